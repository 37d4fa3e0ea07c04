"""Match persistence: the match file written by esfm_results_save / read by esfm_results_load.
CPU: the format against a numpy writer/reader, lookups on a loaded batch, error behaviour, the C++ shim resuming from a
file with no device.  GPU: match -> save -> load gives the same batch."""
import os
import subprocess

import numpy as np
import pytest

import easysfm_b200 as esfm
from test_shim import build_shim

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _fake_batch(seed=0, n_frames=6):
    rng = np.random.default_rng(seed)
    pairs = np.array([(i, j) for i in range(n_frames) for j in range(i)], np.int32)
    per_pair = []
    for _ in pairs:
        n = int(rng.integers(0, 40))
        m = np.zeros(n, esfm.DMATCH_DTYPE)
        m["queryIdx"] = np.sort(rng.choice(500, n, replace=False))
        m["trainIdx"] = rng.integers(0, 500, n)
        m["distance"] = rng.random(n).astype(np.float32)
        per_pair.append(m)
    return pairs, per_pair


def test_load_numpy_written_file_and_lookups(tmp_path):
    pairs, per_pair = _fake_batch()
    path = str(tmp_path / "batch.matches")
    esfm.write_match_file(path, pairs, per_pair, esfm.KIND_B256, 0.8, True)
    r = esfm.load_results(path)
    assert r.n_pairs == len(pairs) and r.n_matches == sum(len(m) for m in per_pair)
    assert r.params() == (esfm.KIND_B256, 0.8, True)
    assert (r.pair_counts() == [len(m) for m in per_pair]).all()
    for k, (q, t) in enumerate(pairs):
        qq, tt, m = r.pair_at(k)
        assert (qq, tt) == (q, t) and m.tobytes() == per_pair[k].tobytes()
        assert r.pair(int(q), int(t)).tobytes() == per_pair[k].tobytes()
    with pytest.raises(esfm.EsfmError):
        r.pair(0, 5)                                   # not part of the batch (query > train only)
    # the library's writer produces byte-for-byte the same file
    path2 = str(tmp_path / "again.matches")
    r.save(path2)
    assert open(path, "rb").read() == open(path2, "rb").read()
    hdr, p2, c2, m2 = esfm.read_match_file(path2)
    assert hdr["magic"] == b"ESFMMTCH" and (p2 == pairs).all() and m2.tobytes() == np.concatenate(per_pair).tobytes()
    r.close()


def test_version2_file_carries_frame_rows_and_validates(tmp_path):
    """Version 2 stores the bank's per-frame row counts; esfm_results_validate refuses a batch that does not fit the caller's
    frames (other row counts, fewer frames, or -- for version-1 files without row counts -- any out-of-range match index)."""
    n_frames = 6
    pairs, per_pair = _fake_batch(2, n_frames)
    rows = [500, 500, 501, 640, 500, 777]
    v2 = str(tmp_path / "v2.matches")
    esfm.write_match_file(v2, pairs, per_pair, esfm.KIND_F32X64, 0.5, False, frame_rows=rows)
    r = esfm.load_results(v2)
    assert r.frame_rows().tolist() == rows
    r.validate(rows)
    with pytest.raises(esfm.EsfmError):
        r.validate(rows[:-1] + [778])               # a frame changed size since the batch was matched
    with pytest.raises(esfm.EsfmError):
        r.validate(rows[:-1])                       # fewer frames
    again = str(tmp_path / "v2_again.matches")
    r.save(again)
    assert open(v2, "rb").read() == open(again, "rb").read()
    hdr, p2, c2, m2 = esfm.read_match_file(again)
    assert int(hdr["version"]) == 2 and (p2 == pairs).all() and m2.tobytes() == np.concatenate(per_pair).tobytes()
    r.close()
    v1 = str(tmp_path / "v1.matches")
    esfm.write_match_file(v1, pairs, per_pair, esfm.KIND_F32X64, 0.5, False)
    r1 = esfm.load_results(v1)
    assert len(r1.frame_rows()) == 0
    r1.validate(rows)                               # every index < 500: fits
    with pytest.raises(esfm.EsfmError):
        r1.validate([500, 500, 501, 640, 500, 10])  # frame 5 has 10 rows now: some queryIdx of pairs (5, j) is out of range
    with pytest.raises(esfm.EsfmError):
        r1.validate(rows[:4])                       # pairs reference frames 4 and 5
    r1.close()


def test_empty_batch_and_bad_files(tmp_path):
    path = str(tmp_path / "empty.matches")
    esfm.write_match_file(path, np.zeros((0, 2), np.int32), [], esfm.KIND_F32X64, 0.5, False)
    r = esfm.load_results(path)
    assert r.n_pairs == 0 and r.n_matches == 0
    r.close()
    with pytest.raises(esfm.EsfmError):
        esfm.load_results(str(tmp_path / "missing.matches"))
    bad = str(tmp_path / "foreign.matches")
    open(bad, "wb").write(b"not a match file at all, just some bytes" * 4)
    with pytest.raises(esfm.EsfmError):
        esfm.load_results(bad)
    pairs, per_pair = _fake_batch(1)
    full = str(tmp_path / "full.matches")
    esfm.write_match_file(full, pairs, per_pair, esfm.KIND_B256, 0.8, False)
    trunc = str(tmp_path / "trunc.matches")
    open(trunc, "wb").write(open(full, "rb").read()[:-24])
    with pytest.raises(esfm.EsfmError):
        esfm.load_results(trunc)


def test_shim_resumes_from_match_file_without_a_device(tmp_path):
    """esfm_load_matches + matchFeaturesORB lookups in the reference's loop order, on this (GPU-less) box."""
    exe = build_shim(str(tmp_path))
    rows = [50, 0, 30, 44]
    pairs = np.array([(i, j) for i in range(len(rows)) for j in range(i)], np.int32)
    rng = np.random.default_rng(3)
    per_pair = []
    for q, t in pairs:
        n = min(rows[q], rows[t], int(rng.integers(0, 20)))
        m = np.zeros(n, esfm.DMATCH_DTYPE)
        m["queryIdx"] = np.sort(rng.choice(max(rows[q], 1), n, replace=False))
        m["trainIdx"] = rng.integers(0, max(rows[t], 1), n)
        m["distance"] = rng.integers(0, 256, n).astype(np.float32)
        per_pair.append(m)
    out = str(tmp_path / "out.bin")
    esfm.write_match_file(out + ".matches", pairs, per_pair, esfm.KIND_B256, 0.8, False)
    blob = str(tmp_path / "desc.bin")
    open(blob, "wb").write(bytes(sum(rows) * 256))     # enough for either kind; the lookups never read it
    subprocess.check_call([exe, "O", str(len(rows))] + [str(r) for r in rows] + [blob, "3", out], stdout=subprocess.DEVNULL)
    raw = open(out, "rb").read()
    pos = 0
    for k in range(len(pairs)):
        cnt = int(np.frombuffer(raw, np.int32, 1, pos)[0]); pos += 4
        assert raw[pos:pos + 16 * cnt] == per_pair[k].tobytes(); pos += 16 * cnt
    assert pos == len(raw)
    # a file matched with another ratio is refused (the shim's SURF default ratio is 0.5, the file says 0.8 / ORB)
    assert subprocess.call([exe, "S", str(len(rows))] + [str(r) for r in rows] + [blob, "3", out],
                           stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL) == 7
    # ... and so is a file that belongs to other frames: stored row counts differ / an index would be out of range
    esfm.write_match_file(out + ".matches", pairs, per_pair, esfm.KIND_B256, 0.8, False, frame_rows=[50, 0, 31, 44])
    assert subprocess.call([exe, "O", str(len(rows))] + [str(r) for r in rows] + [blob, "3", out],
                           stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL) == 7
    esfm.write_match_file(out + ".matches", pairs, per_pair, esfm.KIND_B256, 0.8, False)
    small = [5, 0, 3, 4]
    assert subprocess.call([exe, "O", str(len(small))] + [str(r) for r in small] + [blob, "3", out],
                           stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL) == 7


@pytest.mark.gpu
def test_save_load_roundtrip_gpu(ctx, tmp_path):
    from easysfm_b200 import synth
    for frames in (synth.orb_like(4, [300, 0, 257, 190], seed=5), synth.surf_like(4, [300, 0, 257, 190], seed=5)):
        res = ctx.bank_from_frames(frames).match_all_pairs(0.8, True)
        path = str(tmp_path / "gpu.matches")
        res.save(path)
        back = esfm.load_results(path)
        assert back.n_pairs == res.n_pairs and back.n_matches == res.n_matches and back.params() == res.params()
        for k in range(res.n_pairs):
            a, b = res.pair_at(k), back.pair_at(k)
            assert a[:2] == b[:2] and a[2].tobytes() == b[2].tobytes()
