"""CPU tests that pin the oracle: the C restatement (oracle/bf_oracle.c) against the reference's own cv2
calls (oracle/cv2_oracle.py) run live here, and against the committed golden fixtures (test_golden.py)."""
import numpy as np
import pytest

import oracle
from oracle import cv2_oracle
from easysfm_b200 import synth
from util import assert_matches_equal, dist64, justify_l2, check_knn_l2


@pytest.mark.parametrize("nq,nt", [(1, 2), (5, 1), (0, 7), (64, 64), (300, 517), (1000, 800)])
def test_hamming_oracle_equals_cv2(nq, nt):
    Q, T = synth.orb_like(2, [nq, nt], seed=nq + nt)
    idx, dist = oracle.knn2(Q, T)
    cidx, cdist = cv2_oracle.knn2(Q, T)
    np.testing.assert_array_equal(idx, cidx)
    np.testing.assert_array_equal(dist, cdist)
    for ratio in (0.5, 0.8, 0.7):
        for cc in (False, True):
            assert_matches_equal(oracle.match(Q, T, ratio, cc), cv2_oracle.match(Q, T, ratio, cc))


@pytest.mark.parametrize("nq,nt", [(1, 2), (5, 1), (0, 7), (64, 64), (300, 517), (1000, 800)])
def test_l2_oracle_vs_cv2(nq, nt):
    Q, T = synth.surf_like(2, [nq, nt], seed=nq + nt)
    idx, dist = oracle.knn2(Q, T)
    cidx, cdist = cv2_oracle.knn2(Q, T)
    if nq and nt:
        D = dist64(Q, T)
        check_knn_l2(Q, T, idx[:, : min(2, nt)], dist[:, : min(2, nt)], D)
        check_knn_l2(Q, T, cidx[:, : min(2, nt)], cdist[:, : min(2, nt)], D)
        same = idx == cidx
        assert same.mean() > 0.999
        np.testing.assert_allclose(dist[same], cdist[same], rtol=2.4e-7)   # summation-order noise only
    for ratio in (0.5, 0.8):
        for cc in (False, True):
            a, b = oracle.match(Q, T, ratio, cc), cv2_oracle.match(Q, T, ratio, cc)
            if nq and nt:
                justify_l2(Q, T, ratio, cc, a, b)
            else:
                assert len(a) == len(b) == 0


def test_tie_break_lowest_index_both_ranks():
    """a query equal to three identical train rows gets the two LOWEST indices, in order."""
    rng = np.random.default_rng(0)
    T = rng.standard_normal((9, 64)).astype(np.float32)
    T[4] = T[1]; T[5] = T[1]
    Q = T[[1]].copy()
    for mod in (oracle, cv2_oracle):
        idx, dist = mod.knn2(Q, T)
        assert idx[0].tolist() == [1, 4] and dist[0].tolist() == [0.0, 0.0]
    Tb = rng.integers(0, 256, (9, 32), dtype=np.uint8)
    Tb[3] = Tb[0]
    Qb = Tb[[3]].copy(); Qb[0, 0] ^= 0x80
    for mod in (oracle, cv2_oracle):
        idx, dist = mod.knn2(Qb, Tb)
        assert idx[0].tolist() == [0, 3] and dist[0].tolist() == [1.0, 1.0]


def test_cross_check_is_strict_mutual_nn():
    """queries A=0, B=1 and trains j1=0.6, j2=-1 on one axis: only (B, j1) survives."""
    Q = np.zeros((2, 64), np.float32); Q[1, 0] = 1.0
    T = np.zeros((2, 64), np.float32); T[0, 0] = 0.6; T[1, 0] = -1.0
    for mod in (oracle, cv2_oracle):
        m = mod.mutual_nn(Q, T)
        assert [(int(x["queryIdx"]), int(x["trainIdx"])) for x in m] == [(1, 0)]
    # duplicate queries: only the lower index is matched; duplicate trains: matched to the lower index
    Q2 = np.zeros((2, 64), np.float32); Q2[:, 1] = 1.0
    T2 = np.zeros((2, 64), np.float32); T2[:, 1] = 0.9
    for mod in (oracle, cv2_oracle):
        m = mod.mutual_nn(Q2, T2)
        assert [(int(x["queryIdx"]), int(x["trainIdx"])) for x in m] == [(0, 0)]


def test_ratio_is_evaluated_in_double():
    """d1 = 44, d2 = 55, ratio 0.8 -> 0.8 * 55.0 == 44.0 in double, so `<` fails."""
    T = np.zeros((2, 32), np.uint8)
    Q = np.zeros((1, 32), np.uint8)
    T[0, :5] = 0xFF; T[0, 5] = 0x0F          # 44 bits
    T[1, :6] = 0xFF; T[1, 6] = 0x7F          # 55 bits
    for mod in (oracle, cv2_oracle):
        idx, dist = mod.knn2(Q, T)
        assert dist[0].tolist() == [44.0, 55.0]
        assert len(mod.match(Q, T, 0.8, False)) == 0
        assert len(mod.match(Q, T, 0.8000001, False)) == 1


def test_rows_lt_2_defined_as_no_match():
    Q = synth.orb_like(1, 10, seed=1)[0]
    assert len(oracle.match(Q, Q[:1], 0.8, False)) == 0
    assert len(cv2_oracle.match(Q, Q[:1], 0.8, False)) == 0
    assert len(oracle.match(Q[:0], Q, 0.8, True)) == 0


def test_oracle_thread_count_does_not_change_results():
    Q, T = synth.surf_like(2, [200, 300], seed=3)
    oracle.set_threads(1)
    a = oracle.match(Q, T, 0.8, True)
    oracle.set_threads(0)
    b = oracle.match(Q, T, 0.8, True)
    assert_matches_equal(a, b)


def _cv2_cross_check_call(Q, T):
    """The reference's own mutual-NN call, python_code/feature_match.py:26-27, issued directly."""
    import cv2
    norm = cv2.NORM_L2 if Q.dtype == np.float32 else cv2.NORM_HAMMING
    m = cv2.BFMatcher(norm, crossCheck=True).match(np.ascontiguousarray(Q), np.ascontiguousarray(T))
    return sorted((x.queryIdx, x.trainIdx, x.distance) for x in m)


def _as_tuples(m):
    return [(int(x["queryIdx"]), int(x["trainIdx"]), float(x["distance"])) for x in m]


@pytest.mark.parametrize("kind", ["surf", "orb"])
def test_ratio_inf_is_no_ratio_test_even_for_exact_duplicates(kind):
    """ratio = +inf means NO ratio test (include/esfm_match.h), not `d1 < inf * d2`: that product is NaN when the second
    neighbour is an exact duplicate (d2 = 0) and would drop matches that BFMatcher(crossCheck=True).match keeps.  Both
    oracles in this mode must equal that cv2 call, duplicates on both sides included; one train row is enough."""
    if kind == "surf":
        rng = np.random.default_rng(5)
        Q = (rng.integers(-3, 4, (90, 64)) / 8.0).astype(np.float32)
        T = (rng.integers(-3, 4, (140, 64)) / 8.0).astype(np.float32)
    else:
        Q, T = synth.orb_like(2, [90, 140], seed=6)
    T[7] = T[3]; T[139] = T[3]; Q[11] = T[3]; Q[12] = T[3]
    want = _cv2_cross_check_call(Q, T)
    assert (11, 3, 0.0) in want and not any(q == 12 for q, _, _ in want)
    for mod in (oracle, cv2_oracle):
        assert _as_tuples(mod.match(Q, T, float("inf"), True)) == want
        assert _as_tuples(mod.mutual_nn(Q, T)) == want
        fwd = mod.match(Q, T, float("inf"), False)          # every query keeps its nearest neighbour
        idx, dist = mod.knn2(Q, T)
        assert len(fwd) == len(Q) and (fwd["trainIdx"] == idx[:, 0]).all() and (fwd["distance"] == dist[:, 0]).all()
        one = mod.match(Q, T[:1], float("inf"), True)        # a single train row: its nearest query, nothing else
        assert _as_tuples(one) == _cv2_cross_check_call(Q, T[:1])
        assert len(mod.match(Q, T[:1], 1e30, True)) == 0     # a finite ratio still needs two neighbours
