"""One device-resident batch of pairs for a profiler to look at (ncu -k regex:^sweep_ ... python tools/profile_step.py ...).
usage: python tools/profile_step.py [surf|orb] [n_images] [n_feat] [repeats] [engine]
All n_images*(n_images-1)/2 pairs of a synthetic bank go through ONE esfm_match_pairs call per repeat (one sweep launch)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import easysfm_b200 as esfm
from easysfm_b200 import scheduler, synth

kind = sys.argv[1] if len(sys.argv) > 1 else "surf"
n_images = int(sys.argv[2]) if len(sys.argv) > 2 else 38
n_feat = int(sys.argv[3]) if len(sys.argv) > 3 else 8000
repeats = int(sys.argv[4]) if len(sys.argv) > 4 else 1
engine = sys.argv[5] if len(sys.argv) > 5 else None
dev = torch.device("cuda:0")
ctx = esfm.Context(0)
if engine:
    (ctx.set_l2_engine if kind == "surf" else ctx.set_hamming_engine)(engine)
bank = ctx.bank(esfm.KIND_F32X64 if kind == "surf" else esfm.KIND_B256, n_images)
for f in range(n_images):
    bank.set_frame_rows(f, n_feat)
bank.alloc_device()
ptr, nbytes = bank.device_rows()
raw = scheduler._wrap_device_bytes(ptr, nbytes, 0)
data = (synth.surf_like_torch if kind == "surf" else synth.orb_like_torch)(n_images, n_feat, 4 if kind == "surf" else 5, dev)
raw.copy_(data.reshape(-1).view(torch.uint8))
del data
torch.cuda.synchronize()
bank.commit_device()
pairs = scheduler.all_pairs(n_images)
for r in range(repeats):
    t0 = time.perf_counter()
    res = bank.match_pairs(pairs, 0.8, os.environ.get("PROFILE_CROSS_CHECK", "1") != "0", device_resident=True)
    n = res.n_matches
    res.close()
    ctx.synchronize()
    dt = time.perf_counter() - t0
    st = ctx.stats()
    print(f"{kind} engine={ctx.l2_engine() if kind == 'surf' else ctx.hamming_engine()} {len(pairs)} pairs of {n_feat}x{n_feat}: {n} matches, {dt * 1e3:.1f} ms wall, "
          f"finalize {st['last_finalize_ms']:.2f} ms, sweep {st['last_sweep_ms']:.2f} ms -> {len(pairs) * n_feat * n_feat / (st['last_sweep_ms'] * 1e-3):.3e} cmp/s", flush=True)
bank.close()
ctx.close()
