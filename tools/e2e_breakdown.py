"""Wall-clock breakdown of one end-to-end step through the public API (host frames in, host matches out)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import easysfm_b200 as esfm
from easysfm_b200 import synth

kind = sys.argv[1] if len(sys.argv) > 1 else "surf"
M = int(sys.argv[2]) if len(sys.argv) > 2 else 69
F = int(sys.argv[3]) if len(sys.argv) > 3 else (8000 if kind == "surf" else 4000)
dev = torch.device("cuda:0")
data = (synth.surf_like_torch if kind == "surf" else synth.orb_like_torch)(M, F, 4, dev)
host = torch.empty(data.shape, dtype=data.dtype, pin_memory=True)
host.copy_(data); torch.cuda.synchronize()
frames = host.numpy()
ctx = esfm.Context(0)
kid = esfm.KIND_F32X64 if kind == "surf" else esfm.KIND_B256
for it in range(3):
    t = [time.perf_counter()]
    b = ctx.bank(kid, M); t.append(time.perf_counter())
    for k in range(M):
        b.set_frame(k, frames[k])
    t.append(time.perf_counter())
    b.commit(); t.append(time.perf_counter())
    res = b.match_all_pairs(0.8, True); t.append(time.perf_counter())
    n = res.n_matches
    res.close(); t.append(time.perf_counter())
    b.close(); t.append(time.perf_counter())
    names = ["create", "set_frame x%d" % M, "commit", "match_all_pairs", "results.close", "bank.close"]
    st = ctx.stats()
    print(f"iter {it}: total {1e3*(t[-1]-t[0]):.1f} ms | " + " | ".join(f"{n_} {1e3*(t[i+1]-t[i]):.1f}" for i, n_ in enumerate(names)) +
          f" | sweep {st['last_sweep_ms']:.1f} finalize {st['last_finalize_ms']:.1f} matches {n}")
