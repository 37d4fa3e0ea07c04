"""CPU model (numpy) of the EXPERIMENTAL branch-free SURF row selection of sweep_l2_tc.cu ($ESFM_TC_SURF_BF=1): thread-local column in the
low 5 mantissa bits of every (negative) accumulator, the two largest of a 32-column pass, one merge per pass with the earlier tile winning
ties on the truncated value, final merge of the four column parts by (value, index) keys.  Checks the ALGORITHM against the oracle
(random SURF-like data: mismatches must be float64 near-ties; quantised exact-tie data with duplicates: identical) -- the CUDA code
itself has not run yet (tools/surf_bf_probe.py does that on a GPU).  usage: python tools/surf_bf_emulate.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, oracle
from easysfm_b200 import synth

def emulate(Q, T, delta=2.0**-10):
    nq, nt = len(Q), len(T)
    ntt = (nt + 127)//128
    # ranking value as the sweep computes it (float32 of the exact value; the 3xTF32 noise is not modelled)
    d2 = ((Q[:,None,:].astype(np.float64) - T[None,:,:].astype(np.float64))**2).sum(-1)
    v = (-(0.5*d2 + delta)).astype(np.float32)
    vpad = np.full((nq, ntt*128), -1e30, np.float32); vpad[:, :nt] = v
    bits = vpad.view(np.uint32)
    keys = []   # per part: (value(½d²+δ, float32), idx)
    for part in range(4):
        bf1 = np.full(nq, -3e38, np.float32); bf2 = bf1.copy()
        i1 = np.full(nq, 0xffffffff, np.uint32); i2 = i1.copy()
        for tt in range(ntt):
            col0 = tt*128 + part*32
            w = ((bits[:, col0:col0+32] & np.uint32(0xffffffe0)) | np.arange(32, dtype=np.uint32)[None,:]).view(np.float32)
            srt = -np.sort(-w, axis=1)
            m1, m2 = srt[:,0], srt[:,1]
            enter = m1 > bf2
            a1 = m1 > bf1
            s_new = np.where(a1, m2, m1); s_old = np.where(a1, bf1, bf2); s_oldi = np.where(a1, i1, i2)
            c2 = s_new > s_old
            mb = m1.view(np.uint32); sb = s_new.view(np.uint32)
            nbf1 = np.where(enter & a1, (mb & np.uint32(0xffffffe0)).view(np.float32), bf1)
            ni1 = np.where(enter & a1, col0 + (mb & 31), i1)
            nbf2 = np.where(enter, np.where(c2, (sb & np.uint32(0xffffffe0)).view(np.float32), s_old), bf2)
            ni2 = np.where(enter, np.where(c2, col0 + (sb & 31), s_oldi), i2)
            bf1, bf2, i1, i2 = nbf1.astype(np.float32), nbf2.astype(np.float32), ni1.astype(np.uint32), ni2.astype(np.uint32)
        for b, i in ((bf1, i1), (bf2, i2)):
            val = np.maximum(-b, 0).astype(np.float32)
            ok = (i != 0xffffffff) & (b > -1e29)
            k = (val.view(np.uint32).astype(np.uint64) << np.uint64(32)) | i.astype(np.uint64)
            keys.append(np.where(ok, k, np.uint64(0xffffffffffffffff)))
    K = np.sort(np.stack(keys, 1), axis=1)[:, :2]
    idx = (K & np.uint64(0xffffffff)).astype(np.int64)
    idx[K == np.uint64(0xffffffffffffffff)] = -1
    return idx, d2

def check(Q, T, name):
    idx, d2 = emulate(Q, T)
    ridx, rdist = oracle.knn2(Q, T)
    nt = len(T)
    bad = np.nonzero((idx[:, :min(2,nt)] != ridx[:, :min(2,nt)]).any(1))[0]
    unjust = 0
    d = np.sqrt(d2)
    for r in bad:
        for c in range(min(2, nt)):
            a, b = idx[r,c], ridx[r,c]
            if a != b and (a < 0 or abs(d[r,a]-d[r,b]) > 1e-5*d[r,b]): unjust += 1
    print(f"{name}: {len(Q)}x{nt}: rows with index mismatch {len(bad)}, unjustified {unjust}")

for nq, nt in [(128,128),(1,2),(37,53),(257,1025),(500,700),(300,1)]:
    Q, T = synth.surf_like(2, [nq, nt], seed=nq*3+nt)
    check(Q, T, "surf_like")
for (nq, nt, lv) in ((300,700,3),(129,1025,2),(515,260,5)):
    rng = np.random.default_rng(nq*31+nt)
    Q = (rng.integers(-lv, lv+1, (nq,64))/8.0).astype(np.float32)
    T = (rng.integers(-lv, lv+1, (nt,64))/8.0).astype(np.float32)
    T[7]=T[3]; T[nt-1]=T[3]; Q[11]=T[3]; Q[12]=T[3]
    idx, d2 = emulate(Q, T)
    ridx, _ = oracle.knn2(Q, T)
    print(f"exact ties {nq}x{nt}: identical {bool((idx == ridx).all())}")
