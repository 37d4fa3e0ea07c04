"""Aggregate the ncu source page by contiguous SASS regions: executed instructions + stall samples.
usage: python tools/ncu_source_top.py src.csv [bucket]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
bucket = int(sys.argv[2]) if len(sys.argv) > 2 else 64
hdr = rows[1]
ia, isrc, isamp, iexec = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
ins = [(r[isrc].strip(), int(r[isamp] or 0), int(r[iexec] or 0)) for r in rows[2:] if len(r) > iexec]
tot_s = sum(x[1] for x in ins); tot_e = sum(x[2] for x in ins)
print(f"total instr={len(ins)} samples={tot_s} executed={tot_e:.4g}")
for b in range(0, len(ins), bucket):
    chunk = ins[b:b + bucket]
    s = sum(x[1] for x in chunk); e = sum(x[2] for x in chunk)
    ops = {}
    for x in chunk:
        op = x[0].split()[0] if not x[0].startswith("@") else x[0].split()[1]
        ops[op] = ops.get(op, 0) + 1
    top = sorted(ops.items(), key=lambda kv: -kv[1])[:4]
    print(f"[{b:5d}-{b+len(chunk):5d}) samples {100*s/tot_s:5.1f}%  exec {100*e/tot_e:5.1f}%  {top}")
