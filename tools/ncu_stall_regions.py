"""Per-region stall-reason breakdown from an ncu source-page CSV.
usage: python tools/ncu_stall_regions.py src.csv start:end[:name] ..."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
idx = {h: hdr.index(h) for h in reasons}
isamp = hdr.index("# Samples"); iexec = hdr.index("Instructions Executed")
ins = rows[2:]
tot = sum(int(r[isamp] or 0) for r in ins)
for spec in sys.argv[2:]:
    parts = spec.split(":")
    a, b = int(parts[0]), int(parts[1]); name = parts[2] if len(parts) > 2 else spec
    sub = ins[a:b]
    s = sum(int(r[isamp] or 0) for r in sub)
    ex = sum(int(r[iexec] or 0) for r in sub)
    br = {h: sum(int(r[idx[h]] or 0) for r in sub) for h in reasons}
    top = sorted(br.items(), key=lambda kv: -kv[1])[:7]
    print(f"{name:18s} samples {100*s/tot:5.1f}% exec {ex:.3g} | " + "  ".join(f"{k[6:]} {100*v/max(s,1):.0f}%" for k, v in top))
