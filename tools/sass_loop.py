"""Instruction mix of the hottest FFMA2/POPC loop in a cubin/object: finds backward branches and reports the
body that contains the most `key` instructions.  usage: python tools/sass_loop.py file.o [key=FFMA2]"""
import re, subprocess, sys
obj = sys.argv[1]; key = sys.argv[2] if len(sys.argv) > 2 else "FFMA2"
txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
cur = []; kernels = {}
name = None
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = m.group(1); kernels[name] = []; continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m and name:
        kernels[name].append((int(m.group(1), 16), m.group(2).strip()))
for name, ins in kernels.items():
    addr_idx = {a: i for i, (a, _) in enumerate(ins)}
    best = None
    for i, (a, s) in enumerate(ins):
        m = re.search(r"BRA\S*\s+(?:\S+,\s*)?(0x[0-9a-f]+)", s)
        if m:
            t = int(m.group(1), 16)
            if t < a and t in addr_idx:
                body = ins[addr_idx[t]: i + 1]
                n = sum(1 for _, x in body if re.search(r"\b" + re.escape(key) + r"\b", x))
                if n and (best is None or n > best[0] or (n == best[0] and len(body) < len(best[1]))):
                    best = (n, body)
    if best:
        ops = {}
        for _, x in best[1]:
            op = x.split()[1] if x.startswith("@") else x.split()[0]
            ops[op] = ops.get(op, 0) + 1
        print(name[:60], "loop body:", len(best[1]), "instrs;", sorted(ops.items(), key=lambda kv: -kv[1])[:10])
