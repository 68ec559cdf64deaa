#!/usr/bin/env python
"""Warp-stall summary from the source page of an .ncu-rep: python tools/ncu_stalls.py rep [topN]"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = collections.Counter(); byop = collections.Counter(); execd = collections.Counter(); samples = 0
lines = []
for r in rows[2:]:
    if len(r) < len(hdr): continue
    try: n = int(r[idx["# Samples"]])
    except ValueError: continue
    samples += n
    for s in stalls:
        try: tot[s] += int(r[idx[s]])
        except ValueError: pass
    toks = r[idx["Source"]].split()
    op = toks[1] if toks and toks[0].startswith("@") and len(toks) > 1 else (toks[0] if toks else "")
    op = op.split(".")[0]
    byop[op] += n
    try: execd[op] += int(r[idx["Instructions Executed"]])
    except ValueError: pass
    lines.append((n, r[idx["Address"]], r[idx["Source"]], {s: r[idx[s]] for s in stalls if r[idx[s]] not in ("0", "")}))
print("samples", samples)
print("stalls:", "  ".join(f"{s[6:]}:{100*v/samples:.1f}%" for s, v in tot.most_common(10)))
print("samples by op:", "  ".join(f"{k}:{100*v/samples:.1f}%" for k, v in byop.most_common(14)))
te = sum(execd.values())
print("executed by op:", "  ".join(f"{k}:{100*v/te:.1f}%" for k, v in execd.most_common(16)), " total", te)
for n, a, s, st in sorted(lines, key=lambda x: -x[0])[:top]:
    print(n, a, s[:70], st)
