#!/usr/bin/env python
"""Per-CUDA-source-line instruction / FP64 / stall attribution from an .ncu-rep (needs -lineinfo)."""
import collections, csv, io, re, subprocess, sys

def num(s):
    try:
        return int(s)
    except Exception:
        return 0


rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 50
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
fname = None
hdr = None
lines = {}      # (file, line) -> [src, inst, samples, fp64]
cur = None
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]; continue
    if len(r) >= 2 and r[0] == "Function Name":
        continue
    if r and r[0] == "Line No":
        hdr = r
        iI = hdr.index("Instructions Executed"); iS = hdr.index("# Samples")
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    if r[0] != "-" and r[0] != "":      # a CUDA source line summary row
        cur = (fname, int(r[0]))
        lines.setdefault(cur, [r[1], 0, 0, 0])
        lines[cur][1] += num(r[iI]); lines[cur][2] += num(r[iS])
    else:                               # a SASS row belonging to cur
        sass = r[3]
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", sass)
        if m and cur and m.group(2).split(".")[0] in ("DFMA", "DMUL", "DADD", "DSETP"):
            lines[cur][3] += num(r[iI])
tot = sum(v[1] for v in lines.values()); tots = sum(v[2] for v in lines.values()); totf = sum(v[3] for v in lines.values())
print(f"total inst {tot}  fp64 {totf}  samples {tots}")
byfile = collections.Counter()
for (f, l), v in lines.items():
    byfile[f] += v[1]
print({k: round(v / tot, 3) for k, v in byfile.items()})
for (f, l), v in sorted(lines.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{v[1] / tot:6.3f} inst {v[3] / max(totf, 1):6.3f} fp64 {v[2] / max(tots, 1):6.3f} smp  {f[:22]:22s}:{l:4d}  {v[0][:90]}")
