"""Summarise an `ncu --page source --csv --print-source cuda,sass` export: per CUDA source line, the warp
instructions executed and the stall samples, for the first launch of each kernel in the file.

    ncu -i rep.ncu-rep --page source --csv --print-source cuda,sass -k regex:<kernel> > src.csv
    python profiles/srcsum.py src.csv [top_n] [launch_index]
"""
import csv
import sys
from collections import defaultdict

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0

rows = list(csv.reader(open(path)))
# the export is a sequence of (File Path, Function Name, header, rows...) blocks, one per source file per
# launch; per-line values sit on the CUDA-line rows (the SASS rows under them repeat the text only)
blocks, cur, fpath, func = [], None, "", ""
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fpath = r[1]
    elif r[0] == "Function Name":
        func = r[1]
    elif r[0] == "Line No":
        cur = {"hdr": r, "rows": [], "file": fpath.split("/")[-1], "func": func}
        blocks.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
funcs = []
for b in blocks:
    if b["func"] not in funcs:
        funcs.append(b["func"])
sel = funcs[min(which, len(funcs) - 1)]
print("function:", sel[:120])


def num(x):
    try:
        return int(float(x.replace(",", "")))
    except Exception:
        return 0


per = {}
tot_inst = tot_samp = 0
for b in blocks:
    if b["func"] != sel:
        continue
    h = b["hdr"]
    src_i = h.index("Source")
    inst_i = h.index("Instructions Executed")
    samp_i = h.index("# Samples")
    thr_i = h.index("Avg. Threads Executed")
    stall_cols = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
    for r in b["rows"]:
        if not r[0].strip().isdigit() or len(r) <= max(stall_cols):
            continue
        key = (b["file"], int(r[0]))
        i, s = num(r[inst_i]), num(r[samp_i])
        st = {h[c]: num(r[c]) for c in stall_cols if num(r[c])}
        e = per.setdefault(key, [0, 0, r[src_i].strip(), {}, r[thr_i]])
        e[0] += i
        e[1] += s
        for k, v in st.items():
            e[3][k] = e[3].get(k, 0) + v
        tot_inst += i
        tot_samp += s
print(f"total warp instructions {tot_inst}  samples {tot_samp}")
for (f, ln), (i, s, txt, st, thr) in sorted(per.items(), key=lambda kv: -kv[1][1])[:top]:
    tops = ", ".join(f"{k[6:]}={v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(f"{f}:{ln:<4d} inst {100.0 * i / max(tot_inst, 1):5.1f}%  samp {100.0 * s / max(tot_samp, 1):5.1f}%  thr {thr:>5s} | {tops} | {txt[:90]}")
