#!/usr/bin/env python
"""hottest CUDA source lines (warp-stall samples) of the kernels in an .ncu-rep:
   python scripts/ncu_hot.py rep.ncu-rep [N]     (needs -lineinfo and --import-source on)"""
import csv, io, subprocess, sys
rep = sys.argv[1]; N = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = None; cur = None; agg = {}; fname = ""; kname = ""
def flush():
    global agg
    if not agg: return
    tot = sum(v[0] for v in agg.values()) or 1
    files = {}
    for (f, ln, src), v in agg.items(): files[f] = files.get(f, 0) + v[0]
    print("== %s  [%s]  samples=%d" % (kname[:100], max(files, key=files.get), tot))
    for (f, ln, src), v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:N]:
        st = sorted(v[2].items(), key=lambda kv: -kv[1])[:3]
        print("%6.2f%% ex=%-10d %-44s | %4s %s" % (100.0 * v[0] / tot, v[1], ",".join("%s:%d" % (k[6:], n) for k, n in st if n), ln, src.strip()[:100]))
    agg = {}
for r in rows:
    if not r: continue
    if r[0] in ("Kernel Name", "Function Name"):
        if r[1] != kname: flush()
        kname = r[1]; continue
    if r[0] in ("File Name", "File Path"): fname = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; isamp = hdr.index("# Samples"); iex = hdr.index("Instructions Executed"); stalls = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]; continue
    if hdr is None or len(r) < len(hdr) - 2: continue
    if r[0] != "": cur = (fname, r[0], r[1]); continue   # a CUDA source line; its SASS rows follow
    if cur is None: continue
    try: sm = int(r[isamp] or 0); ex = int(r[iex] or 0)
    except ValueError: continue
    a = agg.setdefault(cur, [0, 0, {}])
    a[0] += sm; a[1] += ex
    for i in stalls:
        try: a[2][hdr[i]] = a[2].get(hdr[i], 0) + int(r[i] or 0)
        except (ValueError, IndexError): pass
flush()
