"""Aggregate an .ncu-rep's per-line samples / instructions of the transport kernel by source region.
usage: python tools/ncu_regions.py REP  (regions = functions / blocks of mcs_device.cuh found by marker comments)"""
import csv, io, subprocess, sys, re
rep = sys.argv[1]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
i = 0; secs = []
while i < len(rows):
    if rows[i] and rows[i][0] == "File Path":
        f, fn, Hh = rows[i][1], rows[i+1][1], rows[i+2]; j = i + 3; body = []
        while j < len(rows) and not (rows[j] and rows[j][0] == "File Path"): body.append(rows[j]); j += 1
        secs.append((f, fn, Hh, body)); i = j
    else: i += 1
per = {}
for f, fn, Hh, body in secs:
    if "transport_kernel" not in fn: continue
    ci = {h: k for k, h in enumerate(Hh)}
    key = f.split("/")[-1]
    for r in body:
        try: ln = int(r[0]); s = int(r[ci["# Samples"]]); ie = int(r[ci["Instructions Executed"]]); te = int(r[ci["Thread Instructions Executed"]])
        except Exception: continue
        a = per.setdefault((key, ln), [0, 0, 0]); a[0] += s; a[1] += ie; a[2] += te
    break
ts = sum(v[0] for v in per.values()); ti = sum(v[1] for v in per.values())
srcl = open("montecarloscattering.jl_b200/csrc/mcs_device.cuh").read().splitlines()
# function boundaries in mcs_device.cuh
funcs = [(n + 1, m.group(1)) for n, l in enumerate(srcl) for m in [re.match(r"^(?:template.*>\s*)?__(?:device|global)__.*?\b(\w+)\s*\(", l)] if m]
def func_of(ln):
    name = "?"
    for n, nm in funcs:
        if n <= ln: name = nm
        else: break
    return name
agg = {}
for (key, ln), (s, ie, te) in per.items():
    if key != "mcs_device.cuh": reg = key
    else:
        reg = func_of(ln)
        if reg == "transport_kernel":
            t = srcl[ln - 1]
            reg = "transport_kernel"
    a = agg.setdefault(reg, [0, 0, 0]); a[0] += s; a[1] += ie; a[2] += te
print(f"total samples {ts}, warp-inst {ti/1e9:.2f} G")
for reg, (s, ie, te) in sorted(agg.items(), key=lambda x: -x[1][0]):
    print(f"  {reg:32s} smp {s/ts*100:5.1f}%  inst {ie/ti*100:5.1f}%  lanes {te/max(ie,1):5.1f}")
# transport_kernel body by lane-occupancy class
cl = {"hot (lanes>=24)": [0, 0], "mid (8..24)": [0, 0], "rare (<8)": [0, 0]}
for (key, ln), (s, ie, te) in per.items():
    if key == "mcs_device.cuh" and func_of(ln) == "transport_kernel" and ie:
        l = te / ie
        k = "hot (lanes>=24)" if l >= 24 else ("mid (8..24)" if l >= 8 else "rare (<8)")
        cl[k][0] += s; cl[k][1] += ie
for k, (s, ie) in cl.items(): print(f"  transport_kernel {k:18s} smp {s/ts*100:5.1f}%  inst {ie/ti*100:5.1f}%")
if len(sys.argv) > 2:
    print("\nlines of transport_kernel with < 8 active lanes, by samples")
    lim = float(sys.argv[3]) if len(sys.argv) > 3 else 8
    L = [(s, ie, te, ln) for (key, ln), (s, ie, te) in per.items() if key == "mcs_device.cuh" and func_of(ln) == sys.argv[2] and ie and te / ie < lim]
    for s, ie, te, ln in sorted(L, reverse=True)[:45]:
        print(f"  {ln:4d} smp {s/ts*100:5.2f}% inst {ie/ti*100:5.2f}% lanes {te/ie:5.1f} | {srcl[ln-1].strip()[:110]}")
