"""Instruction-cache footprint of the transport kernel from an .ncu-rep: SASS instructions by how often they execute
relative to the hottest one, with the address span they cover.  usage: python tools/ncu_icache.py REP"""
import csv, io, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
H = rows[1]; ia, ie, isrc, ismp = H.index("Address"), H.index("Instructions Executed"), H.index("Source"), H.index("# Samples")
ins = []
for r in rows[2:]:
    try: ins.append((int(r[ia], 16), int(r[ie]), r[isrc].strip(), int(r[ismp])))
    except Exception: pass
base = ins[0][0]; mx = max(e for _, e, _, _ in ins); tot = sum(e for _, e, _, _ in ins); ts = sum(s for *_, s in ins)
print(f"{len(ins)} SASS instructions ({len(ins)*16/1024:.0f} KB), hottest executes {mx:.3e} times, {tot/1e9:.2f} G warp-inst")
for lo, hi in ((0.5, 1.01), (0.2, 0.5), (0.05, 0.2), (0.01, 0.05), (0.001, 0.01), (0, 0.001)):
    sel = [(a, e, s) for a, e, _, s in ins if lo * mx <= e < hi * mx]
    if not sel: continue
    # contiguous runs
    runs = 1 + sum(1 for k in range(1, len(sel)) if sel[k][0] - sel[k - 1][0] > 16)
    print(f"  executed {lo:5.3f}..{hi:4.2f} of max: {len(sel):5d} instr = {len(sel)*16/1024:6.1f} KB in {runs:4d} runs, span {(sel[-1][0]-sel[0][0])/1024:7.1f} KB, "
          f"{sum(e for _, e, _ in sel)/tot*100:5.1f}% of executed, {sum(s for *_, s in sel)/ts*100:5.1f}% of samples")
if len(sys.argv) > 3:
    lo, hi = float(sys.argv[2]), float(sys.argv[3])
    prev = None
    for a, e, src, s in ins:
        if lo * mx <= e < hi * mx:
            if prev is not None and a - prev > 16: print("   ...")
            print(f"  +{a-base:6x} x{e/mx:5.2f} smp {s/ts*100:5.2f}% {src[:90]}")
            prev = a
