"""Summarise an .ncu-rep: key metrics, stall reasons and the hottest CUDA source lines.
usage: python tools/ncu_summary.py gpurun_out/v2.ncu-rep [kernel-substring]
       python tools/ncu_summary.py REP --json profiles/r02_traffic.json WORKLOAD N_PER_PCUT LAUNCH_NO
The second form appends the capture's DRAM traffic and L2 atomic sector counts to the tracked file bench.py reads
`roofline.traffic` from (one record per workload and particle count; an existing record for the same pair is replaced)."""
import csv, io, json, os, subprocess, sys
rep = sys.argv[1]
json_out = None
if len(sys.argv) > 2 and sys.argv[2] == "--json":
    json_out, j_workload, j_n, j_launch = sys.argv[3], sys.argv[4], int(sys.argv[5]), sys.argv[6]
    kern = "transport_kernel"
else:
    kern = sys.argv[2] if len(sys.argv) > 2 else "transport_kernel"
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
H = rows[0]
want = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "smsp__warps_eligible.avg.per_cycle_active",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"]
def _bytes(val, unit):
    mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1.0)
    return float(val.replace(",", "")) * mult
for r in rows[2:]:
    print("kernel:", r[H.index("Kernel Name")][:60])
    for w in want:
        if w in H: print(f"  {w:75s} {r[H.index(w)]} {rows[1][H.index(w)]}")
    if json_out and kern in r[H.index("Kernel Name")]:
        g = lambda m: (r[H.index(m)], rows[1][H.index(m)]) if m in H else (None, None)
        rec = {"workload": j_workload, "n_per_pcut": j_n, "launch": j_launch, "source": os.path.basename(rep),
               "kernel": r[H.index("Kernel Name")][:80],
               "dram_bytes_read": _bytes(*g("dram__bytes_read.sum")), "dram_bytes_write": _bytes(*g("dram__bytes_write.sum")),
               "lts_sectors_red": float(g("lts__t_sectors_op_red.sum")[0].replace(",", "")) if g("lts__t_sectors_op_red.sum")[0] else None,
               "lts_sectors_atom": float(g("lts__t_sectors_op_atom.sum")[0].replace(",", "")) if g("lts__t_sectors_op_atom.sum")[0] else None,
               "gpu_time_ms": g("gpu__time_duration.sum")[0], "gpu_time_unit": g("gpu__time_duration.sum")[1]}
        doc = {"captures": []}
        if os.path.exists(json_out):
            doc = json.load(open(json_out))
        doc["captures"] = [c for c in doc["captures"] if not (c["workload"] == j_workload and int(c["n_per_pcut"]) == j_n)] + [rec]
        json.dump(doc, open(json_out, "w"), indent=1)
        print("wrote", json_out)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
i = 0; secs = []
while i < len(rows):
    if rows[i] and rows[i][0] == "File Path":
        f, fn, Hh = rows[i][1], rows[i+1][1], rows[i+2]; j = i + 3; body = []
        while j < len(rows) and not (rows[j] and rows[j][0] == "File Path"): body.append(rows[j]); j += 1
        secs.append((f, fn, Hh, body)); i = j
    else: i += 1
done = set()
for f, fn, Hh, body in secs:
    if kern not in fn or "mcs_device" not in f or fn in done: continue
    done.add(fn)
    ci = {h: k for k, h in enumerate(Hh)}
    srcfile = open(f).read().splitlines() if __import__("os").path.exists(f) else []
    lines = []; stall = {}
    for r in body:
        if r[0] == "" or r[0] == "Line No": continue
        try:
            ln = int(r[0]); s = int(r[ci["# Samples"]]); ie = int(r[ci["Instructions Executed"]]); te = int(r[ci["Thread Instructions Executed"]])
        except Exception: continue
        lines.append((ln, s, ie, te))
        for h, k in ci.items():
            if h.startswith("stall_") and "Not Issued" not in h:
                try: stall[h] = stall.get(h, 0) + int(r[k])
                except Exception: pass
    ts = sum(l[1] for l in lines); ti = sum(l[2] for l in lines); tt = sum(l[3] for l in lines)
    print(f"\n{fn[:80]}\n  samples {ts} warp-inst {ti/1e9:.2f}G avg active lanes {tt/max(ti,1):.1f}")
    sall = sum(stall.values())
    print("  stalls: " + ", ".join(f"{h[6:]} {v/sall*100:.1f}%" for h, v in sorted(stall.items(), key=lambda x: -x[1])[:8]))
    for ln, s, ie, te in sorted(lines, key=lambda x: -x[2])[:40]:
        txt = srcfile[ln-1].strip()[:90] if 0 < ln <= len(srcfile) else ""
        print(f"  {ln:4d} inst {ie/ti*100:5.1f}% smp {s/ts*100:5.1f}% lanes {te/max(ie,1):5.1f} | {txt}")
