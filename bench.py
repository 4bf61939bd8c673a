#!/usr/bin/env python
"""bench.py — scattering steps/s and s/iteration of the transport loop (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # CUDA arm; N>1 under torchrun, one rank per GPU
    python bench.py --impl reference --steps K --warmup W    # CPU arm: the C restatement of the reference loop

A "step" is one pass of the hot path over one batch: ONE ITERATION of one ion species — every pcut of
`main_loops.jl:179-317` — on the workload BASELINE.json quotes the metric on (configs[1]: unmodified planar
non-relativistic shock, protons, 1e6 particles per pcut, no smoothing).  Weak scaling: every GPU gets 1e6 particles
per pcut, sharded by contiguous index blocks with global RNG counters; the only collectives are one 8-byte
all-gather per pcut and one all-reduce of the packed tallies per ion (NCCL inside the library).

  value        scattering steps/s, population resident in HBM: total steps of all ranks / max-over-ranks device time of
               the pcut loops (CUDA events on the library's stream).
  e2e          the same through the C-ABI with HOST buffers: H2D of the population from pinned memory and D2H of the
               tallies inside the timed region (wall clock between synchronisations, max over ranks).
  roofline     the transport kernel against the FP64 pipe: steps x 185 flop (SURVEY 8d / DESIGN.md) / kernel time,
               over a DFMA peak measured live on this GPU (MEASURED_PEAKS.json has no FP64 entry).
  cpu_baseline the CPU oracle (C restatement; Julia is not available) on a bounded sample, all host threads.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

FLOP_PER_STEP = 185.0  # SURVEY.md 8d, Appendix D; restated in DESIGN.md
# dram__bytes_read.sum + dram__bytes_write.sum of ONE transport-kernel launch at 1e6 particles per pcut, from the
# `ncu --set full` capture summarised in profiles/r01_v5_transport_kernel_summary.md (154.4 MB + 92.6 MB, launch 7 of 11).
# Algorithmic bytes of that launch: 1e6 particles x (82 B in + 83 B out) = 165 MB; the rest is PSD / log atomics.
NCU_DRAM_BYTES_PER_LAUNCH_1E6 = 247.0e6
METRIC = "scattering_steps_per_sec"


def build_run(workload: str, n_per_pcut: int):
    from mcs_b200 import problem
    mk = {"planar": problem.planar_test_particle_input, "relativistic": problem.relativistic_input,
          "nonlinear": problem.nonlinear_input, "multi": problem.multi_species_input}[workload]
    inp = mk(n_per_pcut)
    inp.num_iterations = 1
    run = problem.setup_run(inp)
    prof = problem.synthetic_precursor(run) if workload == "nonlinear" else run.profile
    return run, prof


def workload_config(workload, run, n_per_pcut, extra=None):
    names = {"planar": "configs[1]: planar non-relativistic test-particle shock (u0=1e4 km/s), protons, unmodified profile",
             "relativistic": "configs[3]: gamma0=10 test-particle shock, protons",
             "nonlinear": "configs[2]: smoothed-precursor shock (r_comp=8), protons",
             "multi": "configs[4]: p+He+e-, gamma0=1.5"}
    c = {"workload": names[workload], "particles_per_pcut_per_gpu": n_per_pcut,
         "pcuts_mpc": [float(p) for p in run.inp.momentum_cutoffs], "n_grid": run.n_grid,
         "psd_bins": [run.num_psd_mom_bins + 2, run.num_psd_theta_bins + 2, run.n_grid],
         "xn_per": [run.inp.fine_scattering_Ng, run.inp.coarse_scattering_Ng], "rng": "philox4x32-10 seed 210",
         "step": "one iteration of one ion species (all pcuts)",
         "population_order": "momentum-sorted injection list dealt into 64 strided sub-sequences (balances contiguous rank shards)"}
    if extra:
        c.update(extra)
    return c


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                       "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            pass

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.p.terminate()
        self.p.wait()
        self.f.flush()
        rows = [r.split(",") for r in open(self.f.name).read().strip().splitlines() if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons, pw = [], [], set(), []
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.strip().lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def cpu_arm(workload, sample_per_pcut, steps, warmup, threads):
    """The reference's CPU implementation of the path = the C restatement (oracle/), all host threads."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_engine
    from mcs_b200 import abi, driver
    lib = oracle_engine.load_oracle_library()
    run, prof = build_run(workload, sample_per_pcut)
    run.profile = prof
    e = abi.Engine(lib, driver.make_config(lib, run, threads=threads, na_cr=1000))
    tot_steps, tot_t = 0, 0.0
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        r = driver.main_loops(run, e, n_iters=1, want_psd=True, want_log=False, shuffle_population=True)[0][0]
        dt = time.perf_counter() - t0
        st = r["tallies"].stats["n_helix_steps"] + r["tallies"].stats["n_retro_steps"]
        if it >= warmup:
            tot_steps += st
            tot_t += dt
    return tot_steps / tot_t, tot_t / max(steps, 1), tot_steps // max(steps, 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="planar", choices=["planar", "relativistic", "nonlinear", "multi"])
    ap.add_argument("--n-per-pcut", type=int, default=1_000_000)
    ap.add_argument("--cpu-sample", type=int, default=10000, help="particles per pcut of the CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--generate-in-library", action="store_true",
                    help="SURVEY 8(f2): hand init_pop to the library in run-length form instead of copying host arrays")
    a = ap.parse_args()
    # stdout carries exactly one JSON line: libraries that write to fd 1 (NCCL prints its version there) go to stderr
    sys.stdout.flush()
    out_fd = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        sys.stdout.flush()
        os.write(out_fd, (json.dumps(obj) + "\n").encode())


    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    threads = os.cpu_count() or 1

    if a.impl == "reference":
        if rank != 0:
            return 0
        v, s_per_it, st = cpu_arm(a.workload, a.cpu_sample, max(a.steps, 1), min(a.warmup, 1), threads)
        run, _ = build_run(a.workload, a.cpu_sample)
        sample = f"{a.cpu_sample} particles per pcut (of {a.n_per_pcut}), full pcut ladder, {st} steps per iteration"
        emit(({
            "impl": "reference", "metric": METRIC, "value": v, "unit": "steps/s", "n_gpus": a.gpus, "steps": a.steps,
            "warmup": min(a.warmup, 1), "ms_per_step": s_per_it * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(a.workload, run, a.n_per_pcut, {"cpu_sample_per_pcut": a.cpu_sample}),
            "cpu_baseline": {"value": v, "unit": "steps/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "C restatement of the reference Julia loop (no Julia runtime in this image), OpenMP over particles",
        }))
        return 0

    import torch
    import mcs_b200
    from mcs_b200 import abi, driver, problem

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the transport loop has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t)
        return float(t.item())

    n_global = a.n_per_pcut * world
    run, prof = build_run(a.workload, n_global)
    run.profile = prof
    lib = mcs_b200.load_cuda_library()
    cfg = driver.make_config(lib, run, n_pts_cap=n_global + 8, na_cr=1_000_000, device=local_rank)
    eng = abi.Engine(lib, cfg)
    if world > 1:
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid = torch.tensor(list(eng.comm_unique_id()), dtype=torch.uint8, device="cuda")
        dist.broadcast(uid, 0)
        eng.comm_init(rank, world, bytes(uid.cpu().tolist()))

    # host inputs of one step, in pinned memory (the population main_loops.jl hands to the particle loop)
    spec = problem.injection_spec(run, prof, 1)
    lo, hi = driver.shard_bounds(spec.n, rank, world)
    pinned, pop = [], {}
    if not a.generate_in_library:
        ip = problem.expand_injection(spec, np.random.default_rng(0), shuffle=True)
        for k, v in ip.pop.items():
            t = torch.from_numpy(np.ascontiguousarray(v[lo:hi])).pin_memory()
            pinned.append(t)
            pop[k] = t.numpy()
    eps = problem.populate_eps_target(run, prof)
    sp = driver.species_struct(run, 1)
    p_hi = problem.pcut_hi(run.inp.en_pcut_hi, run.species[0].mass)
    h2d = sum(v.nbytes for v in pop.values()) + 11 * (run.n_grid + 2) * 8
    if a.generate_in_library:
        h2d += 6 * 8 * len(spec.bin_ptot) + 8
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def one_step():
        """The call sequence a Julia user makes per (iteration, ion): H2D, the device pcut loop, D2H."""
        eng.set_profile(prof, eps, np.zeros(run.n_grid))
        if a.generate_in_library:
            eng.begin_ion_generate(1, 1, sp, spec, first_global=lo, n_local=hi - lo, shuffle=True)
        else:
            eng.begin_ion(1, 1, sp, pop, first_global=lo)
        n_run, n_used, n_saved = eng.run_ion(run.pcuts, p_hi, run.inp.n_pts_pcut, run.inp.n_pts_pcut_hi)
        t = eng.end_ion(want_psd=True, want_log=False)
        return t, n_run

    for _ in range(a.warmup):
        one_step()
    fp64_peak = eng.measure_fp64_peak()
    scatter_peak = eng.measure_scatter_peak()
    atomic_peak = eng.measure_atomic_peak(run.n_grid * (run.num_psd_mom_bins + 2) * (run.num_psd_theta_bins + 2))
    eng.timing(reset=True)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    wall, steps_local, d2h, n_run = 0.0, 0, 0, 0
    barrier()
    for _ in range(a.steps):
        flush.fill_(1)  # L2 flush between timed iterations (not timed)
        barrier()
        t0 = time.perf_counter()
        t, n_run = one_step()
        torch.cuda.synchronize()
        wall += time.perf_counter() - t0
        steps_local = t.stats["n_helix_steps"] + t.stats["n_retro_steps"]  # already summed over ranks by the all-reduce
        d2h = sum(getattr(t, nm).nbytes for nm in ("pxx_flux", "pxz_flux", "energy_flux", "psd", "num_crossings",
                                                  "esc_psd_feb_upstream", "esc_psd_feb_downstream", "esc_energy_eff",
                                                  "esc_num_eff", "weight_coupled", "spectra_coupled",
                                                  "energy_transfer_pool")) + 8 * 8 + 24 * 8
    barrier()
    clocks = sampler.stop() if sampler else None
    tm = eng.timing()
    steps_per_iter = float(steps_local)  # global (counters are all-reduced in mcs_end_ion)
    dev_s = allmax(tm["ion_loop_ms"]) * 1e-3
    wall_s = allmax(wall)
    kern_s = allmax(tm["transport_ms"]) * 1e-3
    per_rank_kernel_ms = per_rank_steps = per_rank_particles = None
    if dist is not None:
        g = [torch.zeros(1, dtype=torch.float64, device="cuda") for _ in range(world)]
        dist.all_gather(g, torch.tensor([tm["transport_ms"]], dtype=torch.float64, device="cuda"))
        per_rank_kernel_ms = [round(float(x.item()) / a.steps, 1) for x in g]
        dist.all_gather(g, torch.tensor([float(tm["local_steps"])], dtype=torch.float64, device="cuda"))
        per_rank_steps = [float(x.item()) / a.steps for x in g]
        dist.all_gather(g, torch.tensor([float(tm["local_particles"])], dtype=torch.float64, device="cuda"))
        per_rank_particles = [float(x.item()) / a.steps for x in g]
    value = steps_per_iter * a.steps / dev_s
    e2e = steps_per_iter * a.steps / wall_s
    # dominant kernel: this rank's steps over this rank's kernel time
    kern_steps_per_s = (steps_per_iter / world) * a.steps / kern_s
    achieved = kern_steps_per_s * FLOP_PER_STEP / 1e12
    launches = int(tm["transport_launches"] + tm["other_launches"])

    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": "steps/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": dev_s / a.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": workload_config(a.workload, run, a.n_per_pcut, {
                "population": "generated in the library from the run-length injection list" if a.generate_in_library
                              else "host arrays in pinned memory, copied every step",
                "pcuts_run": int(n_run), "steps_per_iteration": int(steps_per_iter), "l2": "flushed between timed iterations",
                "s_per_iteration_device": dev_s / a.steps, "s_per_iteration_e2e": wall_s / a.steps}),
            "e2e": {"value": e2e, "unit": "steps/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": wall_s / a.steps * 1e3},
            "gpu_launches": launches, "per_rank_kernel_ms_per_step": per_rank_kernel_ms,
            "per_rank_steps_per_step": per_rank_steps, "per_rank_particles_per_step": per_rank_particles,
            "clocks": clocks,
            "roofline": {"bound": "fp64", "kernel": "transport_kernel<false>", "achieved": achieved, "peak": fp64_peak,
                         "unit": "TFLOP/s", "frac": achieved / fp64_peak if fp64_peak else None,
                         "traffic": NCU_DRAM_BYTES_PER_LAUNCH_1E6 * a.n_per_pcut / 1e6 if a.workload == "planar" else None,
                         "traffic_source": "ncu --set full capture at 1e6 particles per pcut (profiles/), scaled by particle count",
                         "flop_per_step": FLOP_PER_STEP, "kernel_ms_per_launch": kern_s * 1e3 / max(tm["transport_launches"], 1),
                         "kernel_share_of_step": kern_s / dev_s,
                         "peak_source": "measured live: DFMA microbenchmark in libmcs_b200.so (MEASURED_PEAKS.json has no FP64 entry)",
                         "hot_path_ceiling_steps_per_s": scatter_peak,
                         "frac_of_hot_path_ceiling": kern_steps_per_s / scatter_peak if scatter_peak else None,
                         "hot_path_ceiling_source": "measured live: scatter_only_kernel = Philox + kick + phase + move, no control flow",
                         "fp64_red_scattered_gops": atomic_peak},
        }
        if world == 1 and not a.no_cpu_baseline:
            v, s_it, st = cpu_arm(a.workload, a.cpu_sample, 1, 0, threads)
            out["cpu_baseline"] = {"value": v, "unit": "steps/s", "cores": threads, "kind": "port",
                                   "sample": f"{a.cpu_sample} particles per pcut, full pcut ladder, {st} steps, {s_it:.1f} s"}
            # the reference loop itself is serial (the threading directive at main_loops.jl:227 is a comment): one thread too
            n1 = max(a.cpu_sample // 10, 200)
            v1, s1, st1 = cpu_arm(a.workload, n1, 1, 0, 1)
            out["cpu_baseline"]["single_thread"] = {"value": v1, "cores": 1,
                                                    "sample": f"{n1} particles per pcut, {st1} steps, {s1:.1f} s"}
        emit(out)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
