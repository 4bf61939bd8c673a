"""N>1 host logic on CPU: world_size-2 gloo run of the sharded driver against the single-rank run.

Sharding is by contiguous blocks of the particle index range with GLOBAL indices kept for the RNG counter
(SURVEY 8e), so every particle's trajectory is identical at any rank count; only the tally summation order
changes.  The compute engine in this test is the CPU oracle (test infrastructure); on the GPU the same
driver runs with the CUDA engine and NCCL inside the library (tests/test_parity_gpu.py)."""
import os
import pickle
import socket
import subprocess
import sys
import tempfile

import numpy as np
import pytest

from helpers import LADDER, make_engine, rel_close
from mcs_b200 import driver, problem

HERE = os.path.dirname(os.path.abspath(__file__))


def _inp():
    return problem.planar_test_particle_input(400, momentum_cutoffs=LADDER[:4])


def test_shard_bounds_cover_range():
    for n in (0, 1, 7, 400, 1001):
        for w in (1, 2, 3, 8):
            b = [driver.shard_bounds(n, r, w) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n and all(b[i][1] == b[i + 1][0] for i in range(w - 1))


def _worker(rank, world, port, out):
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(os.path.dirname(HERE), "oracle"))
    import oracle_engine
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    run = problem.setup_run(_inp())
    e = make_engine(oracle_engine.load_oracle_library(), run, bin_thermal=True)
    comm = driver.TorchComm()
    r = driver.main_loops(run, e, n_iters=1, comm=comm, thermo=True)[0][0]
    if rank == 0:
        t = r["tallies"]
        pickle.dump(dict(pxx=r["pxx_flux"], en=r["energy_flux"], psd=t.psd, stats=t.stats, scalars=t.scalars,
                         n_used=r["n_used"], n_saved=r["n_saved"], ncross=t.num_crossings, esc=t.esc_psd_feb_downstream,
                         thpf=t.therm_d2N_pf, P_par=r["P_psd_par"], P_perp=r["P_psd_perp"], e_dens=r["energy_density_psd"]),
                    open(out, "wb"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_ranks_match_one_rank(olib):
    run = problem.setup_run(_inp())
    one = driver.main_loops(run, make_engine(olib, run, bin_thermal=True), n_iters=1, thermo=True)[0][0]
    host = driver.main_loops(run, make_engine(olib, run), n_iters=1, host_pcut_loop=True)[0][0]
    assert np.array_equal(one["n_saved"], host["n_saved"]) and np.array_equal(one["pxx_flux"], host["pxx_flux"])
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    with tempfile.TemporaryDirectory() as d:
        out = os.path.join(d, "r0.pkl")
        code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r); import test_driver_multirank as m; "
                "m._worker(int(sys.argv[1]), 2, %d, %r)" % (HERE, os.path.dirname(HERE), port, out))
        procs = [subprocess.Popen([sys.executable, "-c", code, str(r)]) for r in range(2)]
        for p in procs:
            assert p.wait(timeout=500) == 0
        two = pickle.load(open(out, "rb"))
    t = one["tallies"]
    assert np.array_equal(two["n_used"], one["n_used"]) and np.array_equal(two["n_saved"], one["n_saved"])
    assert two["stats"] == t.stats                      # identical trajectories: every integer count agrees
    assert np.array_equal(two["ncross"], t.num_crossings)
    assert rel_close(one["pxx_flux"], two["pxx"], 0) < 1e-12 and rel_close(one["energy_flux"], two["en"], 0) < 1e-11
    assert np.array_equal(t.psd != 0, two["psd"] != 0) and rel_close(t.psd, two["psd"], 0) < 1e-11
    assert rel_close(t.esc_psd_feb_downstream, two["esc"], 0) < 1e-11
    for k, v in t.scalars.items():
        assert two["scalars"][k] == pytest.approx(v, rel=1e-11)
    # SURVEY 8(f1) across ranks: the thermal histograms are summed on the host like the other tallies, and the pressure consumer
    # is fed the summed arrays (each rank's resident tallies are only its own share): same pressures as on one rank
    assert rel_close(t.therm_d2N_pf, two["thpf"], 0) < 1e-11
    for a, b in ((one["P_psd_par"], two["P_par"]), (one["P_psd_perp"], two["P_perp"]), (one["energy_density_psd"], two["e_dens"])):
        assert np.all(np.isfinite(b)) and rel_close(a, b, 0) < 1e-10
