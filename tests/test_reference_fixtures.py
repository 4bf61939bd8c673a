"""Fixtures recorded from the REAL reference (tools/record_reference.jl, needs Julia) pin the oracle and the CUDA path:
every directory under tests/golden/julia/ is replayed — the recorded uniform streams through every pcut — and compared
pass by pass and at the end of each pcut.  None is committed yet (no Julia runtime exists in the build image): those two
tests SKIP LOUDLY, and parity stays "unpinned by reference outputs" (DESIGN.md).  The consuming machinery itself is
exercised here with a fixture of the same layout written from the oracle."""
import ctypes as C
import glob
import os

import numpy as np
import pytest

import fixture_io
from helpers import LADDER
from mcs_b200 import problem

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
JULIA_FIXTURES = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "julia", "*", "manifest.json")))
NO_FIXTURE = ("no fixture recorded from the Julia reference under tests/golden/julia/ — parity of the oracle is NOT pinned by "
              "reference outputs; run tools/record_reference.jl with Julia >= 1.12 and copy its output there")


def _philox(olib):
    olib.mcso_philox.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.POINTER(C.c_double)]

    def f(seed, gidx, c2, c3, k):
        buf = (C.c_double * max(int(k), 1))()
        olib.mcso_philox(seed, gidx, c2, c3, int(k), buf)
        return np.frombuffer(buf, dtype=np.float64)[: int(k)].copy()
    return f


@pytest.fixture(scope="module")
def oracle_made_fixture(olib, tmp_path_factory):
    run = problem.setup_run(problem.planar_test_particle_input(300, momentum_cutoffs=LADDER[:4]))
    return fixture_io.write_fixture_from_engine(str(tmp_path_factory.mktemp("fx") / "planar"), olib, run, _philox(olib))


def test_fixture_machinery_on_oracle(olib, oracle_made_fixture):
    """Writer -> reader -> replay on the oracle: the recorded Philox streams replayed through every pcut reproduce the run."""
    fixture_io.run_fixture(fixture_io.Fixture(oracle_made_fixture), olib, log=lambda s: None)


@pytest.mark.gpu
def test_fixture_machinery_on_cuda(clib, oracle_made_fixture):
    fixture_io.run_fixture(fixture_io.Fixture(oracle_made_fixture), clib, log=lambda s: None)


@pytest.mark.skipif(not JULIA_FIXTURES, reason=NO_FIXTURE)
@pytest.mark.parametrize("manifest", JULIA_FIXTURES or ["-"])
def test_oracle_against_julia_fixture(olib, manifest):
    fixture_io.run_fixture(fixture_io.Fixture(os.path.dirname(manifest)), olib)


@pytest.mark.gpu
@pytest.mark.skipif(not JULIA_FIXTURES, reason=NO_FIXTURE)
@pytest.mark.parametrize("manifest", JULIA_FIXTURES or ["-"])
def test_cuda_against_julia_fixture(clib, manifest):
    fixture_io.run_fixture(fixture_io.Fixture(os.path.dirname(manifest)), clib)
