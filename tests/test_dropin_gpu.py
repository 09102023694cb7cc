"""The drop-in claim, exercised with the reference's OWN caller code: test/matvectest.cpp, test/discretization-verification/gridconv.cpp
and src/perftesting/{runcase,scaling_openmp,runtest_openmp}.cpp, compiled UNCHANGED against struct3d_b200/host/include (tests/dropin/
Makefile: only a PETSc/MPI shim header is added) and linked with libstruct3d_host.so, are run here with the reference's own test
configurations (CMake test names in the ids; the control / option files under tests/input/ restate the reference's, whose sources do not
travel to the GPU box).  The programs carry their own assertions (second-order defect, convergence orders, solver convergence)."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
BIN = os.path.join(HERE, "dropin", "_ref")
INPUT = os.path.join(HERE, "input")

CASES = {
    "MatVec": ["matvectest", "poisson_conv.control"],
    "NativeRunCase": ["runcase", "convdiff16_cheb.control", "-options_file", "run_defaults.perc"],
    "ScalingTest_OpenMP": ["runtest_openmp", "convdiff16_cheb.control", "-options_file", "scaling_openmp.perc"],
    "Poisson_NativeGridConv_UniformGrid": ["gridconv", "poisson_conv.control", "-options_file", "native.perc"],
    "CD_NativeGridConv_UniformGrid": ["gridconv", "convdiff_conv.control", "-options_file", "native.perc"],
    "CDCirc_NativeGridConv_UniformGrid": ["gridconv", "cdcirc_conv.control", "-options_file", "native.perc"],
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_reference_driver_runs_unchanged(name):
    args = CASES[name]
    exe = os.path.join(BIN, args[0])
    if not os.path.exists(exe):
        pytest.fail(f"{exe} is missing: build() compiles it from /root/reference where that exists, and the built file travels with the repo")
    argv = [exe] + [os.path.join(INPUT, a) if a.endswith((".control", ".perc")) else a for a in args[1:]]
    out = subprocess.run(argv, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=900)
    text = out.stdout.decode(errors="replace")
    print(text[-2500:])
    assert out.returncode == 0, text[-3000:]
    if args[0] == "runcase":
        assert "Avg solver iters:" in text and "Time taken by native linear solver" in text
    if args[0] == "runtest_openmp":
        rep = open("/tmp/s3d_dropin_scaling-output.txt").read()
        assert len(rep.strip().splitlines()) >= 3, rep
