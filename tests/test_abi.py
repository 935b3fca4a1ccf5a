"""CPU tier: the C-ABI library loads, exports every symbol include/ldeq.h declares, and its option struct has
the layout the Python binding assumes.  No compute calls are made here (there is no GPU in this tier)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "ldeq.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ldeq_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported_and_bound(ldeq):
    syms = _header_symbols()
    assert syms == sorted(ldeq._cabi.SYMBOLS)
    lib = ctypes.CDLL(ldeq._cabi.LIB_PATH)
    for s in syms:
        assert hasattr(lib, s), f"libldeq.so does not export {s}"


def test_no_torch_or_oracle_in_the_abi_library(ldeq):
    import subprocess
    out = subprocess.run(["ldd", ldeq._cabi.LIB_PATH], capture_output=True, text=True).stdout
    assert "torch" not in out and "oracle" not in out and "libcuda.so" not in out
    syms = subprocess.run(["nm", "-D", "--undefined-only", ldeq._cabi.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle_" not in syms


def test_opts_layout_and_defaults(ldeq):
    o = ldeq.default_opts()
    assert (o.abstol, o.reltol, o.adaptive, o.controller_pow) == (1e-6, 1e-3, 1, 0)
    assert (o.gamma, o.qmin, o.qmax, o.qoldinit, o.qsteady_min, o.qsteady_max) == (0.9, 0.2, 10.0, 1e-4, 1.0, 1.0)
    assert abs(o.beta1 - 7 / 50) < 1e-16 and abs(o.beta2 - 2 / 25) < 1e-16 and o.maxiters == 1000000
    assert o.dt == 0.0 and o.dtmax == 0.0 and o.dtmin == 0.0 and o.tape_steps == 0 and o.norm_mode == 0 and o.mlp_math == 0
    assert ctypes.sizeof(o) == 144
    # what the reference's diffeq structs request (pendulum.jl:11): ForwardDiffSensitivity(), Tsit5()
    assert o.sensealg == ldeq.SENSE_FORWARD_DUAL == 1 and o.solver == 0
    o2 = ldeq.default_opts(adaptive=False, dt=0.05, abstol=1e-8)
    assert (o2.adaptive, o2.dt, o2.abstol) == (0, 0.05, 1e-8)
    with pytest.raises(TypeError):
        ldeq.default_opts(no_such_option=1)


def test_version_and_clean_failure_without_a_device(ldeq):
    lib = ldeq._cabi.load()
    assert lib.ldeq_version() == 210      # round 2: + ldeq_opts_default_solver, ldeq_pattern_extractor_*, solver enum DP5 / BS3 / RK4
    import torch
    if not torch.cuda.is_available():
        # the product path must fail loudly, not fall back to the CPU
        with pytest.raises(ldeq.LdeqError):
            ldeq._cabi.Handle(0)
        with pytest.raises(RuntimeError):
            ldeq.goku_solve(torch.zeros(4, 2), torch.ones(4, 1), [0.0, 0.05, 0.1])
        with pytest.raises(RuntimeError):
            ldeq.mlp_solve(torch.zeros(4, 16), torch.zeros(46816), [16, 200, 200, 16], [0.0, 0.05])


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "latentdiffeq.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "libldeq_oracle" not in txt, f


def test_c_program_compiles_against_the_header_alone():
    """tests/cabi_smoke.c includes only include/ldeq.h + the CUDA runtime header; it must compile and link here (no GPU
    needed for that) -- the executable stand-in for the Julia `ccall` sequence."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("_entry", os.path.join(ROOT, "__graft_entry__.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    exe = m.build_cabi_smoke()
    assert os.path.exists(exe)
    src = open(os.path.join(ROOT, "tests", "cabi_smoke.c")).read()
    assert "#include <Python.h>" not in src and "#include <torch" not in src


@pytest.mark.gpu
def test_c_program_drives_the_abi():
    import subprocess
    exe = os.path.join(ROOT, "tests", "cabi_smoke")
    if not os.path.exists(exe):
        pytest.skip("tests/cabi_smoke not built (run __graft_entry__.build())")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "cabi_smoke ok" in r.stdout, r.stdout + r.stderr
