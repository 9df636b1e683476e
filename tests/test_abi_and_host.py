"""CPU tests of the drop-in boundary: the C-ABI libraries load and export every symbol that
include/dzopt.h declares (no compute calls: there is no GPU here), and the host-side mirror of the
Julia API validates its arguments like the reference does."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "dzopt.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dzo_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_both_families():
    syms = declared_symbols()
    assert "dzo_bfgs_step" in syms and "dzo_cpu_bfgs_step" in syms and "dzo_gd_step" in syms
    assert len(syms) > 80


def test_cuda_library_exports_every_declared_symbol(dz):
    lib = C.CDLL(dz.lib_path)
    missing = [s for s in declared_symbols() if not s.startswith("dzo_cpu_") and not hasattr(lib, s)]
    assert not missing, f"libdzopt_b200.so lacks: {missing}"


def test_oracle_library_exports_every_declared_symbol(orc):
    lib = C.CDLL(orc.LIB_PATH)
    missing = [s for s in declared_symbols() if s.startswith("dzo_cpu_") and not hasattr(lib, s)]
    assert not missing, f"libdzo_oracle.so lacks: {missing}"


def test_product_library_does_not_link_the_oracle(dz):
    """The product path must not route through the oracle (or any CPU fallback)."""
    import subprocess
    out = subprocess.run(["ldd", dz.lib_path], capture_output=True, text=True).stdout
    assert "dzo_oracle" not in out
    nm = subprocess.run(["nm", "-D", "--defined-only", dz.lib_path], capture_output=True, text=True).stdout
    assert "dzo_cpu_" not in nm
    pkg_dir = os.path.dirname(dz.lib_path)
    for root, _, files in os.walk(os.path.dirname(pkg_dir)):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".jl")):
                text = open(os.path.join(root, f)).read()
                assert "import oracle" not in text and "libdzo_oracle" not in text, f


def test_no_gpu_means_loud_failure_not_fallback(dz):
    """On this CPU-only box the constructor must raise DZO_ERR_NO_DEVICE (-6), never compute."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    EF = dz.ExampleFunctions
    with pytest.raises(dz.DZOptError) as e:
        dz.BFGSOptimizer(EF.rosenbrock_function, EF.rosenbrock_gradient_, np.zeros(2), 1.0)
    assert e.value.code == -6
    with pytest.raises(dz.DZOptError) as e:
        dz.GradientDescentOptimizer(EF.rosenbrock_function, EF.rosenbrock_gradient_, dz.QuadraticLineSearch(0),
                                    np.zeros(64), 1.0)
    assert e.value.code == -6


def test_constructor_argument_validation(dz):
    EF = dz.ExampleFunctions
    with pytest.raises(TypeError):                                   # host closures cannot run in the CUDA step!
        dz.BFGSOptimizer(lambda x: 0.0, EF.rosenbrock_gradient_, np.zeros(2), 1.0)
    with pytest.raises(TypeError):                                   # objective / gradient of different functions
        dz.BFGSOptimizer(EF.rosenbrock_function, EF.riesz_gradient_, np.zeros(2), 1.0)
    with pytest.raises(TypeError):
        dz.BFGSOptimizer(EF.rosenbrock_function, EF.rosenbrock_gradient_, np.zeros(2))          # arity (:753-766)
    with pytest.raises(TypeError):
        dz.GradientDescentOptimizer(EF.riesz_energy, EF.riesz_gradient_, "not a line search", np.zeros((4, 3)), 1.0)
    with pytest.raises(ValueError):
        dz.BFGSOptimizer(EF.rosenbrock_function, EF.rosenbrock_gradient_, np.zeros((3, 2)), 1.0)   # needs batched=True
    assert dz.StepType.NullStep == 0 and dz.StepType.GradientDescentStep == 1 and dz.StepType.BFGSStep == 2
    assert dz.QuadraticLineSearch().max_increases == 0


def test_layout_matches_julia_column_major(dz):
    a, n, batch, pshape, dim = dz._layout(np.zeros((5, 3)), dz.OBJ_RIESZ, False)   # Julia 3 x 5 Matrix
    assert (n, batch, pshape, dim) == (15, 1, (5, 3), 3)
    a, n, batch, pshape, dim = dz._layout(np.zeros((7, 16)), dz.OBJ_ROSENBROCK, True)
    assert (n, batch, pshape, dim) == (16, 7, (16,), 0)


def test_error_paths_need_no_gpu(dz):
    """argument validation happens before any CUDA call: null handles / pointers and bad shapes give
    DZO_ERR_INVALID_ARGUMENT (-1) with a message, on any machine"""
    lib = dz.lib()
    assert lib.dzo_bfgs_step(None, 1) == -1 and b"bad arguments" in lib.dzo_last_error()
    assert lib.dzo_gd_step(None, 1) == -1
    assert lib.dzo_lbfgs_step(None, 1) == -1
    assert lib.dzo_adgd_step(None, 1) == -1
    out = np.zeros(4)
    dp = out.ctypes.data_as(dz._capi.c_double_p)
    assert lib.dzo_bfgs_get_point(None, dp) == -1
    h = C.c_void_p()
    assert lib.dzo_bfgs_create(C.byref(h), 1, 0, 0, 3, 1, dp, 1.0, 0) == -1          # odd n for Rosenbrock
    assert b"even n" in lib.dzo_last_error()
    assert lib.dzo_bfgs_create(C.byref(h), 7, 0, 0, 4, 1, dp, 1.0, 0) == -1          # unknown objective id
    assert lib.dzo_bfgs_create(C.byref(h), 2, 0, 5, 4, 1, dp, 1.0, 0) == -1          # Riesz: n not a multiple of dim
    assert lib.dzo_lbfgs_create(C.byref(h), 1, 0, 0, 4, dp, 1.0, 0, 0) == -1         # history_length < 1
    assert lib.dzo_set_tuning(b"no_such_knob", 1) == -1
    assert lib.dzo_dot(0, 0, dp, dp, dp, 0) == -1                                     # n <= 0
    e = C.c_double()
    assert lib.dzo_pairwise_energy(9, 0, 4, dp, dp, dp, None, C.byref(e), 0) == -1    # unknown potential


def test_bench_reference_arm_contract():
    """bench.py --impl reference runs without a GPU and prints ONE JSON line with the contract's keys."""
    import json
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "3", "--warmup", "3"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["cpu_baseline"]["kind"] == "port" and d["value"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"]


def test_product_pcg_generator_equals_the_oracle_and_the_golden_vectors(dz, orc):
    """dzoptimization.jl_b200/pcg.py (legacy/PCG.jl:7-22, numpy): the inputs of bench.py come from the product, not
    from oracle/; bit-identical to dzo_cpu_pcg_fill and to the committed golden vectors."""
    import json
    import os
    here = os.path.dirname(os.path.abspath(__file__))
    with open(os.path.join(here, "golden", "golden.json")) as f:
        golden = json.load(f)["pcg"]
    for seed, vals in golden.items():
        assert [float(v).hex() for v in dz.pcg_fill(8, int(seed))] == vals
    for count, seed in ((1, 0), (4095, 3), (4096, 2024), (4097, 5), (10000, 2 ** 63 + 5), (300000, 9)):
        assert np.array_equal(dz.pcg_fill(count, seed), orc.pcg_fill(count, seed))
    assert dz.pcg_fill(0, 1).size == 0
