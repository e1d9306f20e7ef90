"""The C-ABI library loads and exports every symbol include/ldpcb200.h declares; without a GPU
it refuses to work instead of falling back to anything."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    text = open(os.path.join(ROOT, "include", "ldpcb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ldpcb200_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree(pkg):
    assert header_functions() == sorted(pkg._lib.SYMBOLS)


def test_library_exports_every_symbol(pkg):
    lib = ctypes.CDLL(pkg._lib.LIB_PATH)
    for name in header_functions():
        assert hasattr(lib, name), name
    assert pkg._lib.load().ldpcb200_version() >= 100


def test_product_never_imports_the_oracle():
    pkg_dir = os.path.join(ROOT, "ldpcdecoders.jl_b200")
    for dirpath, _, files in os.walk(pkg_dir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle", src, flags=re.M), f
                assert not re.search(r"#\s*include[^\n]*oracle", src), f
                assert "libbporacle" not in src and "bp_oracle_batch" not in src, f


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only behaviour")
def test_no_gpu_means_loud_failure(pkg, codes):
    H = codes.gross_x()
    with pytest.raises(pkg._lib.LibraryError) as ei:
        pkg.BeliefPropagationDecoder(H, 0.01, 32)
    assert ei.value.code == pkg._lib.ENODEVICE
    assert "no CPU fallback" in str(ei.value)


def test_constructor_argument_types(pkg, codes):
    H = codes.gross_x()
    with pytest.raises(TypeError):
        pkg.BeliefPropagationDecoder(H, 1, 32)          # per::Float64
    with pytest.raises(TypeError):
        pkg.BeliefPropagationDecoder(H, 0.01, 32.0)     # max_iters::Int


def test_bad_arguments_are_reported_not_crashed(pkg):
    lib = pkg._lib.load()
    h = ctypes.c_void_p()
    colptr = np.array([0, 1], dtype=np.int64)
    rc = lib.ldpcb200_create(1, 1, colptr.ctypes.data, None, 0, 0.1, 5, 0, None, 0, ctypes.byref(h))
    assert rc != 0 and lib.ldpcb200_last_error()
    rc = lib.ldpcb200_create(1, 1, colptr.ctypes.data, colptr.ctypes.data, 7, 0.1, 5, 0, None, 0, ctypes.byref(h))
    assert rc == pkg._lib.EINVAL
    assert lib.ldpcb200_destroy(None) == 0


def test_limits_are_reported(pkg):
    """Degree / edge-count limits give LDPCB200_EUNSUPPORTED or the no-device error, never a crash
    (graph validation happens before any CUDA call only when a device exists, so on CPU the
    no-device error wins)."""
    lib = pkg._lib.load()
    h = ctypes.c_void_p()
    n = 200
    colptr = np.arange(n + 1, dtype=np.int64) * 1      # degree-1 columns, all in check 0 -> check degree 200 > 128
    rowval = np.zeros(n, dtype=np.int64)
    rc = lib.ldpcb200_create(1, n, colptr.ctypes.data, rowval.ctypes.data, 0, 0.1, 5, 0, None, 0, ctypes.byref(h))
    assert rc in (pkg._lib.EUNSUPPORTED, pkg._lib.ENODEVICE)
    assert lib.ldpcb200_last_error()


def test_reference_arm_of_the_bench_emits_the_contract_keys():
    """`bench.py --impl reference` (the CPU arm the driver times next to ours) runs without a GPU and prints one JSON line
    with the contract's keys, the same metric / unit / config as our arm, and a cpu_baseline describing the run."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in line, k
    assert line["impl"] == "reference" and line["unit"] == "syndromes/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert "C3" in line["config"]["workload"]


def test_documented_options_exist_in_the_library_source():
    """Every option name the header documents for ldpcb200_set_option is handled by the library, and every option the
    library handles is documented (guards the header against drifting from csrc/ldpcb200.cu)."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "include", "ldpcb200.h")).read()
    src = open(os.path.join(root, "ldpcdecoders.jl_b200", "csrc", "ldpcb200.cu")).read()
    a = hdr.index("Tunables, set before the first decode")
    b = hdr.index("int ldpcb200_set_option")
    documented = set(re.findall(r'"([a-z_0-9]+)"', hdr[a:b]))
    body = src[src.index("int ldpcb200_set_option("):]
    body = body[:body.index("\nint ldpcb200_info(")]
    handled = set(re.findall(r'k == "([a-z_0-9]+)"', body))
    assert handled, "option parser not found"
    assert documented - handled == set(), "documented but not handled: %s" % sorted(documented - handled)
    undocumented = handled - documented - {"slots"}            # ("slots": accepted for compatibility, unused)
    assert undocumented == set(), "handled but not documented in the header: %s" % sorted(undocumented)
