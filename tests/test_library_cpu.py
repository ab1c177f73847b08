"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol include/vgqa_b200.h declares.
No compute call is made (there is no GPU here) — the product has no CPU fallback, which is also checked."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib_path():
    from vgqa_b200 import build
    return build.build()


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "vgqa_b200.h")).read()
    return sorted(set(re.findall(r"\b(vgqa_[a-z0-9_]+)\s*\(", hdr)))


def test_header_declares_the_expected_entry_points():
    syms = declared_symbols()
    for s in ("vgqa_create", "vgqa_destroy", "vgqa_set_weight", "vgqa_finalize_weights", "vgqa_forward",
              "vgqa_forward_host", "vgqa_last_error", "vgqa_gemm_bf16", "vgqa_mha32", "vgqa_xattn1"):
        assert s in syms


def test_library_loads_and_exports_every_declared_symbol(lib_path):
    L = ctypes.CDLL(lib_path)
    for s in declared_symbols():
        assert hasattr(L, s), f"{s} is declared in include/vgqa_b200.h but not exported"


def test_library_contains_sm100a_tcgen05_and_tma_code(lib_path):
    sass = subprocess.run(["cuobjdump", "-sass", lib_path], capture_output=True, text=True)
    if sass.returncode != 0:
        pytest.skip("cuobjdump not available")
    assert "sm_100a" in sass.stdout
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM", "HMMA"):
        assert mnemonic in sass.stdout, mnemonic


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from oracle import vgqa_oracle as O
    from vgqa_b200.engine import GroundingEngine
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        GroundingEngine(O.synth_state_dict(0))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "vgqa_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in src.lower() or f == "__init__.py" and "oracle" not in src, f"{f} mentions the oracle"


def test_reference_flops_closed_form(lib_path):
    L = ctypes.CDLL(lib_path)
    L.vgqa_reference_flops.restype = ctypes.c_double
    L.vgqa_reference_flops.argtypes = [ctypes.c_int] * 8
    # SURVEY.md §8d table (FlopCounterMode over the reference modules), GFLOP per clip
    for (T, H, W, Lt, ref) in [(32, 7, 7, 20, 86.10), (64, 7, 7, 20, 172.23), (64, 14, 14, 20, 623.54),
                               (256, 7, 7, 20, 690.07), (128, 12, 12, 64, 1068.38)]:
        got = L.vgqa_reference_flops(T, H, W, Lt, 6, 6, 2048, 2) / 1e9
        assert abs(got - ref) / ref < 2e-3, (got, ref)
