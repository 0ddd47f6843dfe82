"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/*.h declares, and fails loudly (no CPU fallback) when there is no CUDA device."""
import ctypes as C
import os

import pytest

from conftest import load_pkg, ROOT


def test_library_exports_every_declared_symbol():
    pkg = load_pkg()
    lib = pkg.load_library()
    names = pkg.exported_symbols()
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert lib.moai_version() >= 100


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    pkg = load_pkg()
    lib = pkg.load_library()
    primes = (C.c_uint64 * 2)(1099511480321, 1099511390209)
    h = C.c_void_p()
    rc = lib.moai_context_create(C.c_int32(12), primes, C.c_int32(2), C.c_int32(0), C.byref(h))
    assert rc == 3  # MOAI_CUDA_ERROR
    assert b"no CPU fallback" in lib.moai_last_error()
    with pytest.raises(RuntimeError):
        pkg.Backend(12, [1099511480321, 1099511390209])


def test_product_never_imports_oracle():
    """The product package and its CUDA sources must not reference oracle/ (③)."""
    # the package, the C-ABI headers and the header-only C++ facade (include/moai_b200_seal.hpp, include/facade/)
    for top in (os.path.join(ROOT, "moai-fhe-transformerinference-public_b200"), os.path.join(ROOT, "include")):
        for dirpath, _, files in os.walk(top):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h", ".cpp")):
                    text = open(os.path.join(dirpath, f)).read()
                    assert "import oracle" not in text and "from oracle" not in text, f
                    assert "ckks_oracle" not in text and "libsealref" not in text, f
                    assert "moai_b200_mock" not in text and "mock_cabi" not in text and "orc_" not in text, f


def test_attention_rotation_steps_cover_the_pipeline():
    """Host-side helper: the steps one head takes in fast mode (csrc/modules.cu) at the repo's packing
    (num_batch = 256, 128 tokens, 32768 slots)."""
    import importlib
    pkg = importlib.import_module("moai-fhe-transformerinference-public_b200")
    st = pkg.attention_rotation_steps(256)
    slots = 128 * 256
    # QK^T: hoisted baby steps b * 256 (b < 16), +-16 a * 256 for the outer part (a < 8)
    assert set(b * 256 for b in range(1, 16)) <= set(st["qk"])
    for a in range(1, 8):
        assert 16 * a * 256 in st["qk"] and (slots - 16 * a * 256) in st["qk"]
    # softmax * V (Ct_ct_matrix_mul.hpp:70-151): g = 12 baby steps of V, 10 group rotations, 10 giant steps
    assert set(k * 256 for k in range(1, 12)) <= set(st["sv"])
    assert set((128 - 12 * i) * 256 for i in range(1, 11)) <= set(st["sv"])
    assert set(12 * j * 256 for j in range(1, 11)) <= set(st["sv"])
    assert all(0 < s < slots for s in st["qk"] + st["sv"])
