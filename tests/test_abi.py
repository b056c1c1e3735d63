"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol
include/coltt_b200.h declares, and refuses to compute without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "coltt_b200.h")).read()
    return sorted(set(re.findall(r"COLTT_API\s+[\w\s\*]+?\b(coltt_b200_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from coltt_b200 import _lib
    L = _lib.lib()
    syms = _header_symbols()
    assert len(syms) >= 20
    assert sorted(_lib.ABI_SYMBOLS) == syms, "python binding list and header disagree"
    for s in syms:
        assert hasattr(L, s), f"libcoltt_b200.so does not export {s}"


def test_signatures_have_no_torch_or_cuda_types():
    src = open(os.path.join(ROOT, "include", "coltt_b200.h")).read()
    assert "torch" not in src.lower().replace("no torch", "")
    assert "cudaStream_t stream" not in src  # streams cross as void*
    assert '#include <cuda' not in src


def test_no_cpu_fallback_without_gpu():
    import coltt_b200
    from coltt_b200 import _lib
    if _lib.lib().coltt_b200_device_count() > 0:
        pytest.skip("a B200 is visible")
    with pytest.raises(coltt_b200.ColttError) as e:
        coltt_b200.VectorSpace("c", coltt_b200.Metadata(8))
    assert e.value.code == -8 and "no CPU fallback" in e.value.message
    with pytest.raises(coltt_b200.ColttError) as e:          # page-locked buffers are a CUDA service too
        coltt_b200.pinned_empty((4, 8))
    assert e.value.code == -8


def test_product_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing under coltt_b200/ may import, link or call it."""
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "coltt_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", "Makefile")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                if re.search(r"\boracle\b|liboracle|orc_", txt):
                    bad.append(os.path.join(dirpath, f))
    assert not bad, bad


def test_score_helper_matches_oracle(oracle):
    import coltt_b200
    s = np.array([0.0, 0.25, 1.0, 2.0, 150.0], dtype=np.float32)
    for metric in (0, 1):
        got = coltt_b200.score_helper(s, metric)
        want = np.array([oracle.lib().orc_score_helper(float(x), metric) for x in s], dtype=np.float32)
        assert got.tobytes() == want.tobytes()


def test_certificate_margin_is_a_function_of_dim():
    """coltt_b200_fast_eps_rel is pure host arithmetic: 1.25 (dim 2^-23 + (dim/8 + 3) 2^-24) + 2^-20."""
    from coltt_b200 import _lib
    L = _lib.lib()
    for dim in (1, 64, 100, 768, 1536, 4096):
        want = 1.25 * (dim * 2.0 ** -23 + (dim / 8.0 + 3.0) * 2.0 ** -24) + 2.0 ** -20
        got = float(L.coltt_b200_fast_eps_rel(dim))
        assert abs(got - want) <= 1e-6 * want, (dim, got, want)
    assert L.coltt_b200_fast_eps_rel(64) < L.coltt_b200_fast_eps_rel(768) < L.coltt_b200_fast_eps_rel(1536)
    assert 1.0e-4 < L.coltt_b200_fast_eps_rel(768) < 1.5e-4


def test_multi_gpu_entry_points_fail_cleanly_without_a_gpu():
    """No device: the communicator calls return codes (COLTT_ERR_NO_DEVICE), never abort; shutdown with nothing open is a no-op."""
    from coltt_b200 import _lib
    L = _lib.lib()
    if L.coltt_b200_device_count() > 0:
        pytest.skip("a GPU is present")
    devs = (C.c_int * 1)(0)
    h = (C.c_void_p * 1)()
    assert L.coltt_b200_init(devs, 1, h) == -8
    assert b"no CPU fallback" in L.coltt_b200_last_error() or b"CUDA" in L.coltt_b200_last_error()
    blob = (C.c_uint8 * 128)()
    out = C.c_void_p()
    assert L.coltt_b200_comm_init_rank(blob, 0, 1, 0, C.byref(out)) == -8
    assert L.coltt_b200_comm_init_rank(blob, 3, 2, 0, C.byref(out)) == -1       # rank >= world: invalid before any device work
    L.coltt_b200_shutdown()
