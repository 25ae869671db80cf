"""The C-ABI library loads and exports exactly what include/qrag.h declares (no GPU needed)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "qrag.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(qrag_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(libqrag):
    from quantum_rag_b200 import _lib
    declared = _declared()
    assert len(declared) >= 14
    assert sorted(_lib.exported_symbols()) == declared           # python prototypes mirror the header
    for name in declared:
        assert hasattr(libqrag, name), name


def test_only_c_linkage_no_torch_dependency():
    import subprocess
    from quantum_rag_b200 import _lib
    out = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "torch" not in out and "c10" not in out
    nm = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = {line.split()[-1] for line in nm.splitlines() if " T " in line}
    assert set(_declared()) <= exported


def test_version_and_error_text(libqrag):
    assert libqrag.qrag_version() == 100
    nbytes = ctypes.c_size_t(0)
    assert libqrag.qrag_search_workspace(4, 1000, 64, 10, ctypes.byref(nbytes)) == 0 and nbytes.value > 0
    assert libqrag.qrag_search_workspace(4, 1000, 64, 5000, ctypes.byref(nbytes)) == -3
    assert b"k" in libqrag.qrag_last_error()
    assert libqrag.qrag_search_workspace(4, 1000, 64, 10, None) == -1


def test_argument_validation_happens_before_any_cuda_call(libqrag):
    # null pointers are rejected on the host side, so this is safe without a device
    assert libqrag.qrag_amp_fidelity(None, 1, None, None, 0, None, 1, 4, 2, 0, None, None, None) == -1
    assert libqrag.qrag_sv_fidelity_angle(None, 1, None, 1, None, 1, 4, 4, 1, None, None) == -1
    assert libqrag.qrag_topk_merge(None, None, 1, 1, 1, 1, 0, None, None, None, 0, None) == -1


def test_fails_loudly_without_device(libqrag):
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    buf = (ctypes.c_double * 64)()
    rc = libqrag.qrag_sv_fidelity_angle(buf, 1, buf, 1, None, 1, 8, 4, 1, buf, None)
    assert rc == -2 and b"no CPU fallback" in libqrag.qrag_last_error()
    sm = ctypes.c_int(0)
    assert libqrag.qrag_device_info(ctypes.byref(sm), None, None) == -2


def test_tensor_core_search_argument_validation(libqrag):
    """Bad arguments of the tcgen05 entry points are rejected on the host, before any CUDA work."""
    import ctypes
    kp = ctypes.c_int(0)
    assert libqrag.qrag_index_prepared_dims(384, 1, ctypes.byref(kp)) == 0 and kp.value == 400   # L2: D + 2 -> mult of 16
    assert libqrag.qrag_index_prepared_dims(384, 2, ctypes.byref(kp)) == 0 and kp.value == 384
    assert libqrag.qrag_index_prepared_dims(10, 0, ctypes.byref(kp)) == 0 and kp.value == 16
    assert libqrag.qrag_index_prepared_dims(384, 9, ctypes.byref(kp)) == -1
    assert libqrag.qrag_set_overlap(5) == -1 and libqrag.qrag_get_overlap() in (0, 1, 2)
    assert libqrag.qrag_set_fmap_kernel(7) == -1 and libqrag.qrag_set_fmap_kernel(0) == 0
    nbytes = ctypes.c_size_t(0)
    rc = libqrag.qrag_search_tc_workspace(4, 1000, 64, 5000, 0, 1, ctypes.byref(nbytes))
    assert rc in (-3, -2)                                  # k too large (or no device: still an error, never a fallback)
    rc = libqrag.qrag_search_topk_tc(None, 1, None, None, None, 10, 8, 1, 0, 0, None, None, None, None, 0, None)
    assert rc == -1 and b"null" in libqrag.qrag_last_error()
