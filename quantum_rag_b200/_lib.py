"""Loader and in-tree builder for ``libqrag.so`` (the C ABI in ``include/qrag.h``).

The library is plain CUDA runtime code with C linkage: no torch types cross the
boundary, only device pointers, sizes and a stream handle.  There is no CPU
fallback -- if the library or a CUDA device is missing the calls raise.
"""
from __future__ import annotations

import ctypes
import os
import shutil
import subprocess
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_size_t, c_uint16, c_uint32, c_void_p
from typing import Dict, List, Optional

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC_DIR = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libqrag.so")
SOURCES = ["lib.cu", "probe.cu", "amp_fidelity.cu", "amp_stream.cu", "sv_kernels.cu", "sv_angle.cu", "sort_scores.cu", "fmap_warp.cu", "fmap_rerank.cu", "search_exact.cu", "search_tc.cu"]
# Kernel-tuning hooks (environment switches that skip or re-shape work inside product kernels) compile only with
# -DQRAG_TUNING, which is NOT in these flags: the shipped library reads no environment variable.
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]

# name -> (restype, argtypes); mirrors include/qrag.h
PROTOTYPES: Dict[str, tuple] = {
    "qrag_last_error": (c_char_p, []),
    "qrag_version": (c_int, []),
    "qrag_device_info": (c_int, [POINTER(c_int), POINTER(c_int), POINTER(c_int)]),
    "qrag_probe_fp64_fma_rate": (c_int, [POINTER(c_double), c_void_p, c_void_p]),
    "qrag_set_overlap": (c_int, [c_int]),
    "qrag_get_overlap": (c_int, []),
    "qrag_set_fmap_kernel": (c_int, [c_int]),
    "qrag_sv_fidelity_angle": (c_int, [c_void_p, c_int, c_void_p, c_int64, c_void_p, c_int64, c_int, c_int, c_int,
                                       c_void_p, c_void_p]),
    "qrag_amp_fidelity": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_int, c_int, c_int,
                                  c_void_p, c_void_p, c_void_p]),
    "qrag_amp_rerank": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_int, c_int, c_int,
                                c_void_p, c_void_p, c_void_p, c_void_p]),
    "qrag_amp_rerank_host_workspace": (c_int, [c_int, c_int64, c_int, c_int, POINTER(c_size_t)]),
    "qrag_amp_rerank_host": (c_int, [c_void_p, c_int, c_void_p, c_int64, c_void_p, c_int64, c_int, c_int, c_int,
                                     c_void_p, c_size_t, c_void_p, c_void_p, c_void_p]),
    "qrag_fmap_rerank_workspace": (c_int, [c_int, c_int64, c_int, POINTER(c_size_t)]),
    "qrag_fmap_filter_error_bound": (c_int, [c_int, POINTER(c_double)]),
    "qrag_fmap_filter_scores": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_int, c_int, c_int,
                                        c_void_p, c_void_p]),
    "qrag_fmap_rerank": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_int, c_int, c_int, c_int,
                                 c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "qrag_sort_scores_workspace": (c_int, [c_int, c_int64, POINTER(c_size_t)]),
    "qrag_sort_scores_stable": (c_int, [c_void_p, c_int, c_int64, c_int, c_int, c_void_p, c_void_p, c_void_p, c_size_t,
                                        c_void_p]),
    "qrag_mock_embedding": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_void_p]),
    "qrag_search_workspace": (c_int, [c_int, c_int64, c_int, c_int, POINTER(c_size_t)]),
    "qrag_search_topk": (c_int, [c_void_p, c_int, c_void_p, c_int64, c_int, c_int, c_int, c_int64, c_void_p,
                                 c_void_p, c_void_p, c_size_t, c_void_p]),
    "qrag_index_prepared_dims": (c_int, [c_int, c_int, POINTER(c_int)]),
    "qrag_index_prepare": (c_int, [c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "qrag_search_tc_workspace": (c_int, [c_int, c_int64, c_int, c_int, c_int, c_int, POINTER(c_size_t)]),
    "qrag_search_topk_tc": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, c_int,
                                    c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "qrag_search_tc_begin": (c_int, [c_void_p, c_int, c_void_p, c_int64, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                                     c_size_t, c_void_p]),
    "qrag_search_tc_filter": (c_int, [c_int, c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_void_p, c_int,
                                      c_void_p, c_void_p, c_size_t, c_void_p]),
    "qrag_search_tc_finish": (c_int, [c_void_p, c_int, c_void_p, c_int64, c_int, c_int, c_int, c_int64, c_void_p,
                                      c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "qrag_search_tc_scores": (c_int, [c_void_p, c_int, c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p, c_size_t,
                                      c_void_p]),
    "qrag_search_tc_exchange_len": (c_int, [c_int, c_int, POINTER(c_int)]),
    "qrag_search_tc_finish_packed": (c_int, [c_void_p, c_int, c_void_p, c_int64, c_int, c_int, c_int, c_int64, c_void_p,
                                             c_int, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    "qrag_owner_finalize": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p,
                                    c_void_p]),
    "qrag_topk_merge_workspace": (c_int, [c_int, c_int, c_int, c_int, POINTER(c_size_t)]),
    "qrag_topk_merge": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                                c_void_p, c_size_t, c_void_p]),
}

ERROR_NAMES = {-1: "QRAG_ERR_INVALID", -2: "QRAG_ERR_CUDA", -3: "QRAG_ERR_UNSUPPORTED", -4: "QRAG_ERR_WORKSPACE",
               -5: "QRAG_ERR_INEXACT"}


class QragError(RuntimeError):
    """Raised for every non-zero return code of libqrag (no silent fallback)."""

    def __init__(self, code: int, message: str):
        super().__init__(f"{ERROR_NAMES.get(code, code)}: {message}")
        self.code = code


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; cannot build libqrag.so")


def _stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC_DIR, f) for f in os.listdir(CSRC_DIR)] + [
        os.path.join(os.path.dirname(PKG_DIR), "include", "qrag.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, tuning: bool = False) -> str:
    """Compile every CUDA source for sm_100a into the in-tree ``libqrag.so``.

    ``tuning=True`` adds -DQRAG_TUNING (kernel-tuning switches read from the environment; tools/sweep_amp.py) -- never
    what ships: rebuild without it (``build(force=True)``) afterwards."""
    if not force and not tuning and not _stale():
        return LIB_PATH
    cmd = [_nvcc()] + NVCC_FLAGS + (["-DQRAG_TUNING"] if tuning else []) + ["-o", LIB_PATH] + [
        os.path.join(CSRC_DIR, s) for s in SOURCES]
    if verbose:
        print(" ".join(cmd))
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + proc.stdout + proc.stderr)
    return LIB_PATH


_LIB: Optional[ctypes.CDLL] = None


def load() -> ctypes.CDLL:
    """dlopen the in-tree library and attach the prototypes.  Raises if it is missing."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(libqrag has no CPU fallback)")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _LIB = lib
    return _LIB


def exported_symbols() -> List[str]:
    return list(PROTOTYPES)


def check(rc: int) -> None:
    if rc != 0:
        raise QragError(rc, load().qrag_last_error().decode("utf-8", "replace"))
