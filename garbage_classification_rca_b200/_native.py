"""ctypes binding of libmmrca.so (include/mmrca.h) and the in-tree nvcc build.

The product path has no CPU fallback: `lib()` raises if the shared library is missing and
every entry point fails on a non-sm_100 device (MMRCA_ERR_NO_DEVICE).
"""
from __future__ import annotations

import ctypes as C
import os
import shutil
import subprocess
import threading

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC_DIR = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libmmrca.so")
INCLUDE_DIR = os.path.join(os.path.dirname(PKG_DIR), "include")

# -cudart shared: the library uses the process's libcudart.so.12 (the one torch has loaded) instead of embedding a
# second, static CUDA runtime with its whole symbol table
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC", "-cudart", "shared", "-Xlinker", "-rpath=/usr/local/cuda/lib64"]
SOURCES = ["mmrca_capi.cu"]

# mmrca.h constants
FLAG_REVERSE, FLAG_FEATURES_ONLY, FLAG_CROSS_ATTENTION_ONLY = 1, 2, 4
FLAG_FEATURE_GRADS = 256
FLAG_TRAINING = 512
FLAG_FEATURES_BF16 = 1024
FLAG_ZERO_GRADS = 2048
COMPUTE_FP32, COMPUTE_BF16, COMPUTE_BF16_FUSED = 0, 1, 2
WS_TEXT_SA_IMAGE, WS_IMAGE_SA_IMAGE = 0, 1
QUERY_ABI_VERSION, QUERY_DEVICE_OK, QUERY_SM_COUNT, QUERY_KERNEL_LAUNCHES, QUERY_RESET_LAUNCHES, QUERY_HAS_BF16 = range(6)
ABI_VERSION = 4

_fp = C.c_void_p  # device pointers travel as integers


class AttnParams(C.Structure):
    _fields_ = [(n, _fp) for n in ("wq", "bq", "wk", "bk", "wv", "bv", "ln_g", "ln_b")]


class HeadParams(C.Structure):
    _fields_ = [("sa_img", AttnParams), ("sa_txt", AttnParams), ("ca1", AttnParams), ("ca2", AttnParams),
                ("wf", _fp), ("bf", _fp)]


class HeadDesc(C.Structure):
    _fields_ = [("batch", C.c_int32), ("d_img", C.c_int32), ("d_txt", C.c_int32), ("n_classes", C.c_int32),
                ("flags", C.c_uint32), ("compute", C.c_int32), ("drop_p", C.c_float), ("drop_seed", C.c_uint64)]


class KernelTime(C.Structure):
    _fields_ = [("name", C.c_char_p), ("ms", C.c_float)]


class HierParams(C.Structure):
    """MmrcaHierParams / MmrcaHierGrads (same layout)."""
    _fields_ = [(n, _fp) for n in ("w_img", "b_img", "w_txt", "b_txt", "w_all", "b_all")]


class HierDesc(C.Structure):
    _fields_ = [("batch", C.c_int32), ("n_classes", C.c_int32), ("drop_p", C.c_float), ("flags", C.c_uint32),
                ("drop_seed", C.c_uint64)]


HIER_FEATURE_GRADS = 1


class FusionParams(C.Structure):
    """MmrcaFusionParams / MmrcaFusionGrads (same layout)."""
    _fields_ = [(n, _fp) for n in ("w_img", "b_img", "w_txt", "b_txt", "w_cat", "b_cat", "w_fc", "b_fc")]


class FusionDesc(C.Structure):
    _fields_ = [("batch", C.c_int32), ("d_img", C.c_int32), ("d_txt", C.c_int32), ("hidden", C.c_int32),
                ("n_classes", C.c_int32), ("flags", C.c_uint32), ("drop_p", C.c_float), ("drop_seed", C.c_uint64)]


FUSION_NORMALIZED = 1
FUSION_BF16 = 2


class TokenDesc(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("batch", "seq_len", "d_in_q", "d_in_kv", "d_kq", "d_v", "reverse")] + \
               [("flags", C.c_uint32)]


TOKEN_WEIGHTS_READY = 1
TOKEN_TRAINING = 2
TOKEN_OUT_BF16 = 4


class CeDesc(C.Structure):
    _fields_ = [("class_weight", _fp), ("label_smoothing", C.c_float)]


EXPORTS = ("mmrca_query", "mmrca_last_error", "mmrca_head_workspace_bytes", "mmrca_head_forward",
           "mmrca_head_backward", "mmrca_cross_entropy", "mmrca_head_train_step", "mmrca_attention_forward",
           "mmrca_attention_backward_scratch_bytes", "mmrca_attention_backward", "mmrca_timing_begin",
           "mmrca_timing_end", "mmrca_dev_umma_selftest", "mmrca_attention_forward_scratch_bytes",
           "mmrca_head_workspace_offset", "mmrca_dropout_mask", "mmrca_dev_set_debug",
           "mmrca_hier_workspace_bytes", "mmrca_hier_forward", "mmrca_hier_backward", "mmrca_hier_train_step",
           "mmrca_hier_backward_features",
           "mmrca_peer_allreduce_mean", "mmrca_peer_allreduce_pad_bytes", "mmrca_peer_allreduce_status",
           "mmrca_feature_handoff", "mmrca_sgd_step", "mmrca_adamw_step",
           "mmrca_token_attention_workspace_bytes", "mmrca_token_attention_forward", "mmrca_token_attention_backward",
           "mmrca_fusion_workspace_bytes", "mmrca_fusion_forward", "mmrca_fusion_backward", "mmrca_fusion_train_step")


def _sources_newer_than_lib() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC_DIR, f) for f in os.listdir(CSRC_DIR)] + [os.path.join(INCLUDE_DIR, "mmrca.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/*.cu for sm_100a into libmmrca.so next to this file (nvcc cross-compiles
    without a GPU).  Returns the library path."""
    if not force and not _sources_newer_than_lib():
        return LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build libmmrca.so")
    tmp = LIB_PATH + ".tmp.%d" % os.getpid()
    cmd = [nvcc] + NVCC_FLAGS + ["-I", INCLUDE_DIR, "-o", tmp] + [os.path.join(CSRC_DIR, s) for s in SOURCES]
    if verbose:
        print(" ".join(cmd), flush=True)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    os.replace(tmp, LIB_PATH)
    return LIB_PATH


_lib = None
_lock = threading.Lock()


def lib() -> C.CDLL:
    """The loaded C-ABI library.  Raises (loudly) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU / PyTorch fallback for the MM-RCA head)")
        L = C.CDLL(LIB_PATH)
        L.mmrca_query.argtypes = [C.c_int]
        L.mmrca_query.restype = C.c_int
        L.mmrca_last_error.argtypes = []
        L.mmrca_last_error.restype = C.c_char_p
        L.mmrca_head_workspace_bytes.argtypes = [C.POINTER(HeadDesc), C.c_int]
        L.mmrca_head_workspace_bytes.restype = C.c_size_t
        L.mmrca_head_workspace_offset.argtypes = [C.POINTER(HeadDesc), C.c_int, C.c_int]
        L.mmrca_head_workspace_offset.restype = C.c_longlong
        L.mmrca_head_forward.argtypes = [C.POINTER(HeadDesc), C.POINTER(HeadParams), _fp, _fp, _fp, C.c_float,
                                         _fp, _fp, C.c_size_t, _fp]
        L.mmrca_head_forward.restype = C.c_int
        L.mmrca_head_backward.argtypes = [C.POINTER(HeadDesc), C.POINTER(HeadParams), _fp, _fp, _fp, C.c_float,
                                          _fp, C.POINTER(HeadParams), _fp, _fp, _fp, C.c_size_t, _fp]
        L.mmrca_head_backward.restype = C.c_int
        L.mmrca_cross_entropy.argtypes = [_fp, _fp, C.POINTER(CeDesc), C.c_int32, C.c_int32, _fp, _fp, _fp]
        L.mmrca_cross_entropy.restype = C.c_int
        L.mmrca_head_train_step.argtypes = [C.POINTER(HeadDesc), C.POINTER(HeadParams), _fp, _fp, _fp, C.c_float,
                                            _fp, C.POINTER(CeDesc), _fp, _fp, C.POINTER(HeadParams), _fp, _fp,
                                            _fp, C.c_size_t, _fp]
        L.mmrca_head_train_step.restype = C.c_int
        L.mmrca_attention_forward_scratch_bytes.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_int32]
        L.mmrca_attention_forward_scratch_bytes.restype = C.c_size_t
        L.mmrca_attention_forward.argtypes = [C.POINTER(AttnParams), _fp, _fp, C.c_int32, C.c_int32, C.c_int32,
                                              C.c_int32, C.c_int32, C.c_int32, _fp, _fp, _fp, C.c_size_t,
                                              C.c_int32, _fp]
        L.mmrca_attention_forward.restype = C.c_int
        L.mmrca_attention_backward_scratch_bytes.argtypes = [C.c_int32, C.c_int32, C.c_int32]
        L.mmrca_attention_backward_scratch_bytes.restype = C.c_size_t
        L.mmrca_attention_backward.argtypes = [C.POINTER(AttnParams), _fp, _fp, _fp, C.c_int32, C.c_int32,
                                               C.c_int32, C.c_int32, C.c_int32, C.POINTER(AttnParams), _fp, _fp,
                                               _fp, C.c_size_t, C.c_int32, _fp]
        L.mmrca_attention_backward.restype = C.c_int
        L.mmrca_timing_begin.argtypes = [C.c_int32]
        L.mmrca_timing_begin.restype = C.c_int
        L.mmrca_timing_end.argtypes = [C.POINTER(KernelTime), C.c_int32]
        L.mmrca_timing_end.restype = C.c_int
        L.mmrca_dev_umma_selftest.argtypes = [C.c_int32, _fp, _fp, _fp, C.c_int32, C.c_int32, _fp]
        L.mmrca_dev_umma_selftest.restype = C.c_int
        L.mmrca_dropout_mask.argtypes = [C.c_uint64, C.c_float, C.c_int32, C.c_int32, _fp, _fp]
        L.mmrca_dropout_mask.restype = C.c_int
        L.mmrca_dev_set_debug.argtypes = [_fp, C.c_int32]
        L.mmrca_dev_set_debug.restype = C.c_int
        L.mmrca_hier_workspace_bytes.argtypes = [C.POINTER(HierDesc)]
        L.mmrca_hier_workspace_bytes.restype = C.c_size_t
        L.mmrca_hier_forward.argtypes = [C.POINTER(HierDesc), C.POINTER(HierParams), C.POINTER(_fp), _fp, C.c_float,
                                         _fp, _fp, C.c_size_t, _fp]
        L.mmrca_hier_forward.restype = C.c_int
        L.mmrca_hier_backward.argtypes = [C.POINTER(HierDesc), C.POINTER(HierParams), _fp, C.POINTER(HierParams),
                                          _fp, C.c_size_t, _fp]
        L.mmrca_hier_backward.restype = C.c_int
        L.mmrca_hier_backward_features.argtypes = [C.POINTER(HierDesc), C.POINTER(HierParams), C.POINTER(_fp), _fp, C.c_float,
                                                   C.POINTER(_fp), _fp, C.c_size_t, _fp]
        L.mmrca_hier_backward_features.restype = C.c_int
        L.mmrca_hier_train_step.argtypes = [C.POINTER(HierDesc), C.POINTER(HierParams), C.POINTER(_fp), _fp,
                                            C.c_float, _fp, C.POINTER(CeDesc), _fp, _fp, C.POINTER(HierParams),
                                            _fp, C.c_size_t, _fp]
        L.mmrca_hier_train_step.restype = C.c_int
        L.mmrca_peer_allreduce_mean.argtypes = [_fp, C.c_int32, C.c_int32, C.POINTER(_fp), C.POINTER(_fp), C.c_int32,
                                                C.c_int32, C.c_uint32, _fp]
        L.mmrca_peer_allreduce_mean.restype = C.c_int
        L.mmrca_peer_allreduce_pad_bytes.argtypes = [C.c_int32]
        L.mmrca_peer_allreduce_pad_bytes.restype = C.c_int
        L.mmrca_peer_allreduce_status.argtypes = [_fp, C.c_int32, _fp]
        L.mmrca_peer_allreduce_status.restype = C.c_int
        L.mmrca_token_attention_workspace_bytes.argtypes = [C.c_void_p]
        L.mmrca_token_attention_workspace_bytes.restype = C.c_size_t
        L.mmrca_token_attention_forward.argtypes = [C.c_void_p, C.c_void_p, _fp, _fp, _fp, _fp, C.c_size_t, _fp]
        L.mmrca_token_attention_forward.restype = C.c_int
        L.mmrca_token_attention_backward.argtypes = [C.c_void_p, C.c_void_p, _fp, _fp, _fp, C.c_void_p, _fp, _fp, _fp,
                                                     C.c_size_t, _fp]
        L.mmrca_token_attention_backward.restype = C.c_int
        L.mmrca_fusion_workspace_bytes.argtypes = [C.c_void_p]
        L.mmrca_fusion_workspace_bytes.restype = C.c_size_t
        L.mmrca_fusion_forward.argtypes = [C.c_void_p, C.c_void_p, _fp, _fp, _fp, C.c_float, _fp, _fp, C.c_size_t, _fp]
        L.mmrca_fusion_forward.restype = C.c_int
        L.mmrca_fusion_backward.argtypes = [C.c_void_p, C.c_void_p, _fp, _fp, _fp, C.c_float, _fp, C.c_void_p, _fp, _fp,
                                            _fp, C.c_size_t, _fp]
        L.mmrca_fusion_backward.restype = C.c_int
        L.mmrca_fusion_train_step.argtypes = [C.c_void_p, C.c_void_p, _fp, _fp, _fp, C.c_float, _fp, C.c_void_p, _fp, _fp,
                                              C.c_void_p, _fp, _fp, _fp, C.c_size_t, _fp]
        L.mmrca_fusion_train_step.restype = C.c_int
        L.mmrca_feature_handoff.argtypes = [_fp, C.c_int32, C.c_int64, C.c_int32, _fp, C.c_int32, C.c_int32, C.c_int32,
                                            C.c_int32, C.c_int32, _fp, _fp, C.c_int32, _fp]
        L.mmrca_feature_handoff.restype = C.c_int
        L.mmrca_sgd_step.argtypes = [_fp, _fp, _fp, C.c_int64, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int32,
                                     C.c_int32, _fp]
        L.mmrca_sgd_step.restype = C.c_int
        L.mmrca_adamw_step.argtypes = [_fp, _fp, _fp, _fp, C.c_int64, C.c_float, C.c_float, C.c_float, C.c_float,
                                       C.c_float, C.c_int32, _fp]
        L.mmrca_adamw_step.restype = C.c_int
        if L.mmrca_query(QUERY_ABI_VERSION) != ABI_VERSION:
            raise RuntimeError("libmmrca.so ABI version mismatch: rebuild it")
        _lib = L
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().mmrca_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed (code {rc}): {msg}")


def last_error() -> str:
    return lib().mmrca_last_error().decode("utf-8", "replace")


def kernel_launches(reset: bool = False) -> int:
    """Kernels launched by libmmrca on the calling thread since the last reset."""
    return lib().mmrca_query(QUERY_RESET_LAUNCHES if reset else QUERY_KERNEL_LAUNCHES)


def timing_begin(max_records: int = 4096) -> None:
    check(-lib().mmrca_timing_begin(max_records), "mmrca_timing_begin")


def timing_end(max_records: int = 4096):
    """-> list of (kernel name, milliseconds) for every launch since timing_begin (blocks on the events)."""
    buf = (KernelTime * max_records)()
    n = lib().mmrca_timing_end(buf, max_records)
    if n < 0:
        check(-n, "mmrca_timing_end")
    return [(buf[i].name.decode(), float(buf[i].ms)) for i in range(min(n, max_records))]
