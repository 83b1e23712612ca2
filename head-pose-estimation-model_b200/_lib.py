"""ctypes binding of libhpose.so (C ABI in include/hpose.h).

This is the thin layer the north-star asks for: Python host code -> ctypes -> hand-written
sm_100a CUDA.  There is deliberately no fallback: if the shared library is missing or no B200 is
visible, every entry point raises ``HposeError``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
# experiment builds: HPOSE_LIB_SUFFIX=_x selects libhpose_x.so (built with the extra nvcc flags in HPOSE_NVCC_EXTRA)
_SUFFIX = os.environ.get("HPOSE_LIB_SUFFIX", "")
LIB_PATH = os.path.join(PKG_DIR, f"libhpose{_SUFFIX}.so")
SOURCES = ["api.cu", "backbone.cu", "blocks_tma.cu", "blocks_tc.cu", "blocks_chain.cu", "preproc.cu", "detect.cu", "stem_tc.cu", "dense_tc.cu", "heads.cu", "postproc.cu", "comm.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC"]

HP_BACKBONE_PARAMS = 101390
HP_MAX_FACES = 100
HP_KEYPOINTS = 6
HP_IMPL_FAST, HP_IMPL_NAIVE, HP_IMPL_CPASYNC, HP_IMPL_TMA = 0, 1, 2, 3
HP_RESULT_HEADER_INTS = 4
HP_DETECT_GRAPH = 1
HP_TRAIN_GRAPH = 1
(HP_OP_DENSE, HP_OP_ACT, HP_OP_ADD, HP_OP_MULCH, HP_OP_GAP, HP_OP_DROPOUT, HP_OP_LAYERNORM,
 HP_OP_MHA) = range(1, 9)
HP_ACT = {"linear": 0, None: 0, "relu": 1, "tanh": 2, "sigmoid": 3, "softsign": 4, "elu": 5, "selu": 6, "softplus": 7, "swish": 8,
          "leaky_relu": 9}
HP_OPT = {"sgd": 0, "adam": 1, "adamax": 2}


class HposeError(RuntimeError):
    """Raised for every failure reported by libhpose (code + hp_last_error())."""

    def __init__(self, code: int, msg: str):
        super().__init__(f"libhpose error {code}: {msg}")
        self.code = code


class hp_head_op(C.Structure):
    _fields_ = [("op", C.c_int32), ("in0", C.c_int32), ("in1", C.c_int32), ("out", C.c_int32),
                ("cin", C.c_int32), ("cout", C.c_int32), ("act", C.c_int32),
                ("w_off", C.c_int32), ("b_off", C.c_int32),
                ("heads", C.c_int32), ("key_dim", C.c_int32), ("op_id", C.c_int32),
                ("fparam", C.c_float), ("l2_w", C.c_float), ("l2_b", C.c_float)]


class hp_head_reg(C.Structure):
    _fields_ = [("channels", C.c_int32), ("per_image", C.c_int32)]


class hp_opt_config(C.Structure):
    _fields_ = [("kind", C.c_int32), ("lr", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float),
                ("eps", C.c_float)]


HEADERS = [os.path.join(CSRC, "common.cuh"), os.path.join(CSRC, "tc_common.cuh"),
           os.path.join(PKG_DIR, "..", "include", "hpose.h")]
OBJ_DIR = os.path.join(CSRC, "build" + _SUFFIX)


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def needs_build() -> bool:
    return _stale(LIB_PATH, [os.path.join(CSRC, s) for s in SOURCES] + HEADERS)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile libhpose.so in-tree for sm_100a (nvcc cross-compiles without a GPU): one object per translation
    unit (only stale ones are recompiled, in parallel), then one link."""
    if not force and not needs_build():
        return LIB_PATH
    from concurrent.futures import ThreadPoolExecutor
    os.makedirs(OBJ_DIR, exist_ok=True)
    flags = [f for f in NVCC_FLAGS if f != "-shared"] + os.environ.get("HPOSE_NVCC_EXTRA", "").split()

    def compile_one(src):
        obj = os.path.join(OBJ_DIR, src.replace(".cu", ".o"))
        if force or _stale(obj, [os.path.join(CSRC, src)] + HEADERS):
            cmd = ["nvcc"] + flags + ["-c", src, "-o", obj]
            if verbose:
                print(" ".join(cmd), file=sys.stderr)
            res = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
            if res.returncode != 0:
                raise RuntimeError(f"nvcc failed on {src}:\n{res.stdout}\n{res.stderr}")
        return obj

    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 4)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    cmd = ["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB_PATH] + objs + ["-ldl"]
    if verbose:
        print(" ".join(cmd), file=sys.stderr)
    res = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    return LIB_PATH


_PROTOS = {
    "hp_last_error": (C.c_char_p, []),
    "hp_version": (C.c_int, []),
    "hp_build_features": (C.c_int, []),
    "hp_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "hp_destroy": (C.c_int, [C.c_void_p]),
    "hp_set_impl": (C.c_int, [C.c_void_p, C.c_int]),
    "hp_launch_count": (C.c_int64, [C.c_void_p]),
    "hp_backbone_load_weights": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]),
    "hp_backbone_generation": (C.c_longlong, [C.c_void_p]),
    "hp_num_anchors": (C.c_int, [C.c_int, C.c_int]),
    "hp_backbone_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_void_p]),
    "hp_backbone_status": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "hp_backbone_read_activation": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                              C.c_void_p, C.c_size_t, C.c_void_p]),
    "hp_preprocess_u8": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "hp_preprocess_resize_u8": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "hp_head_create": (C.c_int, [C.c_void_p, C.POINTER(hp_head_op), C.c_int, C.POINTER(hp_head_reg), C.c_int,
                                 C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "hp_head_destroy": (C.c_int, [C.c_void_p, C.c_void_p]),
    "hp_head_set_weights": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]),
    "hp_head_get_weights": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]),
    "hp_head_get_grads": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]),
    "hp_head_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                  C.c_void_p]),
    "hp_head_train_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                     C.c_int, C.POINTER(hp_opt_config), C.c_uint64, C.c_void_p, C.c_void_p]),
    "hp_head_train_run": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_longlong, C.c_int,
                                    C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(hp_opt_config), C.c_uint64, C.c_int,
                                    C.c_void_p, C.c_void_p]),
    "hp_head_evaluate": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                   C.c_void_p, C.c_void_p]),
    "hp_dropout_hash": (C.c_uint32, [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]),
    "hp_decode_nms": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                C.c_float, C.c_float, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                C.c_void_p, C.c_void_p, C.c_void_p]),
    "hp_filter_detections": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_void_p]),
    "hp_extract_detections": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                        C.c_void_p, C.c_void_p]),
    "hp_nms": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_int, C.c_void_p, C.c_void_p,
                         C.c_void_p]),
    "hp_unified_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                     C.c_float, C.c_float, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "hp_detect_result_faces_offset": (C.c_size_t, [C.c_int]),
    "hp_detect_result_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "hp_detect_frames": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                   C.c_float, C.c_float, C.c_int, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]),
    "hp_comm_unique_id": (C.c_int, [C.c_void_p]),
    "hp_comm_init": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "hp_comm_destroy": (C.c_int, [C.c_void_p]),
    "hp_p2p_alloc": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "hp_p2p_open": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "hp_p2p_close": (C.c_int, [C.c_void_p]),
    "hp_debug_set_p2p": (C.c_int, [C.c_void_p, C.c_int]),
    "hp_fma_peak": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_double)]),
    "hp_debug_set_tile": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "hp_debug_tile_report": (C.c_int, [C.c_void_p, C.c_void_p]),
    "hp_debug_set_stem_tc": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]),
    "hp_debug_set_chain": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int]),
    "hp_debug_chain_status": (C.c_int, [C.c_void_p, C.c_void_p]),
    "hp_debug_chain_describe": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "hp_debug_tc_trace": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "hp_debug_dense": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                 C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "hp_debug_set_tc": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "hp_backbone_profile": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
}

EXPORTED_SYMBOLS = sorted(_PROTOS)
_lib = None


def lib() -> C.CDLL:
    """Load (never build implicitly on the GPU path) libhpose.so; fail loudly when absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise HposeError(-2, f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                                 "(there is no CPU or eager fallback)")
        l = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        for name, (res, args) in _PROTOS.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(code: int):
    if code != 0:
        raise HposeError(code, lib().hp_last_error().decode("utf-8", "replace"))
