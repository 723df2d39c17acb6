"""ctypes binding of libeslam_b200.so (include/eslam_b200.h).

There is NO fallback: if the library is missing or a call fails, a RuntimeError is raised.
The library is built in-tree by `__graft_entry__.build()` / `myslam_b200.build.build_library()`.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ESLAM_B200_LIB") or os.path.join(_HERE, "libeslam_b200.so")  # override: kernel experiments

N_PLANES = 12
DEC_FLOATS = 2700
DEC_BETA = 2696
N_COUNTERS = 8
MAX_COMPACT_BLOCKS = 4096
COUNTER_WORDS = N_COUNTERS + 2 * MAX_COMPACT_BLOCKS  # allocation size of a counters buffer (include/eslam_b200.h)
N_LOSS = 8
MAX_SAMPLES = 64
ABI_VERSION = 1


class Plane(C.Structure):
    _fields_ = [("offset", C.c_int64), ("H", C.c_int32), ("W", C.c_int32)]


class FieldDesc(C.Structure):
    _fields_ = [("plane", Plane * N_PLANES), ("dec_offset", C.c_int64), ("n_floats", C.c_int64),
                ("bound", (C.c_float * 2) * 3)]


class Camera(C.Structure):
    _fields_ = [("H", C.c_int32), ("W", C.c_int32), ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float),
                ("cy", C.c_float), ("H0", C.c_int32), ("H1", C.c_int32), ("W0", C.c_int32), ("W1", C.c_int32)]


class RenderCfg(C.Structure):
    _fields_ = [("n_stratified", C.c_int32), ("n_importance", C.c_int32), ("truncation", C.c_double),
                ("w_fs", C.c_double), ("w_center", C.c_double), ("w_tail", C.c_double), ("w_depth", C.c_double),
                ("w_color", C.c_double)]


MAX_PEERS = 8
EXCH_CTAS = 592


class Peers(C.Structure):
    _fields_ = [("rank", C.c_int32), ("world", C.c_int32), ("epoch", C.c_uint32), ("pad_", C.c_uint32),
                ("flags", C.c_void_p * MAX_PEERS), ("status", C.c_void_p), ("local_sync", C.c_void_p),
                ("adam_seq", C.c_uint64)]


_P = C.c_void_p
_I = C.c_int
_L = C.c_int64
_D = C.c_double
_FP = C.POINTER(FieldDesc)
_CP = C.POINTER(Camera)
_RP = C.POINTER(RenderCfg)
_PP = C.POINTER(Plane)

# name -> argtypes, exactly include/eslam_b200.h
PROTOTYPES = {
    "eslam_plane_import": [_P, _P, _PP, _P],
    "eslam_plane_export": [_P, _P, _PP, _P],
    "eslam_bind_decoders": [_P, _P],
    "eslam_decode_points": [_FP, _P, _P, _L, _P, _I, _P],
    "eslam_decode_backward": [_FP, _P, _P, _L, _P, _P, _P, _P],
    "eslam_sample_plane_feature": [_FP, _P, _P, _L, _I, _P, _P],
    "eslam_grid_sdf": [_FP, _P, _P, _P, _P, _I, _I, _I, _L, _L, _P, _P],
    "eslam_grid_sdf_hull": [_FP, _P, _P, _P, _P, _I, _I, _I, _L, _L, _P, _I, _P, _P],
    "eslam_grid_features": [_FP, _P, _P, _P, _P, _I, _I, _I, _P, _P, _P, _P],
    "eslam_grid_sdf_separable": [_FP, _P, _P, _P, _P, _I, _I, _I, _L, _L, _P, _P, _P, _P, _I, _P, _P],
    "eslam_grid_preact": [_FP, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P, _P, _P, _P],
    "eslam_grid_sdf_factored": [_FP, _P, _P, _P, _I, _I, _I, _L, _L, _P, _P, _P, _P, _I, _P, _P],
    "eslam_grid_sdf_rows": [_FP, _P, _P, _P, _I, _I, _I, _I, _I, _L, _P, _P, _P, _P, _I, _P, _P],
    "eslam_q_build": [_FP, _P, _P, _P],
    "eslam_q_adam_planes": [_FP, _P, _P, _P, _P, _P, _P, _D, _D, _I, _D, _D, _D, _P],
    "eslam_render_forward_q": [_FP, _P, _P, _P, _P, _I, _I, _P, _P, _P, _P, _P, _P, _P],
    "eslam_sample_rays": [_FP, _CP, _RP, _P, _I, _I, _P, _P, _I, _P, _P, _P, _P, _P, _I, _P, _P, _P, _P, _P, _P, _P,
                          _P, _P, _P, _P, _P],
    "eslam_sample_rays_frames": [_FP, _CP, _RP, _P, _I, _I, _P, _P, _I, _P, _P, _P, _P, _P, _I, _P, _P, _P, _P, _P, _P,
                                 _P, _P, _P, _P, _P, _P],
    "eslam_depth_samples": [_RP, _P, _I, _P, _P, _P, _P, _P, _P, _P, _P],
    "eslam_importance_samples": [_FP, _P, _P, _RP, _P, _P, _P, _P, _I, _P, _P, _P, _P, _P],
    "eslam_render_forward": [_FP, _P, _P, _P, _P, _I, _I, _P, _P, _P, _P, _P],
    "eslam_render_forward_act": [_FP, _P, _P, _P, _P, _I, _I, _P, _P, _P, _P, _P, _P, _P],
    "eslam_render_backward": [_FP, _P, _P, _P, _P, _I, _I, _P, _P, _P, _P, _P, _P, _P],
    "eslam_track_mask": [_P, _P, _P, _I, _P, _P, _P, _P],
    "eslam_loss_backward": [_FP, _P, _CP, _RP, _P, _P, _P, _P, _P, _P, _P, _I, _P, _P, _P, _I, _P, _P, _P, _P],
    "eslam_pose_backward_act": [_FP, _P, _CP, _RP, _P, _P, _P, _P, _P, _P, _P, _I, _P, _P, _I, _P, _P, _P, _P, _P, _P],
    "eslam_loss_backward_q": [_FP, _P, _P, _P, _CP, _RP, _P, _P, _P, _P, _P, _P, _P, _I, _P, _P, _P, _I, _P, _P, _P, _P],
    "eslam_loss_backward_q_part": [_FP, _P, _P, _P, _CP, _RP, _P, _P, _P, _P, _P, _P, _P, _I, _P, _P, _P, _I, _P, _P, _I, _P],
    "eslam_pose_backward_q": [_FP, _P, _P, _CP, _RP, _P, _P, _P, _P, _P, _P, _P, _I, _P, _P, _I, _P, _P, _P, _P, _P, _P],
    "eslam_adam_step": [_P, _P, _P, _P, _L, C.POINTER(C.c_int64), C.POINTER(C.c_double), _I, _I, _D, _D, _D, _P],
    "eslam_adam_step_sparse": [_P, _P, _P, _P, _L, C.POINTER(C.c_int64), C.POINTER(C.c_double), _I, _I, _D, _D, _D, _P,
                               _P],
    "eslam_pose_adam_step": [_P, _P, _P, _P, _I, _I, _D, _D, _I, _D, _D, _D, _P, _I, _P],
    "eslam_finalize_loss": [_RP, _P, _I, _P, _P, _P],
    "eslam_ingest_frame": [_P, _P, _I, _I, _I, _D, _D, _P, _P, _P],
    "eslam_ingest_frame_resized": [_P, _I, _I, _P, _I, _I, _I, _D, _D, _P, _P, _P],
    "eslam_undistort_u8": [_P, _P, _I, _I, _D, _D, _D, _D, C.POINTER(C.c_double), C.POINTER(C.c_double), _P],
    "eslam_ingest_frame_crop": [_P, _P, _I, _I, _I, _I, _I, _D, _D, _P, _P, _P],
    "eslam_matrix_to_pose": [_P, _P, _I, _P],
    "eslam_pose_to_matrix": [_P, _P, _I, _P],
    "eslam_keyframe_overlap": [_CP, _P, _P, _P, _I, _P, _I, _P, _I, _P, _P, _P],
    "eslam_mc_count": [_P, _I, _I, _I, _D, _P, _P, _P, _P],
    "eslam_mc_emit": [_P, _P, _P, _P, _I, _I, _I, _D, _P, _P, _P, _P, _P, _P],
    "eslam_cull_frame": [_P, _L, _P, _P, _CP, _D, _I, _P, _P],
    "eslam_exchange_counters": [C.POINTER(Peers), _P, C.POINTER(C.c_void_p), _I, _P, _P],
    "eslam_exchange_aux": [C.POINTER(Peers), _P, C.POINTER(C.c_void_p), _P, _I, _P, C.POINTER(C.c_void_p), _P, _I, _P],
    "eslam_q_adam_exchange": [C.POINTER(Peers), _FP, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), _P, _P, _P, _P, _P, _P,
                              _D, _D, _D, _I, _D, _D, _D, C.POINTER(C.c_void_p), _P, C.POINTER(C.c_void_p), _P, _I, _P,
                              C.POINTER(C.c_void_p), _P, _I, _P],
}

_lib = None


def load():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  myslam_b200 has no CPU or PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    lib.eslam_last_error.restype = C.c_char_p
    lib.eslam_last_error.argtypes = []
    lib.eslam_abi_version.restype = C.c_int
    lib.eslam_abi_version.argtypes = []
    lib.eslam_exchange_flag_words.restype = C.c_int
    lib.eslam_exchange_flag_words.argtypes = []
    lib.eslam_q_exchange_stage_floats.restype = C.c_int64
    lib.eslam_q_exchange_stage_floats.argtypes = [C.POINTER(FieldDesc), C.c_int]
    lib.eslam_mc_blocks.restype = C.c_int64
    lib.eslam_mc_blocks.argtypes = [C.c_int, C.c_int, C.c_int]
    lib.eslam_q_touched_bytes.restype = C.c_int
    lib.eslam_q_touched_bytes.argtypes = [C.POINTER(FieldDesc)]
    for name, argtypes in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = C.c_int
        fn.argtypes = argtypes
    if lib.eslam_abi_version() != ABI_VERSION:
        raise RuntimeError(f"libeslam_b200.so ABI {lib.eslam_abi_version()} != binding {ABI_VERSION}; rebuild")
    _lib = lib
    return lib


def ptr(t):
    """Device (or host) address of a tensor; None -> NULL."""
    if t is None:
        return None
    assert t.is_contiguous(), "eslam kernels take contiguous buffers"
    return t.data_ptr()


_RAW_STREAM = getattr(torch._C, "_cuda_getCurrentRawStream", None)
_CUR_DEVICE = getattr(torch._C, "_cuda_getDevice", None)
_STREAM_OVERRIDE = None


def stream():
    """cudaStream_t the next launch goes to: torch's current stream of the current device (through the raw binding --
    torch.cuda.current_stream() costs as much host time as a launch), or the handle set by `on_stream`."""
    if _STREAM_OVERRIDE is not None:
        return _STREAM_OVERRIDE
    if _RAW_STREAM is not None and _CUR_DEVICE is not None:
        return _RAW_STREAM(_CUR_DEVICE())
    return torch.cuda.current_stream().cuda_stream


class on_stream:
    """`with on_stream(s.cuda_stream):` -- launches of THIS library inside go to that stream.  Only a variable is set:
    torch's own current stream does not change (torch ops inside still need torch.cuda.stream), which is what makes
    it cheap enough for a loop that switches streams several times per iteration.  Process-wide, not per thread: the
    tracker and the mapper are separate processes (ESLAM.py:246-260), each launching from one thread."""

    def __init__(self, handle):
        self.handle = handle

    def __enter__(self):
        global _STREAM_OVERRIDE
        self.prev, _STREAM_OVERRIDE = _STREAM_OVERRIDE, self.handle

    def __exit__(self, *exc):
        global _STREAM_OVERRIDE
        _STREAM_OVERRIDE = self.prev


LAUNCHES = 0  # kernels launched through the ABI (bench.py reports it as gpu_launches)
_NO_KERNEL = ("eslam_bind_decoders",)
BOUND_DECODERS = {}  # device index -> (id(FieldStore), generation) whose decoders the constant bank holds (field.py)
_FN = {}


def call(name, *args):
    global LAUNCHES
    fn = _FN.get(name)
    if fn is None:
        fn = _FN[name] = getattr(load(), name)
    if name == "eslam_bind_decoders":
        BOUND_DECODERS.clear()  # whoever binds records what it bound afterwards (FieldStore.bind)
    else:
        LAUNCHES += 1
    rc = fn(*args)
    if rc != 0:
        raise RuntimeError(f"{name} failed ({rc}): {load().eslam_last_error().decode()}")


def require_cuda(t, what):
    if not (torch.is_tensor(t) and t.is_cuda):
        raise RuntimeError(f"{what}: expected a CUDA tensor; myslam_b200 has no CPU path")
