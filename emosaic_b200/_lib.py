"""ctypes binding of libemosaic_cuda.so (C ABI: include/emosaic_cuda.h).

There is no CPU fallback: if the shared library is missing, or no sm_100 GPU is present,
loading / ``Context()`` raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libemosaic_cuda.so")

EXPORTS = [
    "emo_abi_version", "emo_last_error", "emo_create", "emo_destroy", "emo_set_stream", "emo_sync",
    "emo_device_info", "emo_launch_count", "emo_timer_start", "emo_timer_stop", "emo_mark", "emo_mark_elapsed", "emo_dev_alloc", "emo_dev_free",
    "emo_host_alloc", "emo_host_free", "emo_copy_h2d", "emo_copy_d2h", "emo_analyse", "emo_analyse_dev",
    "emo_analyse_fused", "emo_analyse_fused_dev", "emo_set_library", "emo_set_library_dev", "emo_library_info", "emo_build_index",
    "emo_set_match_mode", "emo_match", "emo_topk", "emo_topk_dev", "emo_no_repeat",
    "emo_match_dev", "emo_compose", "emo_compose_dev", "emo_compose_overlay", "emo_compose_overlay_dev", "emo_mosaic", "emo_mosaic_dev", "emo_reserve", "emo_stats", "emo_stats_dev",
    "emo_resize", "emo_resize_dev", "emo_resize_taps",
    "emo_stripe_bounds", "emo_host_register", "emo_host_unregister",
    "emo_comm_unique_id", "emo_comm_init_rank", "emo_comm_info", "emo_comm_set_library", "emo_comm_set_library_dev",
    "emo_comm_broadcast_dev", "emo_comm_allgather_analysis_dev",
    "emo_group_create", "emo_group_destroy", "emo_group_size", "emo_group_ctx", "emo_group_set_library", "emo_group_analyse",
    "emo_group_analyse_fused", "emo_group_mosaic",
]

_lib = None


class EmosaicError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"[emo_status {code}] {msg}")
        self.code = code
        self.msg = msg


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(make -C emosaic_b200/csrc). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    vp, u8p, i32p, u32p = C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p
    sig = {
        "emo_abi_version": (C.c_int, []),
        "emo_last_error": (C.c_char_p, []),
        "emo_create": (C.c_int, [C.c_int, C.POINTER(vp)]),
        "emo_destroy": (None, [vp]),
        "emo_set_stream": (C.c_int, [vp, vp]),
        "emo_sync": (C.c_int, [vp]),
        "emo_device_info": (C.c_int, [vp, C.c_char_p, C.c_size_t, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
        "emo_launch_count": (C.c_uint64, [vp]),
        "emo_timer_start": (C.c_int, [vp]),
        "emo_timer_stop": (C.c_int, [vp, C.POINTER(C.c_float)]),
        "emo_mark": (C.c_int, [vp, C.c_uint32]),
        "emo_mark_elapsed": (C.c_int, [vp, C.c_uint32, C.c_uint32, C.POINTER(C.c_float)]),
        "emo_dev_alloc": (C.c_int, [vp, C.c_size_t, C.POINTER(vp)]),
        "emo_dev_free": (C.c_int, [vp, vp]),
        "emo_host_alloc": (C.c_int, [vp, C.c_size_t, C.POINTER(vp)]),
        "emo_host_free": (C.c_int, [vp, vp]),
        "emo_copy_h2d": (C.c_int, [vp, vp, vp, C.c_size_t]),
        "emo_copy_d2h": (C.c_int, [vp, vp, vp, C.c_size_t]),
        "emo_analyse": (C.c_int, [vp, u8p, C.c_uint64, C.c_uint32, C.c_uint32, u8p]),
        "emo_analyse_dev": (C.c_int, [vp, u8p, C.c_uint64, C.c_uint32, C.c_uint32, u8p]),
        "emo_analyse_fused": (C.c_int, [vp, u8p, C.c_uint64, C.c_uint32, u8p, u8p]),
        "emo_analyse_fused_dev": (C.c_int, [vp, u8p, C.c_uint64, C.c_uint32, u8p, u8p]),
        "emo_set_library": (C.c_int, [vp, u8p, u8p, C.c_uint32, C.c_uint32, C.c_uint32]),
        "emo_set_library_dev": (C.c_int, [vp, u8p, u8p, C.c_uint32, C.c_uint32, C.c_uint32]),
        "emo_library_info": (C.c_int, [vp, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
        "emo_build_index": (C.c_int, [vp]),
        "emo_set_match_mode": (C.c_int, [vp, C.c_int]),
        "emo_match": (C.c_int, [vp, u8p, C.c_uint32, C.c_uint32, i32p, u32p]),
        "emo_topk": (C.c_int, [vp, u8p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, u8p, i32p, u32p]),
        "emo_topk_dev": (C.c_int, [vp, u8p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, u8p, i32p, u32p]),
        "emo_no_repeat": (C.c_int, [vp, u8p, C.c_uint32, C.c_uint32, C.c_uint32, i32p, u32p, C.POINTER(C.c_uint64)]),
        "emo_match_dev": (C.c_int, [vp, u8p, C.c_uint32, C.c_uint32, i32p, u32p]),
        "emo_compose": (C.c_int, [vp, i32p, u8p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint8, u8p]),
        "emo_compose_dev": (C.c_int, [vp, i32p, u8p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint8, u8p]),
        "emo_compose_overlay": (C.c_int, [vp, i32p, C.c_uint32, C.c_uint32, u8p, C.c_uint32, C.c_uint32, C.c_uint8, u8p]),
        "emo_compose_overlay_dev": (C.c_int, [vp, i32p, C.c_uint32, C.c_uint32, u8p, C.c_uint32, C.c_uint32, C.c_uint8, u8p]),
        "emo_mosaic": (C.c_int, [vp, u8p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint8, i32p, u32p, u8p]),
        "emo_mosaic_dev": (C.c_int, [vp, u8p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint8, i32p, u32p, u8p]),
        "emo_reserve": (C.c_int, [vp, C.c_uint32, C.c_uint32, C.c_uint32]),
        "emo_stats": (C.c_int, [vp, i32p, u32p, C.c_uint64, C.c_uint32, vp, u32p]),
        "emo_stats_dev": (C.c_int, [vp, i32p, u32p, C.c_uint64, C.c_uint32, vp, u32p]),
        "emo_resize": (C.c_int, [vp, u8p] + [C.c_uint32] * 9 + [u8p]),
        "emo_resize_dev": (C.c_int, [vp, u8p] + [C.c_uint32] * 9 + [u8p]),
        "emo_resize_taps": (C.c_int, [C.c_uint32, C.c_uint32, vp, vp, vp, C.c_uint32, C.POINTER(C.c_uint32)]),
        "emo_stripe_bounds": (None, [C.c_uint64, C.c_int, C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
        "emo_host_register": (C.c_int, [vp, C.c_size_t]),
        "emo_host_unregister": (C.c_int, [vp]),
        "emo_comm_unique_id": (C.c_int, [vp]),
        "emo_comm_init_rank": (C.c_int, [vp, vp, C.c_int, C.c_int]),
        "emo_comm_info": (C.c_int, [vp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
        "emo_comm_set_library": (C.c_int, [vp, u8p, u8p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int]),
        "emo_comm_set_library_dev": (C.c_int, [vp, u8p, u8p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int]),
        "emo_comm_broadcast_dev": (C.c_int, [vp, vp, C.c_size_t, C.c_int]),
        "emo_comm_allgather_analysis_dev": (C.c_int, [vp, u8p, C.c_uint64, C.c_uint32, u8p]),
        "emo_group_create": (C.c_int, [C.POINTER(C.c_int), C.c_int, C.POINTER(vp)]),
        "emo_group_destroy": (None, [vp]),
        "emo_group_size": (C.c_int, [vp]),
        "emo_group_ctx": (vp, [vp, C.c_int]),
        "emo_group_set_library": (C.c_int, [vp, u8p, u8p, C.c_uint32, C.c_uint32, C.c_uint32]),
        "emo_group_analyse": (C.c_int, [vp, u8p, C.c_uint64, C.c_uint32, C.c_uint32, u8p]),
        "emo_group_analyse_fused": (C.c_int, [vp, u8p, C.c_uint64, C.c_uint32, u8p, u8p]),
        "emo_group_mosaic": (C.c_int, [vp, u8p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint8, i32p, u32p, u8p]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)  # AttributeError if the export is missing
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        raise EmosaicError(rc, load().emo_last_error().decode("utf-8", "replace"))
