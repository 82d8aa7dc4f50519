"""ctypes access to tools/libemosaic_probe.so — measurement-only microbenchmarks (not part of the product ABI):
pipe issue rates (tools/probe/probe.cu) and the pinned-memory copy ceiling with n GPUs at once (tools/probe/hostcopy.cu)."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libemosaic_probe.so")
_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} not found: make -C tools/probe")
        lib = C.CDLL(LIB_PATH)
        lib.emo_probe_int_pipe.restype = C.c_int
        lib.emo_probe_int_pipe.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_double)]
        lib.emo_probe_host_copy.restype = C.c_int
        lib.emo_probe_host_copy.argtypes = [C.POINTER(C.c_int), C.c_int, C.c_size_t, C.c_size_t, C.c_int, C.c_int,
                                            C.POINTER(C.c_double), C.POINTER(C.c_double)]
        lib.emo_probe_host_copy_at.restype = C.c_int
        lib.emo_probe_host_copy_at.argtypes = lib.emo_probe_host_copy.argtypes + [C.c_int64]
        _lib = lib
    return _lib


def probe_int_pipe(device: int, which: int) -> float:
    """Thread-level instructions per second of one instruction class (see tools/probe/probe.cu for `which`)."""
    v = C.c_double()
    rc = load().emo_probe_int_pipe(device, which, C.byref(v))
    if rc:
        raise RuntimeError(f"emo_probe_int_pipe: cudaError {-rc}")
    return float(v.value)


def host_copy(devices, nbytes: int, chunk: int = 64 << 20, direction: str = "d2h", reps: int = 3, start_unix_ns: int = 0):
    """(aggregate GB/s, [per-device GB/s]) of pinned-memory copies with every listed device copying at the same time.
    start_unix_ns: wall-clock instant at which the timed copies start (several processes on one box: every rank passes the
    same instant); raises if the set-up finished after it."""
    devices = list(devices)
    arr = (C.c_int * len(devices))(*devices)
    agg = C.c_double()
    per = (C.c_double * len(devices))()
    rc = load().emo_probe_host_copy_at(arr, len(devices), nbytes, chunk, 0 if direction == "d2h" else 1, reps, C.byref(agg), per,
                                       int(start_unix_ns))
    if rc < 0:
        raise RuntimeError(f"emo_probe_host_copy: cudaError {-rc}")
    if rc == 1:
        raise RuntimeError("emo_probe_host_copy: set-up finished after the common start instant")
    return float(agg.value), [float(x) for x in per]
