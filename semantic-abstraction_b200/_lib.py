"""ctypes binding of libsemabs_b200.so (the C ABI declared in include/semabs_b200.h).

There is deliberately NO fallback: if the shared library is missing or a call fails, an exception is raised.
torch is used only as the owner of device memory / streams (`tensor.data_ptr()`).
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import torch

PKG_DIR = Path(__file__).resolve().parent
LIB_PATH = PKG_DIR / "libsemabs_b200.so"

_lib = None


class SemabsError(RuntimeError):
    pass


class GemmEpilogue(C.Structure):
    _fields_ = [
        ("bias", C.c_void_p),
        ("residual", C.c_void_p),
        ("aux16", C.c_void_p),
        ("aux_rows", C.c_int32),
        ("ld_aux", C.c_int32),
        ("out_aux16", C.c_void_p),
        ("ld_out_aux", C.c_int32),
        ("out_f32", C.c_void_p),
        ("ld_out", C.c_int32),
        ("out_f16", C.c_void_p),
        ("ld_out16", C.c_int32),
        ("out_f16_splits", C.c_int32),
        ("act", C.c_int32),
        ("scale_cols", C.c_int32),
        ("scale", C.c_float),
    ]


ACT_NONE, ACT_QUICKGELU, ACT_MUL_AUX16 = 0, 1, 2


class _CallProfile:
    """Optional CUDA-event bracket around EVERY C-ABI call, keyed by entry-point name (bench.py's per-kernel-family
    roofline table; never enabled inside a timed region: the events cost ~2 us per call)."""

    def __init__(self):
        self.on = False
        self.recs = {}
        self.meta = {}

    def enable(self):
        self.on, self.recs, self.meta = True, {}, {}

    def note(self, name, **kw):
        """accumulate algorithmic work (flops / bytes) declared by the op wrapper for the call that follows"""
        if self.on:
            m = self.meta.setdefault(name, {})
            for k, v in kw.items():
                m[k] = m.get(k, 0.0) + float(v)

    def collect(self):
        torch.cuda.synchronize()
        out = {}
        for name, evs in self.recs.items():
            out[name] = dict(ms=sum(a.elapsed_time(b) for a, b in evs), calls=len(evs), **self.meta.get(name, {}))
        self.on, self.recs, self.meta = False, {}, {}
        return out


CALL_PROFILE = _CallProfile()


class _ProfiledLib:
    def __init__(self, l):
        self._l = l

    def __getattr__(self, name):
        fn = getattr(self._l, name)
        if not name.startswith("semabs_") or name in ("semabs_last_error", "semabs_abi_version", "semabs_lamb_chunk_bytes"):
            return fn

        def call(*a):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = fn(*a)
            e1.record()
            CALL_PROFILE.recs.setdefault(name, []).append((e0, e1))
            return r

        return call


_profiled = None


def lib():
    """Load (once) and return the shared library; raises if it has not been built."""
    global _lib, _profiled
    if _lib is not None:
        if CALL_PROFILE.on:
            if _profiled is None:
                _profiled = _ProfiledLib(_lib)
            return _profiled
        return _lib
    if not LIB_PATH.exists():
        if os.environ.get("SEMABS_B200_AUTOBUILD", "0") == "1":
            from . import build as _build

            _build.build()
        else:
            raise SemabsError(
                f"{LIB_PATH} is missing: run `python __graft_entry__.py build` (nvcc, sm_100a). "
                "There is no CPU / PyTorch fallback for this path."
            )
    l = C.CDLL(str(LIB_PATH))
    l.semabs_last_error.restype = C.c_char_p
    l.semabs_abi_version.restype = C.c_int
    _lib = l
    return l


def check(status: int) -> None:
    if status != 0:
        msg = lib().semabs_last_error()
        raise SemabsError(f"libsemabs_b200 call failed (status {status}): {msg.decode() if msg else '?'}")


def ptr(t) -> C.c_void_p:
    if t is None:
        return C.c_void_p(0)
    assert t.is_cuda, "libsemabs_b200 works on device memory only (no CPU fallback)"
    return C.c_void_p(t.data_ptr())


def stream_ptr() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def i32(v) -> C.c_int32:
    return C.c_int32(int(v))


def f32(v) -> C.c_float:
    return C.c_float(float(v))
