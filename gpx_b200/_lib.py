"""ctypes binding of libgpxb200.so (the C ABI declared in include/gpxb200.h).

There is no CPU fallback: if the shared library is missing or a call fails, an
exception is raised.  torch is used only for device buffers and streams.
"""
from __future__ import annotations

import ctypes as C
import threading
from pathlib import Path

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "libgpxb200.so"

_c_dp = C.c_void_p  # device pointers travel as void*
_ll = C.c_longlong
_dbl = C.c_double
_int = C.c_int

# name -> (restype, argtypes); mirrors include/gpxb200.h one to one
SIGNATURES = {
    "gpxb_version": (_int, []),
    "gpxb_last_error": (_int, [C.c_char_p, C.c_size_t]),
    "gpxb_ctx_create": (_int, [_int, C.POINTER(C.c_void_p)]),
    "gpxb_ctx_destroy": (_int, [C.c_void_p]),
    "gpxb_ctx_launches": (C.c_ulonglong, [C.c_void_p]),
    "gpxb_prof_enable": (_int, [C.c_void_p, _int]),
    "gpxb_prof_collect": (_int, [C.c_void_p, C.POINTER(_dbl), C.POINTER(_dbl), C.POINTER(_int)]),
    "gpxb_phase_enable": (_int, [C.c_void_p, _int]),
    "gpxb_phase_collect": (_int, [C.c_void_p, C.POINTER(_dbl), C.POINTER(_int), _int]),
    "gpxb_gram": (_int, [C.c_void_p, _int, _c_dp, _int, _c_dp, _int, _int, _dbl, _dbl, _c_dp, _ll, _int, C.c_void_p]),
    "gpxb_gemm": (_int, [C.c_void_p, _int, _int, _int, _int, _int, _dbl, _c_dp, _ll, _c_dp, _ll, _dbl, _c_dp, _ll, _int, C.c_void_p]),
    "gpxb_potrf_inv_size": (C.c_size_t, [_int]),
    "gpxb_potrf": (_int, [C.c_void_p, _c_dp, _ll, _int, _c_dp, C.c_void_p]),
    "gpxb_trsm": (_int, [C.c_void_p, _c_dp, _ll, _c_dp, _int, _c_dp, _ll, _int, _int, C.c_void_p]),
    "gpxb_logdet": (_int, [C.c_void_p, _c_dp, _ll, _int, _c_dp, C.c_void_p]),
    "gpxb_potri": (_int, [C.c_void_p, _c_dp, _ll, _c_dp, _int, _c_dp, _ll, C.c_void_p]),
    "gpxb_gpr_lml_grad": (_int, [C.c_void_p, _int, _c_dp, _c_dp, _int, _int, _int, _dbl, _dbl, _int, _int, _c_dp, C.c_void_p]),
    "gpxb_gpr_fit": (_int, [C.c_void_p, _int, _c_dp, _c_dp, _int, _int, _int, _dbl, _dbl, _int, _c_dp, _c_dp, C.c_void_p]),
    "gpxb_gpr_predict": (_int, [C.c_void_p, _int, _c_dp, _int, _int, _c_dp, _int, _c_dp, _int, _c_dp, _dbl, _dbl, _c_dp, _c_dp, C.c_void_p]),
    "gpxb_gram_composite_fits": (_int, [_int, C.c_void_p]),
    "gpxb_gram_composite": (_int, [C.c_void_p, _int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, _int, _int, C.c_void_p, _int, _dbl, _int, _int, _c_dp, _ll, C.c_void_p]),
    "gpxb_rpcholesky_composite": (_int, [C.c_void_p, _int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, _int, C.c_void_p, _int, C.c_uint32, C.c_uint32, _int, _int, _c_dp, _c_dp, C.c_void_p]),
    "gpxb_gpr_iter_composite": (_int, [C.c_void_p, _int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, _int, C.c_void_p, _int, _c_dp, _dbl, _int, _int, C.c_uint32, C.c_uint32, _int, _dbl, _dbl, _int, _c_dp, _c_dp, C.POINTER(_int), _c_dp, _ll, _int, _int, _c_dp, _c_dp, C.c_void_p, C.c_void_p]),
    "gpxb_gram_dl_contract": (_int, [C.c_void_p, _int, _c_dp, _int, _c_dp, _int, _int, _dbl, _c_dp, _ll, _int, _c_dp, C.c_void_p]),
    "gpxb_elementwise": (_int, [C.c_void_p, _int, _ll, _dbl, _c_dp, _dbl, _c_dp, _c_dp, C.c_void_p]),
    "gpxb_hess_gram": (_int, [C.c_void_p, _int, _c_dp, _c_dp, _int, _int, _c_dp, _c_dp, _int, _int, _int, _dbl, _dbl, _c_dp, _ll, _int, C.c_void_p]),
    "gpxb_gpr_lml_derivs": (_int, [C.c_void_p, _int, _c_dp, _c_dp, _c_dp, _int, _int, _int, _dbl, _dbl, _c_dp, C.c_void_p]),
    "gpxb_gpr_lml_grad_derivs": (_int, [C.c_void_p, _int, _c_dp, _c_dp, _c_dp, _int, _int, _int, _dbl, _dbl, _int, _c_dp, C.c_void_p]),
    "gpxb_gpr_predict_y_derivs": (_int, [C.c_void_p, _int, _c_dp, _c_dp, _int, _int, _c_dp, _int, _dbl, _dbl, _c_dp, C.c_void_p]),
    "gpxb_gpr_predict_derivs_full": (_int, [C.c_void_p, _int, _c_dp, _c_dp, _int, _int, _int, _c_dp, _c_dp, _int, _int, _c_dp, _dbl, _dbl, _dbl, _c_dp, _c_dp, C.c_void_p]),
    "gpxb_gpr_predict_y_derivs_full": (_int, [C.c_void_p, _int, _c_dp, _c_dp, _int, _int, _int, _c_dp, _int, _c_dp, _dbl, _dbl, _dbl, _c_dp, _c_dp, C.c_void_p]),
    "gpxb_hess_matvec": (_int, [C.c_void_p, _int, _c_dp, _c_dp, _int, _int, _c_dp, _c_dp, _int, _int, _int, _dbl, _dbl, _c_dp, _ll, _int, _c_dp, _ll, C.c_void_p]),
    "gpxb_gpr_fit_derivs_iter": (_int, [C.c_void_p, _int, _c_dp, _c_dp, _c_dp, _int, _int, _int, _dbl, _dbl, _int, C.c_uint32, C.c_uint32, _int, _dbl, _dbl, _int, _c_dp, _c_dp, C.POINTER(_int), C.c_void_p]),
    "gpxb_hess_matvec_dl": (_int, [C.c_void_p, _int, _c_dp, _c_dp, _int, _int, _int, _dbl, _c_dp, _ll, _int, _c_dp, _ll, C.c_void_p]),
    "gpxb_lanczos_grad_derivs": (_int, [C.c_void_p, _int, _c_dp, _c_dp, _int, _int, _int, _dbl, _dbl, _c_dp, _ll, _int, _int, _c_dp, _c_dp, _c_dp, _c_dp, _c_dp, C.c_void_p]),
    "gpxb_sgpr_fit_derivs_iter": (_int, [C.c_void_p, _int, _c_dp, _c_dp, _c_dp, _int, _int, _int, _c_dp, _c_dp, _int, _int, _dbl, _dbl, _dbl, _dbl, _int, _c_dp, C.POINTER(_int), C.c_void_p]),
    "gpxb_gpr_td_fit_iter": (_int, [C.c_void_p, _int, _c_dp, _c_dp, _c_dp, _c_dp, _int, _int, _int, _int, _dbl, _dbl, _dbl, _dbl, _int, _int, C.c_uint32, C.c_uint32, _int, _dbl, _dbl, _int, _c_dp, _c_dp, C.POINTER(_int), C.c_void_p]),
    "gpxb_lanczos_td": (_int, [C.c_void_p, _int, _c_dp, _c_dp, _int, _int, _int, _int, _dbl, _dbl, _dbl, _dbl, _c_dp, _ll, _int, _int, _c_dp, _c_dp, _c_dp, C.c_void_p]),
    "gpxb_rpcholesky_td": (_int, [C.c_void_p, _int, _c_dp, _c_dp, _int, _int, _int, _int, _dbl, C.c_uint32, C.c_uint32, _int, _int, _c_dp, _c_dp, C.c_void_p]),
    "gpxb_rpcholesky_derivs": (_int, [C.c_void_p, _int, _c_dp, _c_dp, _int, _int, _int, _dbl, C.c_uint32, C.c_uint32, _int, _int, _c_dp, _c_dp, C.c_void_p]),
    "gpxb_lanczos_derivs": (_int, [C.c_void_p, _int, _c_dp, _c_dp, _int, _int, _int, _dbl, _dbl, _c_dp, _ll, _int, _int, _c_dp, _c_dp, _c_dp, C.c_void_p]),
    "gpxb_td_gram": (_int, [C.c_void_p, _int, _c_dp, _c_dp, _int, _int, _c_dp, _c_dp, _int, _int, _int, _dbl, _dbl, _dbl, _c_dp, _ll, _int, C.c_void_p]),
    "gpxb_gpr_td_lml_grad": (_int, [C.c_void_p, _int, _c_dp, _c_dp, _c_dp, _c_dp, _int, _int, _int, _dbl, _dbl, _dbl, _int, _int, _c_dp, C.c_void_p]),
    "gpxb_gpr_td_fit": (_int, [C.c_void_p, _int, _c_dp, _c_dp, _c_dp, _c_dp, _int, _int, _int, _dbl, _dbl, _dbl, _int, _c_dp, _c_dp, C.c_void_p]),
    "gpxb_gpr_td2_lml_grad": (_int, [C.c_void_p, _int, _c_dp, _c_dp, _c_dp, _c_dp, _int, _int, _int, _int, _dbl, _dbl, _dbl, _dbl, _int, _int, _c_dp, C.c_void_p]),
    "gpxb_gpr_td2_fit": (_int, [C.c_void_p, _int, _c_dp, _c_dp, _c_dp, _c_dp, _int, _int, _int, _int, _dbl, _dbl, _dbl, _dbl, _int, _c_dp, _c_dp, C.c_void_p]),
    "gpxb_gpr_fit_derivs": (_int, [C.c_void_p, _int, _c_dp, _c_dp, _c_dp, _int, _int, _int, _dbl, _dbl, _c_dp, _c_dp, C.c_void_p]),
    "gpxb_kmatvec": (_int, [C.c_void_p, _int, _c_dp, _int, _ll, _c_dp, _int, _int, _dbl, _dbl, _c_dp, _ll, _int, _c_dp, _ll, C.c_void_p]),
    "gpxb_rpcholesky": (_int, [C.c_void_p, _int, _c_dp, _int, _int, _dbl, C.c_uint32, C.c_uint32, _int, _int, _c_dp, _c_dp, C.c_void_p]),
    "gpxb_rademacher_probes": (_int, [C.c_void_p, C.c_uint32, C.c_uint32, _int, _int, _int, _c_dp, C.c_void_p]),
    "gpxb_lanczos_set_restart": (_int, [C.c_void_p, _int, C.c_uint32, C.c_uint32, _int]),
    "gpxb_lanczos": (_int, [C.c_void_p, _int, _c_dp, _int, _int, _dbl, _dbl, _c_dp, _ll, _int, _int, _c_dp, _c_dp, _c_dp, C.c_void_p]),
    "gpxb_lanczos_grad": (_int, [C.c_void_p, _int, _c_dp, _int, _int, _dbl, _dbl, _c_dp, _ll, _int, _int, _c_dp, _c_dp, _c_dp, _c_dp, _c_dp, C.c_void_p]),
    "gpxb_kmatvec_dl": (_int, [C.c_void_p, _int, _c_dp, _int, _int, _dbl, _c_dp, _ll, _int, _c_dp, _ll, C.c_void_p]),
    "gpxb_gpr_lml_iter": (_int, [C.c_void_p, _int, _c_dp, _c_dp, _int, _int, _dbl, _dbl, _int, _int, C.c_uint32, C.c_uint32, _int, _dbl, _dbl, _int, _c_dp, _ll, _int, _int, _c_dp, _c_dp, C.POINTER(_int), _c_dp, _c_dp, C.c_void_p, C.c_void_p]),
    "gpxb_gpr_fit_iter": (_int, [C.c_void_p, _int, _c_dp, _c_dp, _int, _int, _dbl, _dbl, _int, _int, C.c_uint32, C.c_uint32, _int, _dbl, _dbl, _int, _c_dp, _c_dp, C.POINTER(_int), C.c_void_p]),
    "gpxb_sgpr_lml": (_int, [C.c_void_p, _int, _c_dp, _c_dp, _int, _int, _int, _c_dp, _int, _dbl, _dbl, _int, _c_dp, C.c_void_p]),
    "gpxb_sgpr_fit_iter": (_int, [C.c_void_p, _int, _c_dp, _c_dp, _int, _int, _c_dp, _int, _dbl, _dbl, _int, _dbl, _dbl, _int, _c_dp, _c_dp, C.POINTER(_int), C.c_void_p]),
    "gpxb_sgpr_lml_iter_grad_parts": (_int, [C.c_void_p, _int, _c_dp, _c_dp, _int, _int, _c_dp, _int, _dbl, _dbl, _int, _dbl, _dbl, _int, _c_dp, _c_dp, C.POINTER(_int), _c_dp, _ll, _int, _int, _c_dp, _c_dp, _c_dp, _c_dp, C.c_void_p, C.c_void_p]),
    "gpxb_sgpr_lml_iter_parts": (_int, [C.c_void_p, _int, _c_dp, _c_dp, _int, _int, _c_dp, _int, _dbl, _dbl, _int, _dbl, _dbl, _int, _c_dp, C.POINTER(_int), _c_dp, _ll, _int, _int, _c_dp, _c_dp, _c_dp, C.c_void_p]),
    "gpxb_sgpr_lml_grad_locs": (_int, [C.c_void_p, _int, _c_dp, _c_dp, _int, _int, _c_dp, _int, _dbl, _dbl, _int, _c_dp, _c_dp, C.c_void_p]),
    "gpxb_sgpr_lml_grad": (_int, [C.c_void_p, _int, _c_dp, _c_dp, _int, _int, _c_dp, _int, _dbl, _dbl, _int, _int, _c_dp, C.c_void_p]),
    "gpxb_sgpr_fit": (_int, [C.c_void_p, _int, _c_dp, _c_dp, _int, _int, _int, _c_dp, _int, _dbl, _dbl, _int, _c_dp, _c_dp, C.c_void_p]),
    "gpxb_sgpr_predict": (_int, [C.c_void_p, _int, _c_dp, _int, _int, _c_dp, _int, _c_dp, _int, _c_dp, _dbl, _dbl, _c_dp, _c_dp, C.c_void_p]),
    "gpxb_gather_rows": (_int, [C.c_void_p, _c_dp, _int, _c_dp, _int, _c_dp, C.c_void_p]),
    "gpxb_comm_unique_id": (_int, [C.c_void_p]),
    "gpxb_shard_range": (_int, [_ll, _int, _int, C.POINTER(_ll), C.POINTER(_ll)]),
    "gpxb_comm_init": (_int, [C.c_void_p, C.c_void_p, _int, _int]),
}

_lib = None
_lock = threading.Lock()


class GpxbError(RuntimeError):
    pass


def load() -> C.CDLL:
    """Load libgpxb200.so and declare every symbol of the header.  Fails loudly."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not LIB_PATH.exists():
            raise GpxbError(
                f"{LIB_PATH} not found: build it with `python -m gpx_b200._build` "
                "(there is no CPU fallback)"
            )
        lib = C.CDLL(str(LIB_PATH))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the symbol is missing
            fn.restype = res
            fn.argtypes = args
        _lib = lib
        return lib


def last_error() -> str:
    buf = C.create_string_buffer(1024)
    load().gpxb_last_error(buf, 1024)
    return buf.value.decode(errors="replace")


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise GpxbError(f"{what} failed (rc={rc}): {last_error()}")


class Context:
    """One library context per process/GPU (handles, launch counter)."""

    def __init__(self, device: int = 0):
        lib = load()
        self._h = C.c_void_p()
        check(lib.gpxb_ctx_create(int(device), C.byref(self._h)), "gpxb_ctx_create")
        self.device = int(device)
        self.lib = lib

    @property
    def handle(self):
        return self._h

    def launches(self) -> int:
        return int(self.lib.gpxb_ctx_launches(self._h))

    def close(self):
        if self._h:
            self.lib.gpxb_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass


_ctx = {}


def context(device: int | None = None) -> Context:
    import torch

    if not torch.cuda.is_available():
        raise GpxbError("gpx_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    if device is None:
        device = torch.cuda.current_device()
    if device not in _ctx:
        _ctx[device] = Context(device)
    return _ctx[device]


def stream_ptr():
    import torch

    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    """Device pointer of a contiguous float64/int64/uint32 CUDA tensor (or None)."""
    if t is None:
        return C.c_void_p(0)
    if not (t.is_cuda and t.is_contiguous()):  # a host pointer would fault inside a kernel: refuse it here
        raise GpxbError("device buffers must be contiguous CUDA tensors (there is no CPU path)")
    return C.c_void_p(t.data_ptr())
