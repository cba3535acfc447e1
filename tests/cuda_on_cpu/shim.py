"""TEST INFRASTRUCTURE ONLY -- a chatts_b200._cabi.Context whose library is libchatts_shim.so (kernel sources compiled by g++ against the
"CUDA on CPU" shim): the product's own ctypes wrappers, argument marshalling included, drive the product's own kernel source on host
memory.  Entry points whose source is not part of the shim build (tensor cores, TMA, clusters) are simply absent."""
import ctypes as C

from .build import build


def shim_context():
    from chatts_b200 import _cabi
    lib = C.CDLL(build())
    real = _cabi.load_library()
    for name in _cabi.SYMBOLS:
        try:
            fn = getattr(lib, name)
        except AttributeError:
            continue
        rf = getattr(real, name)
        fn.argtypes, fn.restype = rf.argtypes, rf.restype
    _cabi._stream = lambda: None                       # no CUDA streams on the host
    ctx = _cabi.Context.__new__(_cabi.Context)
    ctx.lib = lib
    h = C.c_void_p()
    assert lib.cts_ctx_create(0, C.byref(h)) == 0
    ctx.h, ctx.device, ctx.arch, ctx.launches, ctx._debug_sync = h, 0, lib.cts_arch().decode(), 0, False
    return ctx
