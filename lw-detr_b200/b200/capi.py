"""ctypes binding of include/lwdetr_b200.h.

The shared library is built in-tree by tools/build.py (``__graft_entry__.build()``).  There is no
fallback: if the library is missing or a call fails, a RuntimeError is raised.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "lib", "liblwdetr_b200.so")

F16, BF16 = 0, 1
ACT_NONE, ACT_RELU, ACT_GELU, ACT_SILU = 0, 1, 2, 3

_lib = None

_vp, _i, _f = ctypes.c_void_p, ctypes.c_int, ctypes.c_float

_SIGNATURES = {
    "lwdetr_last_error": (ctypes.c_char_p, []),
    "lwdetr_abi_version": (_i, []),
    "lwdetr_gemm": (_i, [_i, _vp, _i, _i, _i, _vp, _i, _vp, _vp, _vp, _i, _i, _i, _vp, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "lwdetr_conv3x3": (_i, [_i, _vp, _i, _i, _i, _i, _i, _i, _vp, _i, _vp, _i, _vp, _i, _vp]),
}


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "lwdetr_b200: %s not found - run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU / PyTorch fallback)" % LIB_PATH)
        l = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def exported_symbols():
    return sorted(_SIGNATURES)


def check(rc, what):
    if rc != 0:
        raise RuntimeError("%s failed: %s" % (what, lib().lwdetr_last_error().decode()))


def ptr(t):
    """Device/host address of a torch tensor (or None)."""
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def dtype_code(torch_dtype):
    import torch
    if torch_dtype == torch.float16:
        return F16
    if torch_dtype == torch.bfloat16:
        return BF16
    raise RuntimeError("lwdetr_b200 computes in float16 or bfloat16, got %s" % torch_dtype)


def stream_ptr():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def gemm(A, W, out, bias=None, gamma=None, resid=None, resid_mod=0, act=ACT_NONE, M=None, K=None, N=None,
         rows_in=0, remap_rows=0, shuffle_cout=0, IH=0, IW=0):
    """out = epilogue(A @ W.T); A [M, K] (row stride A.stride(0)), W [N, K] contiguous."""
    import torch
    M = A.shape[0] if M is None else M
    K = A.shape[1] if K is None else K
    N = W.shape[0] if N is None else N
    check(lib().lwdetr_gemm(dtype_code(A.dtype), ptr(A), A.stride(0), M, K, ptr(W), N, ptr(bias), ptr(gamma),
                            ptr(resid), 0 if resid is None else resid.stride(0), resid_mod, act, ptr(out),
                            out.stride(0), 1 if out.dtype == torch.float32 else 0, rows_in, remap_rows,
                            shuffle_cout, IH, IW, stream_ptr()), "lwdetr_gemm")
    return out


def conv3x3(X, Wk, out, B, OH, OW, stride, Cin, bias=None, act=ACT_NONE):
    """X NHWC [B, s*OH, s*OW, ldx] (uses channels [0, Cin)), Wk [N, 9*Cin], out [B*OH*OW, ld_out]."""
    check(lib().lwdetr_conv3x3(dtype_code(X.dtype), ptr(X), X.stride(2), B, OH, OW, stride, Cin, ptr(Wk),
                               Wk.shape[0], ptr(bias), act, ptr(out), out.stride(0), stream_ptr()),
          "lwdetr_conv3x3")
    return out
