"""Operator-level wrappers over the C ABI (torch tensors in, raw device pointers across the boundary)."""
from __future__ import annotations

import torch

from . import _lib
from ._lib import ACT_NONE, DT_BF16, DT_FP16, check

_TORCH16 = {DT_BF16: torch.bfloat16, DT_FP16: torch.float16}


def dtype_code(t: torch.dtype) -> int:
    if t == torch.bfloat16:
        return DT_BF16
    if t == torch.float16:
        return DT_FP16
    raise TypeError(f"operand dtype must be bfloat16 or float16, got {t}")


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t):
    return 0 if t is None else t.data_ptr()


def gemm(a16, w16, bias=None, residual=None, act=ACT_NONE, want16=True, want32=False, m=None):
    """act(a16 @ w16.T + bias) (+ residual): a16 [M,K] (row stride free), w16 [N,K]."""
    lib = _lib.load()
    assert a16.is_cuda and a16.dim() == 2 and w16.dim() == 2 and a16.stride(1) == 1 and w16.stride(1) == 1
    M = a16.shape[0] if m is None else m
    K = a16.shape[1]
    N = w16.shape[0]
    code = dtype_code(a16.dtype)
    out16 = torch.empty((M, N), dtype=a16.dtype, device=a16.device) if want16 else None
    out32 = torch.empty((M, N), dtype=torch.float32, device=a16.device) if want32 else None
    check(lib.mmr_gemm(a16.data_ptr(), a16.stride(0), w16.data_ptr(), w16.stride(0), M, N, K, _ptr(bias),
                       _ptr(residual), residual.stride(0) if residual is not None else 0, _ptr(out16), N,
                       _ptr(out32), N, act, code, _stream()))
    return out16, out32


def layernorm(x32, gamma, beta, eps=1e-12, dtype=torch.bfloat16, want16=True, want32=True, scale=1.0,
              accumulate_into=None):
    lib = _lib.load()
    M, H = x32.shape
    out16 = torch.empty((M, H), dtype=dtype, device=x32.device) if want16 else None
    out32 = accumulate_into if accumulate_into is not None else (
        torch.empty((M, H), dtype=torch.float32, device=x32.device) if want32 else None)
    check(lib.mmr_layernorm(x32.data_ptr(), x32.stride(0), gamma.data_ptr(), beta.data_ptr(), eps, M, H,
                            _ptr(out16), H, _ptr(out32), H, scale, 1 if accumulate_into is not None else 0,
                            dtype_code(dtype), _stream()))
    return out16, out32


def attention(q, k, v, key_mask, B, Sq, Sk, heads=12):
    """q [B*Sq, >=heads*64], k/v [B*Sk, ...] (views with arbitrary row stride); key_mask int32 [B,Sk] or None."""
    lib = _lib.load()
    out = torch.empty((B * Sq, heads * 64), dtype=q.dtype, device=q.device)
    check(lib.mmr_attention(q.data_ptr(), q.stride(0), k.data_ptr(), k.stride(0), v.data_ptr(), v.stride(0),
                            _ptr(key_mask), out.data_ptr(), out.stride(0), B, Sq, Sk, heads, dtype_code(q.dtype),
                            _stream()))
    return out


def cast16(x32, dtype=torch.bfloat16):
    lib = _lib.load()
    out = torch.empty(x32.shape, dtype=dtype, device=x32.device)
    check(lib.mmr_cast16(x32.data_ptr(), out.data_ptr(), x32.numel(), dtype_code(dtype), _stream()))
    return out


def gemm_layernorm(a16, w16, bias, x32, gamma, beta, eps=1e-12):
    """x32 <- LN(a16 @ w16.T + bias + x32) * gamma + beta in place (fused kernel); returns (x16, x32)."""
    lib = _lib.load()
    M, K = a16.shape
    assert w16.shape[0] == 768 and x32.shape == (M, 768) and x32.is_contiguous()
    code = dtype_code(a16.dtype)
    if not lib.mmr_gemm_layernorm_supported(M, K, code):
        raise _lib.MmrError(f"fused GEMM+LayerNorm not available for M={M} K={K} on this device")
    x16 = torch.empty((M, 768), dtype=a16.dtype, device=a16.device)
    check(lib.mmr_gemm_layernorm(a16.data_ptr(), a16.stride(0), w16.data_ptr(), w16.stride(0), M, K, bias.data_ptr(),
                                 x32.data_ptr(), 768, gamma.data_ptr(), beta.data_ptr(), eps, x16.data_ptr(), 768,
                                 x32.data_ptr(), 768, code, _stream()))
    return x16, x32
