"""Shared test helpers: ctypes wrappers of the single-operator C-ABI entry points and small
torch restatements used as per-kernel references."""

from __future__ import annotations

import ctypes as C

import torch

from blurr_b200 import capi


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def pack_weight(W):
    """Row-major [N][K] -> the engine's tile-packed layout, through the C-ABI operator."""
    lib = capi.load_library()
    N, K = W.shape
    packed = torch.empty(N * K, device=W.device, dtype=torch.bfloat16)
    capi.check(lib.blurr_op_pack_weight(stream_ptr(), _ptr(W), N, K, K, _ptr(packed)))
    return packed


def op_gemm(W, X, epi, splitk=1, bias=None, out_cols=None, packed=True):
    """W [N][K], X [T][K] bf16 cuda.  Returns bf16 [T][out_cols] or fp32 partial [S][T][N].
    `packed`: stream the weights from the tile-packed layout (what the engine does)."""
    lib = capi.load_library()
    N, K = W.shape
    T = X.shape[0]
    assert X.shape[1] == K and W.is_contiguous() and X.is_contiguous()
    ldw = K
    if packed:
        W = pack_weight(W)
        ldw = 0
    if epi == capi.EPI_PARTIAL:
        partial = torch.zeros((max(splitk, 1), T, N), device=W.device, dtype=torch.float32)
        s = capi.check(lib.blurr_op_gemm(stream_ptr(), _ptr(W), N, K, ldw, _ptr(X), T, K, epi, splitk, None, None, 0,
                                         _ptr(partial)))
        torch.cuda.synchronize()
        return partial.view(-1)[: s * T * N].view(s, T, N)
    cols = out_cols if out_cols is not None else (N // 2 if epi == capi.EPI_GEGLU else N)
    out = torch.zeros((T, cols), device=W.device, dtype=torch.bfloat16)
    capi.check(lib.blurr_op_gemm(stream_ptr(), _ptr(W), N, K, ldw, _ptr(X), T, K, epi, 1, _ptr(bias), _ptr(out), cols,
                                 None))
    torch.cuda.synchronize()
    return out


def op_siglip_attention(qkv, batch, seq, heads, hidden):
    lib = capi.load_library()
    out = torch.zeros((batch * seq, hidden), device=qkv.device, dtype=torch.bfloat16)
    capi.check(lib.blurr_op_siglip_attention(stream_ptr(), _ptr(qkv), qkv.shape[1], batch, seq, heads, hidden,
                                             _ptr(out), hidden))
    torch.cuda.synchronize()
    return out


def op_joint_attention(few, q, q_per_sample, q_row_offset, kc, vc, n_keys, mask, batch, n_heads):
    """q [B*qps][n_heads*256]; kc/vc [B][slots][256]; mask [B][R][C] contiguous bf16."""
    lib = capi.load_library()
    out = torch.zeros_like(q)
    capi.check(lib.blurr_op_joint_attention(stream_ptr(), int(few), _ptr(q), q_per_sample, q_row_offset, _ptr(kc),
                                            _ptr(vc), kc.shape[1], n_keys, _ptr(mask), mask.stride(0),
                                            mask.stride(1), batch, n_heads, _ptr(out)))
    torch.cuda.synchronize()
    return out


def bf16_ulp_err(a: torch.Tensor, b: torch.Tensor):
    """max |a-b| measured in bf16 ulps of max(|a|,|b|) (>= 2^-8 floor)."""
    a, b = a.float(), b.float()
    mag = torch.maximum(a.abs(), b.abs()).clamp_min(2.0 ** -6)
    ulp = torch.pow(2.0, torch.floor(torch.log2(mag)) - 7)
    return ((a - b).abs() / ulp).max().item()


def report(name, got, ref):
    got, ref = got.float(), ref.float()
    d = (got - ref).abs()
    return (f"{name}: max_abs={d.max().item():.3e} mean_abs={d.mean().item():.3e} "
            f"ref_rms={ref.pow(2).mean().sqrt().item():.3e} mismatch_frac={(d > 0).float().mean().item():.4f} "
            f"max_ulp={bf16_ulp_err(got, ref):.1f}")
