"""Python launch wrappers over the C ABI. Each function validates shapes/dtypes, allocates the output with
torch (memory plumbing only) and enqueues ONE library call on the current CUDA stream."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib as L


def _bf16_2d(t: torch.Tensor, name: str) -> None:
    if t.dtype != torch.bfloat16 or t.dim() != 2 or t.stride(1) != 1 or not t.is_cuda:
        raise ValueError(f"{name}: expected a CUDA bf16 matrix with unit inner stride, got {t.dtype} {tuple(t.shape)}")


def gemm(
    a: torch.Tensor,
    w: torch.Tensor,
    *,
    bias: Optional[torch.Tensor] = None,
    gelu: bool = False,
    scale: float = 1.0,
    resid: Optional[torch.Tensor] = None,
    rowadd: Optional[torch.Tensor] = None,
    out: Optional[torch.Tensor] = None,
    out_dtype: torch.dtype = torch.bfloat16,
    mode: int = L.VR_EPI_LINEAR,
    positions: Optional[torch.Tensor] = None,
    rope_cos: Optional[torch.Tensor] = None,
    rope_sin: Optional[torch.Tensor] = None,
    rope_cols: int = 0,
    block_n: int = 0,
) -> torch.Tensor:
    """``out = epilogue(a @ w.T)`` on tcgen05. ``a`` [M,K] bf16, ``w`` [N,K] bf16 (nn.Linear layout).

    LINEAR: ``out = [resid +] scale * gelu?(a@w.T + bias) [+ rowadd[row % period]]``; bf16 or fp32 out.
    ROPE / SWIGLU: see include/visrag_b200.h.
    """
    _bf16_2d(a, "a")
    _bf16_2d(w, "w")
    M, K = a.shape
    N, K2 = w.shape
    if K != K2:
        raise ValueError(f"gemm: K mismatch {K} vs {K2}")
    out_cols = N // 2 if mode == L.VR_EPI_SWIGLU else N
    if mode != L.VR_EPI_LINEAR:
        out_dtype = torch.bfloat16
    if out is None:
        out = torch.empty((M, out_cols), dtype=out_dtype, device=a.device)
    if out.dtype != out_dtype or out.shape != (M, out_cols) or out.stride(1) != 1:
        raise ValueError("gemm: bad `out`")
    for t, n in ((bias, "bias"), (resid, "resid"), (rowadd, "rowadd"), (rope_cos, "rope_cos"), (rope_sin, "rope_sin")):
        if t is not None and (t.dtype != torch.float32 or not t.is_contiguous() and n != "resid"):
            raise ValueError(f"gemm: {n} must be contiguous fp32")
    if resid is not None and (resid.shape != out.shape or resid.stride(0) != out.stride(0)):
        raise ValueError("gemm: resid must match out (shape and row pitch)")
    e = L.GemmEpilogue()
    e.mode = mode
    e.out_dtype = L.VR_F32 if out_dtype == torch.float32 else L.VR_BF16
    e.act_gelu = int(gelu)
    e.scale = float(scale)
    e.bias = L.ptr(bias)
    e.resid = L.ptr(resid)
    e.rowadd = L.ptr(rowadd)
    e.rowadd_period = 0 if rowadd is None else rowadd.shape[0]
    e.positions = L.ptr(positions)
    e.rope_cos = L.ptr(rope_cos)
    e.rope_sin = L.ptr(rope_sin)
    e.rope_cols = rope_cols
    e.out = out.data_ptr()
    e.ldo = out.stride(0)
    L.check(
        L.lib().vr_gemm_tuned(
            a.data_ptr(), a.stride(0), w.data_ptr(), w.stride(0), L.VR_BF16, M, N, K, C.byref(e), block_n, L.stream_ptr()
        )
    )
    return out
