"""Python launch wrappers over the C ABI. Each function validates shapes/dtypes, allocates the output with
torch (memory plumbing only) and enqueues ONE library call on the current CUDA stream."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib as L

_PROF = None  # list of (kind, start_event, end_event, flops) while profiling is on


def profile_begin() -> None:
    """Start recording one CUDA-event pair around every kernel launch (bench.py's roofline pass)."""
    global _PROF
    _PROF = []


def profiling() -> bool:
    """True while per-launch event profiling is on (the engine then avoids CUDA-graph capture and replay)."""
    return _PROF is not None


def profile_end():
    """Stop recording; returns {kind: (launches, total_ms, total_flops)} (synchronises the device)."""
    global _PROF
    rec, _PROF = _PROF, None
    torch.cuda.synchronize()
    out = {}
    for kind, e0, e1, flops in rec or []:
        n, ms, fl = out.get(kind, (0, 0.0, 0.0))
        out[kind] = (n + 1, ms + e0.elapsed_time(e1), fl + flops)
    return out


def _launch(kind: str, flops: float, fn, *args) -> None:
    if _PROF is None:
        L.check(fn(*args))
        return
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    L.check(fn(*args))
    e1.record()
    _PROF.append((kind, e0, e1, flops))


def _bf16_2d(t: torch.Tensor, name: str) -> None:
    if t.dtype != torch.bfloat16 or t.dim() != 2 or t.stride(1) != 1 or not t.is_cuda:
        raise ValueError(f"{name}: expected a CUDA bf16 matrix with unit inner stride, got {t.dtype} {tuple(t.shape)}")
    L.check_device(t)


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
    kind = "gemm" if _PROF is None else f"gemm:{M}x{N}x{K}:" + ("rope" if mode == L.VR_EPI_ROPE else "swiglu" if mode == L.VR_EPI_SWIGLU
                                                               else ("gelu" if gelu else "") + ("+resid" if resid is not None else "") +
                                                               ("f32" if out_dtype == torch.float32 else "bf16"))
    _launch(kind, 2.0 * M * N * K, L.lib().vr_gemm_tuned, a.data_ptr(), a.stride(0), w.data_ptr(), w.stride(0), L.VR_BF16,
            M, N, K, C.byref(e), block_n, L.stream_ptr())
    return out


def attention(
    q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, *, q_col0: int, k_col0: int, v_col0: int, head_stride: int,
    head_dim: int, heads: int, batch: int, cu_k: torch.Tensor, max_k: int, cu_q: Optional[torch.Tensor], max_q: int,
    causal: bool, scale: float, out: torch.Tensor, v_ones_column: bool = False,
) -> torch.Tensor:
    """softmax(QK^T*scale)V on tcgen05; q/k/v are bf16 token matrices (see include/visrag_b200.h)."""
    for t, n in ((q, "q"), (k, "k"), (v, "v"), (out, "out")):
        _bf16_2d(t, n)
    if cu_k.dtype != torch.int32 or (cu_q is not None and cu_q.dtype != torch.int32):
        raise ValueError("attention: cu_seqlens must be int32")
    p = L.AttnParams()
    p.q, p.ldq, p.q_rows = q.data_ptr(), q.stride(0), q.shape[0]
    p.k, p.ldk = k.data_ptr(), k.stride(0)
    p.v, p.ldv, p.kv_rows = v.data_ptr(), v.stride(0), k.shape[0]
    p.q_col0, p.k_col0, p.v_col0 = q_col0, k_col0, v_col0
    p.head_stride, p.head_dim, p.heads, p.batch = head_stride, head_dim, heads, batch
    p.cu_q, p.cu_k = L.ptr(cu_q), cu_k.data_ptr()
    p.max_q, p.max_k = max_q, max_k
    p.causal, p.scale = int(causal), float(scale)
    p.out, p.ldo = out.data_ptr(), out.stride(0)
    p.flags = L.VR_ATTN_V_ONES_COLUMN if v_ones_column else 0
    _launch("attention", 0.0, L.lib().vr_attention, C.byref(p), L.stream_ptr())
    return out


def im2col_norm(pixels: torch.Tensor, patch: int, ld_out: int) -> torch.Tensor:
    """uint8 [S,h,w,3] -> bf16 patch matrix [S*(h/p)*(w/p), ld_out] (normalised, zero padded columns)."""
    if pixels.dtype != torch.uint8 or pixels.dim() != 4 or pixels.shape[3] != 3 or not pixels.is_contiguous():
        raise ValueError("im2col_norm: expected contiguous uint8 [S,h,w,3]")
    L.check_device(pixels)
    S, h, w, _ = pixels.shape
    out = torch.empty((S * (h // patch) * (w // patch), ld_out), dtype=torch.bfloat16, device=pixels.device)
    _launch("im2col", 0.0, L.lib().vr_im2col_norm, pixels.data_ptr(), S, h, w, patch, out.data_ptr(), ld_out, L.stream_ptr())
    return out


def layernorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float, *, add: Optional[torch.Tensor] = None):
    """fp32 [M,D] -> bf16 LN(x); with ``add`` [P,D] also returns LN(x)+add[row % P] (bf16)."""
    M, D = x.shape
    L.check_device(x)
    out = torch.empty((M, D), dtype=torch.bfloat16, device=x.device)
    out2 = torch.empty_like(out) if add is not None else None
    _launch("norm", 0.0, L.lib().vr_layernorm, x.data_ptr(), x.stride(0), gamma.data_ptr(), beta.data_ptr(), eps, M, D,
            out.data_ptr(), out.stride(0), L.ptr(out2), L.ptr(add), 0 if add is None else add.shape[0], L.stream_ptr())
    return out if add is None else (out, out2)


def rmsnorm(x: torch.Tensor, gamma: torch.Tensor, eps: float) -> torch.Tensor:
    M, D = x.shape
    L.check_device(x)
    out = torch.empty((M, D), dtype=torch.bfloat16, device=x.device)
    _launch("norm", 0.0, L.lib().vr_rmsnorm, x.data_ptr(), x.stride(0), gamma.data_ptr(), eps, M, D, out.data_ptr(),
            out.stride(0), L.stream_ptr())
    return out


def build_lm_input(src: torch.Tensor, embed: torch.Tensor, scale_emb: float, vision: Optional[torch.Tensor]) -> torch.Tensor:
    T, D = src.shape[0], embed.shape[1]
    L.check_device(embed)
    h = torch.empty((T, D), dtype=torch.float32, device=embed.device)
    _launch("other", 0.0, L.lib().vr_build_lm_input, src.data_ptr(), T, D, embed.data_ptr(), scale_emb, L.ptr(vision),
            0 if vision is None else vision.stride(0), h.data_ptr(), h.stride(0), L.stream_ptr())
    return h


POOLING = {"wmean": 0, "mean": 1, "lasttoken": 2, "cls": 3}


def pool_norm(h: torch.Tensor, gamma: torch.Tensor, eps: float, cu: torch.Tensor, pooling: str, normalize: bool) -> torch.Tensor:
    B = cu.shape[0] - 1
    L.check_device(h)
    reps = torch.empty((B, h.shape[1]), dtype=torch.float32, device=h.device)
    _launch("other", 0.0, L.lib().vr_pool_norm, h.data_ptr(), h.stride(0), gamma.data_ptr(), eps, cu.data_ptr(), B, h.shape[1],
            POOLING[pooling], int(normalize), reps.data_ptr(), L.stream_ptr())
    return reps
