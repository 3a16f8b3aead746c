"""Launches the hot kernels at ViT shapes for ncu (one short process per kernel family).
usage: python tools/prof_kernels.py {attn|gemm}"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from visrag_b200 import ops  # noqa: E402

dev = "cuda"
torch.manual_seed(0)
what = sys.argv[1]
S = int(os.environ.get("PROF_SLICES", 32))
if what == "attn":
    N, nh, hd, hs = 1024, 16, 72, 80
    qkv = torch.zeros(S * N, 3, nh, hs, device=dev)
    qkv[..., :hd] = torch.randn(S * N, 3, nh, hd, device=dev)
    ones = os.environ.get("PROF_ONES", "1") == "1"
    if ones:
        qkv[:, 2, :, hd] = 1.0
    qkv = qkv.reshape(S * N, 3 * nh * hs).bfloat16()
    cu = torch.arange(0, (S + 1) * N, N, dtype=torch.int32, device=dev)
    out = torch.zeros(S * N, nh * hd, dtype=torch.bfloat16, device=dev)
    for _ in range(3):
        ops.attention(qkv, qkv, qkv, q_col0=0, k_col0=nh * hs, v_col0=2 * nh * hs, head_stride=hs, head_dim=hd, heads=nh,
                      batch=S, cu_k=cu, max_k=N, cu_q=cu, max_q=N, causal=False, scale=hd ** -0.5, out=out, v_ones_column=ones)
    torch.cuda.synchronize()
else:
    M = S * 1024
    a = (torch.randn(M, 1152, device=dev) * 0.5).bfloat16()
    a2 = (torch.randn(M, 4304, device=dev) * 0.5).bfloat16()
    w_qkv = (torch.randn(3840, 1152, device=dev) * 0.03).bfloat16()
    w_proj = (torch.randn(1152, 1152, device=dev) * 0.03).bfloat16()
    w_fc1 = (torch.randn(4304, 1152, device=dev) * 0.03).bfloat16()
    w_fc2 = (torch.randn(1152, 4304, device=dev) * 0.03).bfloat16()
    b1152, b3840, b4304 = torch.randn(1152, device=dev), torch.randn(3840, device=dev), torch.randn(4304, device=dev)
    x = torch.randn(M, 1152, device=dev)
    for _ in range(2):
        ops.gemm(a, w_qkv, bias=b3840)                                                   # launch 0: bf16 out
        ops.gemm(a, w_proj, bias=b1152, resid=x, out=x, out_dtype=torch.float32)         # launch 1: fp32 resid, short K
        ops.gemm(a, w_fc1, bias=b4304, gelu=True)                                        # launch 2: GELU
        ops.gemm(a2, w_fc2, bias=b1152, resid=x, out=x, out_dtype=torch.float32)         # launch 3: fp32 resid, long K
    torch.cuda.synchronize()
