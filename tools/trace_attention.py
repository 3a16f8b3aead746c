"""Cycle-level timeline of the shipped ViT attention kernel (debug build with -DVR_A_TRACE=1, see tools/build_variants.sh):
clock64 stamps of CTA 0's tile-A / tile-B softmax warps and of the MMA issuer over the first 96 key blocks, reduced to
average phase durations.   VR_LIB=build/libvr_trace.so python tools/trace_attention.py"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from visrag_b200 import _lib as L, ops  # noqa: E402

S, N, nh, hd, hs = 128, 1024, 16, 72, 80
qkv = torch.zeros(S * N, 3, nh, hs, device="cuda")
qkv[..., :hd] = torch.randn(S * N, 3, nh, hd, device="cuda")
qkv[:, 2, :, hd] = 1.0
qkv = qkv.reshape(S * N, 3 * nh * hs).bfloat16()
cu = torch.arange(0, (S + 1) * N, N, dtype=torch.int32, device="cuda")
out = torch.zeros(S * N, nh * hd, dtype=torch.bfloat16, device="cuda")
for _ in range(3):
    ops.attention(qkv, qkv, qkv, q_col0=0, k_col0=nh * hs, v_col0=2 * nh * hs, head_stride=hs, head_dim=hd, heads=nh, batch=S,
                  cu_k=cu, max_k=N, cu_q=cu, max_q=N, causal=False, scale=hd ** -0.5, out=out, v_ones_column=True)
torch.cuda.synchronize()
buf = np.zeros((3, 96, 8), dtype=np.int64)
lib = L.lib()
assert lib.vr_attention_trace_read(buf.ctypes.data_as(C.c_void_p)) == 0
names = ["wait s_full (S_x = QK^T ready)", "tcgen05.ld S (128 cols) + wait::ld", "row max (+ rescale)", "128 x (fma, ex2, pack)",
         "wait pv_done (P buffer free)", "tcgen05.st P + wait::st", "arrive p_full -> top of loop"]
for x, tile in enumerate("AB"):
    t = buf[x, 16:88].astype(np.float64)  # steady state
    d = np.diff(t[:, :7], axis=1)
    loop = t[1:, 0] - t[:-1, 6]
    period = np.diff(t[:, 0])
    print(f"tile {tile}: period {period.mean():7.0f} cycles per 128-key block (min {period.min():.0f}, max {period.max():.0f})")
    for i in range(6):
        print(f"    {names[i]:44s} {d[:, i].mean():7.0f}")
    print(f"    {names[6]:44s} {loop.mean():7.0f}")
m = buf[2, 16:88].astype(np.float64)
print("issuer: QK_A->QK_A %.0f  PV_A->PV_A %.0f  QK_A->PV_A(same block) %.0f  PV_A->PV_B %.0f  QK_A->QK_B %.0f" % (
    np.diff(m[:, 0]).mean(), np.diff(m[:, 1]).mean(), (m[:, 1] - m[:, 0]).mean(), (m[:, 3] - m[:, 1]).mean(), (m[:, 2] - m[:, 0]).mean()))
ta, tb = buf[0, 16:88, 3].astype(np.float64), buf[1, 16:88, 3].astype(np.float64)
print("exp phase of B starts %.0f cycles after A's (same block index)" % (tb - ta).mean())
