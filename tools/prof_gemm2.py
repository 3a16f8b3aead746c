"""One CTA-pair GEMM launch for ncu (tools/prof_session style)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from visrag_b200 import ops  # noqa: E402

bn = int(sys.argv[1]) if len(sys.argv) > 1 else 2
M, N, K = 16384, 8192, 4096
a = (torch.randn(M, K, device="cuda") * 0.5).bfloat16()
w = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
for _ in range(3):
    ops.gemm(a, w, block_n=bn)
torch.cuda.synchronize()
