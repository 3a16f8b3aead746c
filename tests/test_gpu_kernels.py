"""Kernel-level parity on the B200 (through the C ABI) against plain PyTorch fp32 references of the same op.
Tolerances: fp32-output paths 2e-3 relative to max|ref| (bf16 operands, fp32 accumulate); bf16-output paths 1e-2
(one bf16 rounding of the result)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from tools import check_encode as CE  # noqa: E402
from tools import check_gemm as CG  # noqa: E402


@pytest.mark.parametrize("bn", [256, 128, 64, 3, 2, 4])  # 3 = feature-major accumulator kernel, 2 = CTA-pair kernel, 4 = pair, 192-wide tiles
def test_gemm_plain_shapes(bn):
    assert CG.case_basic(bn)


@pytest.mark.parametrize("bn", [256, 128, 64, 3, 2, 4])
def test_gemm_fused_epilogues(bn):
    assert CG.case_epilogues(bn)


def test_elementwise_kernels():
    assert CE.stage_elementwise()


def test_attention_three_shapes():
    assert CE.stage_attention()          # two-tile ping-pong kernel wherever max_q > 128


def test_attention_three_shapes_single_tile_kernel():
    assert CE.stage_attention(force_v1=True)


@pytest.mark.parametrize("variant", [2, 5])  # 2 = round-1 two-tile kernel (still the causal long-sequence kernel), 5 = the same with Q in tensor memory
def test_attention_three_shapes_other_variants(variant):
    assert CE.stage_attention(variant=variant)


def test_gemm_rejects_bad_arguments():
    from visrag_b200 import ops

    a = torch.zeros(128, 64, device="cuda", dtype=torch.bfloat16)
    w = torch.zeros(100, 64, device="cuda", dtype=torch.bfloat16)  # N not a multiple of 8
    with pytest.raises(RuntimeError, match="multiple of 8"):
        ops.gemm(a, w)
    with pytest.raises(ValueError):
        ops.gemm(a.float(), w)
