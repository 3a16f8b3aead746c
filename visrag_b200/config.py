"""Architecture constants of VisRAG-Ret (SigLIP-so400m ViT -> Resampler -> MiniCPM-2B).

Values follow the reference: ViT `timm/models/vision_transformer.py:2612-2619` (26 of 27 blocks used,
`modeling_minicpmv.py:57-73`), resampler `modeling_minicpmv.py:75-82`, slicing defaults
`configuration_minicpm.py:197-222`, MiniCPM-2B fields `configuration_minicpm.py:109-160` with the public
checkpoint's values. ``tiny()`` is a reduced config for tests; it keeps every head
dimension (72 / 128 / 64) because the kernels are specialised on them.
"""
from __future__ import annotations

from dataclasses import dataclass, asdict


@dataclass(frozen=True)
class VisRAGConfig:
    # --- SigLIP ViT (vpm)
    patch_size: int = 14
    vit_dim: int = 1152
    vit_depth: int = 26          # blocks actually executed (27 defined, last dropped)
    vit_heads: int = 16          # head_dim 72
    vit_mlp: int = 4304
    vit_pos_grid: int = 27       # pos_embed is [1, 27*27, D] (384 // 14)
    ln_eps: float = 1e-6
    # --- Resampler
    query_num: int = 64          # 8 x 8 learned queries, heads = hidden // 128
    # --- MiniCPM decoder (llm)
    hidden: int = 2304
    layers: int = 40
    heads: int = 36              # head_dim 64, MHA
    inter: int = 5760
    vocab: int = 122753
    scale_emb: float = 12.0
    scale_depth: float = 1.4
    rms_eps: float = 1e-5
    rope_theta: float = 10000.0
    max_pos: int = 2048
    # --- slicing (configuration_minicpm.py:203-210)
    scale_resolution: int = 448
    max_slice_nums: int = 9
    slice_mode: bool = True

    @property
    def vit_head_dim(self) -> int:
        return self.vit_dim // self.vit_heads

    @property
    def rs_heads(self) -> int:
        return self.hidden // 128

    @property
    def head_dim(self) -> int:
        return self.hidden // self.heads

    @property
    def depth_scale(self) -> float:
        """Residual branch multiplier scale_depth / sqrt(num_layers) (modeling_minicpm.py:983-985)."""
        return self.scale_depth / (self.layers ** 0.5)

    def validate(self) -> None:
        assert self.vit_dim % self.vit_heads == 0 and self.vit_head_dim == 72, "ViT kernels are specialised on head_dim 72"
        assert self.hidden % 128 == 0, "resampler heads = hidden // 128"
        assert self.hidden % self.heads == 0 and self.head_dim == 64, "LM kernels are specialised on head_dim 64"
        assert self.vit_mlp % 8 == 0 and self.inter % 32 == 0
        assert int(self.query_num ** 0.5) ** 2 == self.query_num

    @classmethod
    def full(cls) -> "VisRAGConfig":
        return cls()

    @classmethod
    def tiny(cls) -> "VisRAGConfig":
        """2-block ViT (4 heads x 72), 2-layer LM (4 heads x 64, hidden 256 -> 2 resampler heads), vocab 512."""
        return cls(vit_dim=288, vit_depth=2, vit_heads=4, vit_mlp=1008, hidden=256, layers=2, heads=4, inter=640, vocab=512)

    def to_dict(self) -> dict:
        return asdict(self)
