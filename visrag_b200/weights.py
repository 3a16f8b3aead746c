"""Weight naming + synthetic weights.

Tensor names are the reference checkpoint's (module attribute names `modeling_minicpmv.py:36-40`; SURVEY.md
A.7), so a real ``VisRAG-Ret`` state_dict loads unchanged. ``random_state_dict`` draws seeded random weights
(no checkpoint is reachable offline); every value is rounded to bf16 so that the reference (fp32 CPU), the
oracle and the bf16 GPU engine all hold *identical* parameters and differ only in activation arithmetic.
"""
from __future__ import annotations

from typing import Dict

import numpy as np
import torch

from .config import VisRAGConfig


def sincos_2d(embed_dim: int, grid_h: int, grid_w: int) -> np.ndarray:
    """2-D sin/cos table of the resampler (`resampler.py:38-90`), fp32 [grid_h*grid_w, embed_dim].

    Follows the reference literally: ``grid = meshgrid(w, h)`` (w first); the FIRST half of the channels
    encodes ``grid[0]`` (the w index), the second half ``grid[1]`` (the h index); each half is
    ``[sin(pos*omega), cos(pos*omega)]`` with ``omega_i = 10000^(-i/(D/4))``.
    """
    gh = np.arange(grid_h, dtype=np.float32)
    gw = np.arange(grid_w, dtype=np.float32)
    grid = np.stack(np.meshgrid(gw, gh), axis=0).reshape(2, -1)  # [2, H*W]; grid[0] = w index

    def one(dim: int, pos: np.ndarray) -> np.ndarray:
        omega = np.arange(dim // 2, dtype=np.float32)
        omega /= dim / 2.0
        omega = 1.0 / 10000 ** omega
        out = np.einsum("m,d->md", pos, omega)
        return np.concatenate([np.sin(out), np.cos(out)], axis=1)

    return np.concatenate([one(embed_dim // 2, grid[0]), one(embed_dim // 2, grid[1])], axis=1).astype(np.float32)


def expected_shapes(cfg: VisRAGConfig) -> Dict[str, tuple]:
    D, Hv, E, H, I = cfg.vit_dim, cfg.vit_mlp, cfg.hidden, cfg.hidden, cfg.inter
    s: Dict[str, tuple] = {
        "vpm.patch_embed.proj.weight": (D, 3, cfg.patch_size, cfg.patch_size),
        "vpm.patch_embed.proj.bias": (D,),
        "vpm.pos_embed": (1, cfg.vit_pos_grid ** 2, D),
        "vpm.norm.weight": (D,),
        "vpm.norm.bias": (D,),
        "resampler.query": (cfg.query_num, E),
        "resampler.pos_embed": (cfg.query_num, E),
        "resampler.proj": (E, E),
        "resampler.kv_proj.weight": (E, D),
        "resampler.attn.in_proj_weight": (3 * E, E),
        "resampler.attn.in_proj_bias": (3 * E,),
        "resampler.attn.out_proj.weight": (E, E),
        "resampler.attn.out_proj.bias": (E,),
        "llm.model.embed_tokens.weight": (cfg.vocab, H),
        "llm.model.norm.weight": (H,),
    }
    for ln in ("ln_q", "ln_kv", "ln_post"):
        s[f"resampler.{ln}.weight"] = (E,)
        s[f"resampler.{ln}.bias"] = (E,)
    for i in range(cfg.vit_depth):
        p = f"vpm.blocks.{i}."
        s[p + "norm1.weight"] = (D,)
        s[p + "norm1.bias"] = (D,)
        s[p + "attn.qkv.weight"] = (3 * D, D)
        s[p + "attn.qkv.bias"] = (3 * D,)
        s[p + "attn.proj.weight"] = (D, D)
        s[p + "attn.proj.bias"] = (D,)
        s[p + "norm2.weight"] = (D,)
        s[p + "norm2.bias"] = (D,)
        s[p + "mlp.fc1.weight"] = (Hv, D)
        s[p + "mlp.fc1.bias"] = (Hv,)
        s[p + "mlp.fc2.weight"] = (D, Hv)
        s[p + "mlp.fc2.bias"] = (D,)
    for i in range(cfg.layers):
        p = f"llm.model.layers.{i}."
        s[p + "input_layernorm.weight"] = (H,)
        s[p + "post_attention_layernorm.weight"] = (H,)
        for n in ("q_proj", "k_proj", "v_proj", "o_proj"):
            s[p + f"self_attn.{n}.weight"] = (H, H)
        s[p + "mlp.gate_proj.weight"] = (I, H)
        s[p + "mlp.up_proj.weight"] = (I, H)
        s[p + "mlp.down_proj.weight"] = (H, I)
    return s


def random_state_dict(cfg: VisRAGConfig, seed: int = 1234, device: str = "cpu") -> Dict[str, torch.Tensor]:
    """Seeded synthetic checkpoint: fp32 tensors whose values are exactly representable in bf16.

    Linear weights ~ N(0, 0.02); norm gains ~ 1 + 0.1 N(0,1); biases ~ 0.02 N(0,1); pos_embed ~ N(0, 0.02);
    resampler.proj ~ E^-0.5 N(0,1) (`resampler.py:133`); resampler.pos_embed = fixed 8x8 sincos
    (`resampler.py:116-118`). Generated on CPU per tensor (seed + index) so any subset is reproducible.
    """
    cfg.validate()
    out: Dict[str, torch.Tensor] = {}
    for idx, (name, shape) in enumerate(expected_shapes(cfg).items()):
        g = torch.Generator(device="cpu").manual_seed(seed * 100003 + idx)
        if name == "resampler.pos_embed":
            q = int(cfg.query_num ** 0.5)
            t = torch.from_numpy(sincos_2d(cfg.hidden, q, q))
        elif name == "resampler.proj":
            t = torch.randn(shape, generator=g) * (cfg.hidden ** -0.5)
        elif name.endswith("norm.weight") or name.endswith("norm1.weight") or name.endswith("norm2.weight") or \
                name.endswith("layernorm.weight") or ".ln_" in name and name.endswith(".weight"):
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif name.endswith(".bias"):
            t = 0.02 * torch.randn(shape, generator=g)
        elif name == "resampler.query":
            t = 0.02 * torch.randn(shape, generator=g)
        else:
            t = 0.02 * torch.randn(shape, generator=g)
        out[name] = t.to(torch.bfloat16).to(torch.float32).to(device)
    return out


def random_state_dict_device(cfg: VisRAGConfig, seed: int, device: str) -> Dict[str, torch.Tensor]:
    """Same distributions as `random_state_dict`, drawn directly on `device` (benchmarks: the 3.1 B parameter model
    takes minutes to draw on host cores). Values differ from the CPU stream, so goldens use `random_state_dict`."""
    cfg.validate()
    g = torch.Generator(device=device).manual_seed(seed)
    out: Dict[str, torch.Tensor] = {}
    for name, shape in expected_shapes(cfg).items():
        if name == "resampler.pos_embed":
            q = int(cfg.query_num ** 0.5)
            t = torch.from_numpy(sincos_2d(cfg.hidden, q, q)).to(device)
        elif name == "resampler.proj":
            t = torch.randn(shape, generator=g, device=device) * (cfg.hidden ** -0.5)
        elif name.endswith("norm.weight") or name.endswith("norm1.weight") or name.endswith("norm2.weight") or \
                name.endswith("layernorm.weight") or ".ln_" in name and name.endswith(".weight"):
            t = 1.0 + 0.1 * torch.randn(shape, generator=g, device=device)
        else:
            t = 0.02 * torch.randn(shape, generator=g, device=device)
        out[name] = t.to(torch.bfloat16)
    return out


def hf_config_dict(cfg: VisRAGConfig, name: str = "VisRAG-Ret-synthetic") -> dict:
    """`config.json` of a checkpoint directory (MiniCPM-V field names, `configuration_minicpm.py:109-160,197-222`);
    `_name_or_path` is what the reference driver dispatches the tokenizer class on (`driver/eval.py:102-110,306-316`)."""
    d = {
        "_name_or_path": name, "architectures": ["VisRAG_Ret"], "model_type": "minicpmv",
        "hidden_size": cfg.hidden, "num_hidden_layers": cfg.layers, "num_attention_heads": cfg.heads,
        "num_key_value_heads": cfg.heads, "intermediate_size": cfg.inter, "vocab_size": cfg.vocab, "scale_emb": cfg.scale_emb,
        "scale_depth": cfg.scale_depth, "dim_model_base": 256, "rms_norm_eps": cfg.rms_eps, "rope_theta": cfg.rope_theta,
        "max_position_embeddings": cfg.max_pos, "hidden_act": "silu", "query_num": cfg.query_num, "patch_size": cfg.patch_size,
        "scale_resolution": cfg.scale_resolution, "max_slice_nums": cfg.max_slice_nums, "slice_mode": cfg.slice_mode,
        "drop_vision_last_layer": True, "torch_dtype": "bfloat16",
    }
    if (cfg.vit_dim, cfg.vit_depth, cfg.vit_heads, cfg.vit_mlp) == (1152, 26, 16, 4304):
        d["vision_encoder"] = "vit_so400m_patch14_siglip_384"
    else:  # reduced test towers have no timm name
        d.update(vit_dim=cfg.vit_dim, vit_depth=cfg.vit_depth, vit_heads=cfg.vit_heads, vit_mlp=cfg.vit_mlp)
    return d


def save_checkpoint(path: str, cfg: VisRAGConfig, state_dict: Dict[str, torch.Tensor], name: str = "VisRAG-Ret-synthetic",
                    shards: int = 2) -> None:
    """Write a HF-style checkpoint directory (`config.json` + `model-0000i-of-0000n.safetensors`, bf16) that
    `DRModelForInference.build` / `VisRAGRetB200.from_pretrained` load - the layout of the public VisRAG-Ret checkpoint.
    Used to exercise the loading path without network access to the real weights."""
    import json
    import os

    from safetensors.torch import save_file

    os.makedirs(path, exist_ok=True)
    with open(os.path.join(path, "config.json"), "w") as f:
        json.dump(hf_config_dict(cfg, name), f, indent=1)
    names = list(state_dict)
    per = (len(names) + shards - 1) // shards
    for i in range(shards):
        part = {k: state_dict[k].detach().to("cpu", torch.bfloat16).contiguous() for k in names[i * per:(i + 1) * per]}
        if part:
            save_file(part, os.path.join(path, f"model-{i + 1:05d}-of-{shards:05d}.safetensors"))
