"""TEST INFRASTRUCTURE. Generates tests/golden/*.npz by running the REAL reference (through
oracle/reference_shim.py) in the build container:   python -m oracle.gen_golden [--full]

The .npz files hold everything needed to replay the case without the reference: the config, the weight seed
(weights are re-drawn by visrag_b200.weights.random_state_dict), the page sizes + pixel seed (pages are
re-drawn by synth_pages), the query strings, and the reference outputs (fp32 embeddings, score top-k,
slice geometry).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import warnings

import numpy as np
from PIL import Image

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.filterwarnings("ignore")

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

from visrag_b200.synth import QUERY_PREFIX, synth_pages, synth_queries  # noqa: E402,F401


def geometry_cases():
    sizes = [(224, 224), (448, 448), (336, 336), (1344, 1344), (564, 3040), (1114, 1670), (700, 900), (1200, 500),
             (500, 1200), (640, 480), (2000, 300), (300, 2000), (449, 449), (1000, 1000), (896, 448), (447, 448)]
    rs = np.random.RandomState(11)
    for _ in range(150):
        sizes.append((int(rs.randint(100, 1500)), int(rs.randint(100, 1500))))
    return sizes


def gen_geometry():
    """Slice geometry from the reference's own slice_image (modeling_minicpmv.py:482-537)."""
    from oracle import reference_shim as RS

    RS._import_reference()
    from openmatch.modeling.modeling_minicpmv.modeling_minicpmv import slice_image

    rows = []
    for (w, h) in geometry_cases():
        src, patches, grid = slice_image(Image.new("RGB", (w, h)), 9, 448, 14)
        g = grid if grid is not None else [0, 0]
        pw, ph = (patches[0][0].size if patches else (0, 0))
        rows.append([w, h, src.size[0], src.size[1], g[0], g[1], pw, ph, sum(len(r) for r in patches)])
    np.savez(os.path.join(GOLDEN_DIR, "geometry_v1.npz"), cases=np.asarray(rows, dtype=np.int64),
             columns=np.asarray(["W", "H", "src_w", "src_h", "grid_x", "grid_y", "patch_w", "patch_h", "n_patches"]))
    print("geometry:", len(rows), "cases")


def gen_model_case(name, cfg, weight_seed, page_sizes, page_seed, n_queries, query_seed, topk):
    import torch
    from oracle import reference_shim as RS
    from visrag_b200.tokenizer_stub import StubTokenizer
    from visrag_b200.weights import random_state_dict

    sd = random_state_dict(cfg, weight_seed)
    model = RS.build_reference_model(cfg, sd, attn_implementation="sdpa")
    tok = StubTokenizer(cfg.vocab)
    pages = synth_pages(page_sizes, page_seed)
    queries = synth_queries(n_queries, query_seed)
    p_items = [{"id": f"d{i}", "text": "", "image": im} for i, im in enumerate(pages)]
    q_items = [{"id": f"q{i}", "text": t, "image": None} for i, t in enumerate(queries)]
    # the reference encodes in batches; padding must not matter -> encode pages in two uneven batches
    half = max(1, len(p_items) // 2)
    p = np.concatenate([RS.encode(model, tok, p_items[:half], False), RS.encode(model, tok, p_items[half:], False)])
    q = RS.encode(model, tok, q_items, True)
    # reference scoring: torch.matmul + torch.topk (dense_retriever.py:25-30)
    S = torch.matmul(torch.from_numpy(q), torch.from_numpy(p).T)
    ts, ti = torch.topk(S, min(topk, p.shape[0]), dim=1)
    np.savez(os.path.join(GOLDEN_DIR, f"{name}.npz"), config=json.dumps(cfg.to_dict()), weight_seed=weight_seed,
             page_sizes=np.asarray(page_sizes, dtype=np.int64), page_seed=page_seed, queries=np.asarray(queries),
             query_seed=query_seed, page_reps=p.astype(np.float32), query_reps=q.astype(np.float32),
             topk_scores=ts.numpy(), topk_indices=ti.numpy())
    print(name, "pages", p.shape, "queries", q.shape)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--full", action="store_true", help="also generate the full-size (3.1 B parameter) case")
    a = ap.parse_args()
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    from visrag_b200.config import VisRAGConfig

    gen_geometry()
    sizes = [(224, 224), (448, 448), (224, 224), (700, 900), (760, 141), (1200, 500), (320, 240), (448, 448)]
    gen_model_case("tiny_v1", VisRAGConfig.tiny(), 1234, sizes, 7, 4, 5, 5)
    if a.full:
        gen_model_case("full_v1", VisRAGConfig.full(), 4321, [(448, 448), (224, 224), (640, 480)], 17, 3, 15, 3)


if __name__ == "__main__":
    main()
