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


REAL_PAGES = os.path.join(GOLDEN_DIR, "real_pages.npz")


def gen_real_pages():
    """The reference's own example inputs, kept as the ENCODED bytes the reference ships (data, not code): the two
    (query, page) rows of examples/training_data/0.parquet and the demo's cat/dog photos
    (visrag_scripts/demo/retriever/test_image, README.md:315-319). Decoded with PIL at test time."""
    import pyarrow.parquet as pq
    from oracle.reference_shim import REF_ROOT

    t = pq.read_table(os.path.join(REF_ROOT, "examples", "training_data", "0.parquet"))
    out, queries = {}, []
    for i in range(t.num_rows):
        out[f"parquet{i}"] = np.frombuffer(t.column("image")[i].as_py()["bytes"], dtype=np.uint8)
        queries.append(t.column("query")[i].as_py())
    for n in ("cat.jpeg", "dog.jpg"):
        with open(os.path.join(REF_ROOT, "visrag_scripts", "demo", "retriever", "test_image", n), "rb") as f:
            out[n.split(".")[0]] = np.frombuffer(f.read(), dtype=np.uint8)
    np.savez(REAL_PAGES, queries=np.asarray(queries), **out)
    print("real pages:", {k: v.size for k, v in out.items()})


def full_v2_spec():
    """>= 32 pages (>= 8 multi-slice + the reference's 4 real example images) and >= 8 queries: a corpus on which the
    top-5 ranking is a real statement (VERDICT r01 'next' #1)."""
    single = [(448, 448), (400, 500), (300, 600), (224, 224), (336, 336), (420, 420), (500, 390), (448, 448), (360, 540),
              (640, 300), (448, 448), (280, 280), (512, 384), (384, 512), (448, 448), (330, 600), (600, 330), (448, 440),
              (224, 224), (436, 452)]
    multi = [(700, 900), (640, 480), (1000, 700), (900, 450), (1344, 336), (600, 600), (800, 1000), (448, 900)]
    spec = [{"kind": "doc", "size": list(s), "seed": 5000 + i} for i, s in enumerate(single + multi)]
    spec += [{"kind": "noise", "size": list(s), "seed": 6000 + i} for i, s in enumerate([(448, 448), (448, 448), (224, 224), (640, 480)])]
    spec += [{"kind": "real", "name": n} for n in ("parquet0", "parquet1", "cat", "dog")]
    return spec


def gen_spec_case(name, cfg, weight_seed, spec, n_queries, query_seed, topk, batch=6):
    """Like gen_model_case, for a page list described by a spec (see tests/helpers.pages_from_spec)."""
    import time

    import torch
    from oracle import reference_shim as RS
    from tests.helpers import pages_from_spec, real_queries
    from visrag_b200.tokenizer_stub import StubTokenizer
    from visrag_b200.weights import random_state_dict

    sd = random_state_dict(cfg, weight_seed)
    model = RS.build_reference_model(cfg, sd, attn_implementation="sdpa")
    tok = StubTokenizer(cfg.vocab)
    pages = pages_from_spec(spec)
    queries = synth_queries(n_queries, query_seed) + [QUERY_PREFIX + q for q in real_queries()]
    p_items = [{"id": f"d{i}", "text": "", "image": im} for i, im in enumerate(pages)]
    q_items = [{"id": f"q{i}", "text": t, "image": None} for i, t in enumerate(queries)]
    t0 = time.time()
    parts = []
    for s in range(0, len(p_items), batch):  # the reference's own batch loop, right-padded batches of mixed pages
        parts.append(RS.encode(model, tok, p_items[s:s + batch], False))
        print(f"  pages {s + len(parts[-1])}/{len(p_items)}  {time.time() - t0:.0f}s", flush=True)
    p = np.concatenate(parts)
    q = RS.encode(model, tok, q_items, True)
    S = torch.matmul(torch.from_numpy(q), torch.from_numpy(p).T)      # dense_retriever.py:25-30
    ts, ti = torch.topk(S, topk, dim=1)
    full = np.sort(S.numpy(), axis=1)[:, ::-1]
    gaps = full[:, :topk] - full[:, 1:topk + 1]
    np.savez(os.path.join(GOLDEN_DIR, f"{name}.npz"), config=json.dumps(cfg.to_dict()), weight_seed=weight_seed,
             page_spec=json.dumps(spec), queries=np.asarray(queries), query_seed=query_seed,
             page_reps=p.astype(np.float32), query_reps=q.astype(np.float32), topk_scores=ts.numpy(),
             topk_indices=ti.numpy(), min_gap=np.float32(gaps.min()))
    print(name, "pages", p.shape, "queries", q.shape, "min score gap inside top-(k+1):", gaps.min(), "median", np.median(gaps))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--full", action="store_true", help="also generate the full-size (3.1 B parameter) case")
    ap.add_argument("--full-v2", action="store_true", help="only generate full_v2 (36 pages, 10 queries, top-5; ~15 min of CPU)")
    ap.add_argument("--tiny-v2", action="store_true", help="only generate tiny_v2 (same corpus as full_v2, tiny model)")
    a = ap.parse_args()
    from visrag_b200.config import VisRAGConfig as _C

    if a.full_v2 or a.tiny_v2:
        if not os.path.exists(REAL_PAGES):
            gen_real_pages()
        if a.tiny_v2:
            gen_spec_case("tiny_v2", _C.tiny(), 1234, full_v2_spec(), 8, 25, 5)
        if a.full_v2:
            gen_spec_case("full_v2", _C.full(), 4321, full_v2_spec(), 8, 25, 5)
        return
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    from visrag_b200.config import VisRAGConfig

    gen_geometry()
    sizes = [(224, 224), (448, 448), (224, 224), (700, 900), (760, 141), (1200, 500), (320, 240), (448, 448)]
    gen_model_case("tiny_v1", VisRAGConfig.tiny(), 1234, sizes, 7, 4, 5, 5)
    if a.full:
        gen_model_case("full_v1", VisRAGConfig.full(), 4321, [(448, 448), (224, 224), (640, 480)], 17, 3, 15, 3)


if __name__ == "__main__":
    main()
