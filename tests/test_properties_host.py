"""Property-based checks (hypothesis) of the integer / host logic against the oracle and against brute force. CPU only.
The geometry properties run the oracle's `slice_image` (the restated reference algorithm) on real PIL images, so they
also exercise the exact sizes PIL is asked to produce."""
import math

import numpy as np
from hypothesis import given, settings
from hypothesis import strategies as st
from PIL import Image

from oracle import pil_resample as PR
from oracle import restated as O
from visrag_b200 import frontend as F
from visrag_b200 import host, inference as I, retriever as R
from visrag_b200.config import VisRAGConfig

CFG = VisRAGConfig.tiny()
SIDES = st.integers(min_value=8, max_value=2600)


@settings(max_examples=120, deadline=None)
@given(w=SIDES, h=SIDES)
def test_plan_equals_oracle_slice_image_geometry(w, h):
    """Thumbnail size, grid and cell size of any page = what the restated `slice_image` produces (sizes only: a 1-bit
    image keeps the oracle's PIL work cheap)."""
    img = Image.new("1", (w, h))
    source, patches, grid = O.slice_image(img, CFG.max_slice_nums, CFG.scale_resolution, CFG.patch_size)
    plan = host.plan_slices(w, h, CFG)
    assert plan.source_size == source.size
    if grid is None:
        assert plan.grid is None and plan.n_slices == 1
    else:
        assert plan.grid == tuple(grid) and plan.n_slices == 1 + sum(len(r) for r in patches)
        assert all(p.size == plan.cell_size for row in patches for p in row)
        assert len(patches) == plan.grid[1] and len(patches[0]) == plan.grid[0]
    # invariants the engine relies on
    for sw, sh in plan.slice_sizes():
        assert sw % CFG.patch_size == 0 and sh % CFG.patch_size == 0 and sw > 0 and sh > 0
    assert plan.n_slices <= 1 + CFG.max_slice_nums


@settings(max_examples=60, deadline=None)
@given(n_in=st.integers(1, 4000), n_out=st.integers(1, 1500))
def test_resample_tables_equal_oracle_and_are_well_formed(n_in, n_out):
    ksize, bounds, kk = F.resample_coeffs(n_in, n_out)
    oks, ob, okk = PR.precompute_coeffs(n_in, n_out)
    assert ksize == oks and np.array_equal(bounds, np.asarray(ob, dtype=np.int32).reshape(-1, 2))
    assert np.array_equal(kk, np.asarray(PR.normalize_coeffs_8bpc(okk), dtype=np.int32))
    assert (bounds[:, 0] >= 0).all() and (bounds[:, 1] >= 1).all() and (bounds[:, 0] + bounds[:, 1] <= n_in).all()
    assert (bounds[:, 1] <= ksize).all()
    # taps past the window are zero, windows move monotonically
    for i in range(0, n_out, max(1, n_out // 7)):
        assert (kk[i, bounds[i, 1]:] == 0).all()
    assert (np.diff(bounds[:, 0]) >= 0).all()


@settings(max_examples=25, deadline=None)
@given(w=st.integers(1, 90), h=st.integers(1, 90), ow=st.integers(1, 120), oh=st.integers(1, 120), seed=st.integers(0, 10 ** 6))
def test_oracle_resampler_equals_pillow_on_random_shapes(w, h, ow, oh, seed):
    img = np.random.RandomState(seed).randint(0, 256, (h, w, 3), dtype=np.uint8)
    want = np.asarray(Image.fromarray(img).resize((ow, oh), Image.Resampling.BICUBIC))
    assert np.array_equal(PR.resize_bicubic(img, ow, oh), want)


@settings(max_examples=200, deadline=None)
@given(n=st.integers(0, 10 ** 6), world=st.integers(1, 16))
def test_shard_ranges_partition_the_corpus(n, world):
    parts = [R.shard_range(n, r, world) for r in range(world)]
    assert parts[0][0] == 0 and parts[-1][1] == n
    assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
    sizes = [hi - lo for lo, hi in parts]
    assert max(sizes) - min(sizes) <= 1 and all(s >= 0 for s in sizes)


@settings(max_examples=60, deadline=None)
@given(nq=st.integers(1, 5), shards=st.integers(1, 4), per=st.integers(1, 12), k=st.integers(1, 8), seed=st.integers(0, 10 ** 6))
def test_merging_per_shard_topk_equals_global_topk(nq, shards, per, k, seed):
    """Top-k of the union of per-shard top-k lists == top-k of the whole corpus (tie rule: score desc, id asc), the
    property the one-all-gather multi-GPU retrieval relies on. Scores are drawn from a small set to force ties."""
    rs = np.random.RandomState(seed)
    S = rs.randint(0, 5, (nq, shards * per)).astype(np.float32)
    ids = np.arange(shards * per, dtype=np.int64)
    full_order = np.lexsort((np.broadcast_to(ids, S.shape), -S), axis=1)[:, :k]
    parts = []
    for s in range(shards):
        sl = slice(s * per, (s + 1) * per)
        order = np.lexsort((np.broadcast_to(ids[sl], S[:, sl].shape), -S[:, sl]), axis=1)[:, :k]
        parts.append((np.take_along_axis(S[:, sl], order, 1), ids[sl][order]))
    ms, mi = O.merge_topk(parts, min(k, shards * per))
    kk = min(k, shards * per)
    assert np.array_equal(mi[:, :kk], full_order[:, :kk]) and np.array_equal(ms[:, :kk], np.take_along_axis(S, full_order[:, :kk], 1))


@settings(max_examples=80, deadline=None)
@given(nd=st.integers(1, 30), nrel=st.integers(0, 6), k=st.integers(1, 12), seed=st.integers(0, 10 ** 6))
def test_metrics_equal_their_definitions(nd, nrel, k, seed):
    """recall.k / ndcg_cut.k / MRR@k against direct evaluation of their definitions on a random run with graded
    relevance, unjudged documents and score ties (trec_eval order: score desc, doc id desc)."""
    rs = np.random.RandomState(seed)
    docs = [f"d{i:02d}" for i in range(nd)]
    run = {"q": {d: float(rs.randint(0, 4)) for d in docs}}
    rel_docs = list(rs.choice(docs, size=min(nrel, nd), replace=False)) if nrel else []
    qrel = {"q": {d: int(rs.randint(1, 4)) for d in rel_docs}}
    if not qrel["q"]:
        qrel["q"]["unjudged-elsewhere"] = 0
    ranking = sorted(docs, key=lambda d: (run["q"][d], d), reverse=True)[:k]
    rel = {d for d, g in qrel["q"].items() if g > 0}
    want_recall = len(rel & set(ranking)) / len(rel) if rel else 0.0
    dcg = sum(max(qrel["q"].get(d, 0), 0) / math.log2(i + 2) for i, d in enumerate(ranking))
    ideal = sorted((g for g in qrel["q"].values() if g > 0), reverse=True)[:k]
    idcg = sum(g / math.log2(i + 2) for i, g in enumerate(ideal))
    assert I.recall_at_k(qrel, run, k)["all"] == want_recall
    assert abs(I.ndcg_at_k(qrel, run, k)["all"] - (dcg / idcg if idcg > 0 else 0.0)) < 1e-12
    # MRR uses Python's stable sort on score only (utils.py:285-308): first relevant doc in that order
    order = [d for d, _ in sorted(run["q"].items(), key=lambda x: x[1], reverse=True)][:k]
    rr = next((1.0 / (i + 1) for i, d in enumerate(order) if qrel["q"].get(d, 0) > 0), 0.0)
    assert I.eval_mrr(qrel, run, k)["all"] == rr
