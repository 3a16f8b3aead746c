"""The oracle restatement against golden vectors produced by the REAL reference (oracle/gen_golden.py). CPU only."""
import numpy as np
import pytest

from oracle import restated as O
from tests.helpers import GOLDEN, cosine_rows, load_case
from visrag_b200.tokenizer_stub import StubTokenizer
from visrag_b200.weights import random_state_dict


@pytest.fixture(scope="module")
def tiny():
    cfg, wseed, pages, queries, z = load_case("tiny_v1")
    return cfg, random_state_dict(cfg, wseed), pages, queries, z


def test_oracle_reproduces_reference_embeddings(tiny):
    cfg, sd, pages, queries, z = tiny
    tok = StubTokenizer(cfg.vocab)
    p = O.encode(sd, cfg, tok, [""] * len(pages), pages)
    q = O.encode(sd, cfg, tok, queries, [None] * len(queries))
    assert p.dtype == np.float32 and p.shape == z["page_reps"].shape
    # fp32 vs fp32: only summation order differs (reference batches/pads, oracle goes sequence by sequence)
    assert np.abs(p - z["page_reps"]).max() < 2e-6
    assert np.abs(q - z["query_reps"]).max() < 2e-6
    assert np.allclose(np.linalg.norm(p, axis=1), 1.0, atol=1e-6)


def test_oracle_score_topk_matches_reference_topk(tiny):
    _, _, _, _, z = tiny
    s, i = O.score_topk(z["query_reps"], z["page_reps"], z["topk_indices"].shape[1])
    assert np.array_equal(i, z["topk_indices"])
    assert np.allclose(s, z["topk_scores"], atol=1e-6)


def test_oracle_geometry_matches_reference():
    z = np.load(f"{GOLDEN}/geometry_v1.npz")
    from PIL import Image

    for W, H, sw, sh, gx, gy, pw, ph, npatch in z["cases"][:60]:
        src, patches, grid = O.slice_image(Image.new("RGB", (int(W), int(H))), 9, 448, 14)
        assert src.size == (sw, sh)
        assert (grid or [0, 0]) == [gx, gy]
        assert sum(len(r) for r in patches) == npatch
        if patches:
            assert patches[0][0].size == (pw, ph)


def test_pooling_variants_and_edge_cases():
    import torch

    h = torch.arange(12, dtype=torch.float32).reshape(3, 4)
    assert torch.allclose(O.pool(h, "wmean"), (h[0] + 2 * h[1] + 3 * h[2]) / 6)
    assert torch.allclose(O.pool(h, "mean"), h.mean(0))
    assert torch.equal(O.pool(h, "lasttoken"), h[2]) and torch.equal(O.pool(h, "cls"), h[0])
    with pytest.raises(ValueError):
        O.pool(h, "nope")


def test_score_topk_ties_short_corpus_and_merge():
    Q = np.eye(3, 8, dtype=np.float32)
    D = np.concatenate([np.eye(3, 8, dtype=np.float32)] * 2)  # docs 0..2 duplicated as 3..5 -> exact ties
    s, i = O.score_topk(Q, D, 4)
    assert [list(r[:2]) for r in i] == [[0, 3], [1, 4], [2, 5]]  # ties: lower id first
    s2, i2 = O.score_topk(Q, D[:2], 5)                          # k > nd
    assert s2.shape == (3, 2)
    # sharded == global
    rs = np.random.RandomState(0)
    Q = rs.randn(7, 16).astype(np.float32)
    D = rs.randn(50, 16).astype(np.float32)
    gs, gi = O.score_topk(Q, D, 5)
    parts = []
    for lo, hi in ((0, 17), (17, 34), (34, 50)):
        ps, pi = O.score_topk(Q, D[lo:hi], 5)
        parts.append((ps, pi + lo))
    ms, mi = O.merge_topk(parts, 5)
    assert np.array_equal(mi, gi) and np.allclose(ms, gs)
    rel = [set(gi[q, :2].tolist()) for q in range(7)]
    assert O.recall_at_k(gi, rel, 5) == 1.0 and O.recall_at_k(gi[:, ::-1], rel, 1) < 1.0


def test_oracle_reproduces_reference_ranking_on_the_document_corpus():
    """tiny_v2: 36 pages (structured synthetic documents, noise pages and the reference's four real example images) x 10
    queries through the REAL reference; the oracle must give the same embeddings and the same ORDERED top-5 of 36."""
    cfg, wseed, pages, queries, z = load_case("tiny_v2")
    assert len(pages) == 36 and len(queries) == 10 and z["topk_indices"].shape == (10, 5)
    sd = random_state_dict(cfg, wseed)
    tok = StubTokenizer(cfg.vocab)
    p = O.encode(sd, cfg, tok, [""] * len(pages), pages)
    q = O.encode(sd, cfg, tok, queries, [None] * len(queries))
    assert np.abs(p - z["page_reps"]).max() < 2e-6 and np.abs(q - z["query_reps"]).max() < 2e-6
    s, i = O.score_topk(q, p, 5)
    assert np.array_equal(i, z["topk_indices"])
    # the full-size golden shares the page spec and the queries (only the weights differ)
    zf = np.load(f"{GOLDEN}/full_v2.npz")
    assert str(zf["page_spec"]) == str(z["page_spec"]) and list(zf["queries"]) == list(z["queries"])
    assert zf["page_reps"].shape == (36, 2304) and zf["topk_indices"].shape == (10, 5)
