"""Pins the oracle against the REAL reference executed in this container (skipped where /root/reference is absent,
e.g. on the GPU box — the committed goldens carry the same information there)."""
import numpy as np
import pytest

from oracle import reference_shim as RS

pytestmark = [pytest.mark.reference, pytest.mark.skipif(not RS.available(), reason="/root/reference not mounted")]


def test_restatement_equals_reference_on_fresh_inputs():
    from oracle import restated as O
    from tests.helpers import QUERY_PREFIX, synth_pages
    from visrag_b200.config import VisRAGConfig
    from visrag_b200.tokenizer_stub import StubTokenizer
    from visrag_b200.weights import random_state_dict

    cfg = VisRAGConfig.tiny()
    sd = random_state_dict(cfg, 777)
    model = RS.build_reference_model(cfg, sd, attn_implementation="sdpa")
    tok = StubTokenizer(cfg.vocab)
    pages = synth_pages([(300, 300), (1000, 600), (448, 448)], 21)
    items = [{"id": str(i), "text": "doc text" if i == 1 else "", "image": im} for i, im in enumerate(pages)]
    p_ref = RS.encode(model, tok, items, False)
    p = O.encode(sd, cfg, tok, [it["text"] for it in items], pages)
    assert np.abs(p - p_ref).max() < 2e-6
    qs = [QUERY_PREFIX + "what is shown", QUERY_PREFIX + "x"]
    q_ref = RS.encode(model, tok, [{"id": f"q{i}", "text": t, "image": None} for i, t in enumerate(qs)], True)
    assert np.abs(O.encode(sd, cfg, tok, qs, [None, None]) - q_ref).max() < 2e-6
    # B1 boundary: hidden states of the valid positions
    hs, mask = RS.hidden_states(model, tok, [it["text"] for it in items], pages)
    _, hid = O.encode(sd, cfg, tok, [it["text"] for it in items], pages, return_hidden=True)
    for b, h in enumerate(hid):
        n = int(mask[b].sum())
        assert n == h.shape[0] and np.abs(hs[b, :n] - h).max() < 5e-5


@pytest.mark.parametrize("pooling", ["lasttoken", "mean", "cls"])
def test_other_poolings_equal_reference_on_a_ragged_batch(pooling):
    """the pooling variants of `dense_retrieval_model.py:170-218` on a right-padded batch of unequal
    lengths (the oracle and the engine pool unpadded sequences; this pins that they mean the same thing)."""
    from oracle import restated as O
    from tests.helpers import QUERY_PREFIX, synth_pages
    from visrag_b200.config import VisRAGConfig
    from visrag_b200.tokenizer_stub import StubTokenizer
    from visrag_b200.weights import random_state_dict

    cfg = VisRAGConfig.tiny()
    sd = random_state_dict(cfg, 778)
    model = RS.build_reference_model(cfg, sd, attn_implementation="sdpa", pooling=pooling)
    tok = StubTokenizer(cfg.vocab)
    page = synth_pages([(448, 448)], 5)[0]
    texts = [QUERY_PREFIX + "a", QUERY_PREFIX + "a much longer query about the page content", ""]
    images = [None, None, page]
    items = [{"id": str(i), "text": t, "image": im} for i, (t, im) in enumerate(zip(texts, images))]
    ref = RS.encode(model, tok, items, False)
    got = O.encode(sd, cfg, tok, texts, images, pooling=pooling)
    assert np.abs(got - ref).max() < 2e-6, pooling


def test_score_topk_and_run_files_equal_reference(tmp_path):
    """The scoring side of the path against the REAL reference functions on the CPU: `_retrieve_one_shard`
    (`retriever/dense_retriever.py:13-34`) on a pickle shard, `save_as_trec` / `load_from_trec` / `eval_mrr`
    (`utils.py:125-175,285-308`). Random unit vectors: no score ties, so torch.topk's unspecified tie order cannot differ."""
    import pickle

    import torch

    RS._import_reference()
    from openmatch import utils as ref_utils
    from openmatch.retriever.dense_retriever import _retrieve_one_shard as ref_retrieve

    from oracle import restated as O
    from visrag_b200 import inference as I
    from visrag_b200 import retriever as R

    rs = np.random.RandomState(12)
    D = rs.randn(500, 64).astype(np.float32)
    D /= np.linalg.norm(D, axis=1, keepdims=True)
    Q = rs.randn(7, 64).astype(np.float32)
    Q /= np.linalg.norm(Q, axis=1, keepdims=True)
    lookup = [f"doc{i}" for i in range(len(D))]
    shard = str(tmp_path / "embeddings.corpus.rank.0")
    R.save_shard(shard, D, lookup)                      # our writer, the reference's reader
    assert pickle.load(open(shard, "rb"))[1] == lookup
    s_ref, i_ref, look_ref = ref_retrieve(shard, torch.from_numpy(Q), 10, "cpu")
    s, i = O.score_topk(Q, D, 10)
    assert look_ref == lookup and np.array_equal(i, i_ref.numpy()) and np.abs(s - s_ref.numpy()).max() < 1e-6
    # run files and MRR: our functions and the reference's read each other's output and agree
    run = {f"q{q}": {lookup[j]: float(s[q, r]) for r, j in enumerate(i[q])} for q in range(len(Q))}
    qrel = {f"q{q}": {lookup[int(i[q, q % 10])]: 1} for q in range(len(Q))}
    ours, theirs = str(tmp_path / "ours.trec"), str(tmp_path / "theirs.trec")
    I.save_as_trec(run, ours)
    ref_utils.save_as_trec(run, theirs)
    assert open(ours).read() == open(theirs).read()
    assert ref_utils.load_from_trec(ours) == I.load_from_trec(theirs)
    assert ref_utils.eval_mrr(qrel, run, 10) == I.eval_mrr(qrel, run, 10)
    assert ref_utils.eval_mrr(qrel, run, 3) == I.eval_mrr(qrel, run, 3)
