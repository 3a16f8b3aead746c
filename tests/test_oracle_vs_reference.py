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
