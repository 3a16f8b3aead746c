"""Host-side logic (geometry, placeholders, tokenisation, packing) against the reference's golden geometry and
against the oracle restatement. CPU only."""
import numpy as np
import pytest
from PIL import Image

from oracle import restated as O
from tests.helpers import GOLDEN, synth_pages
from visrag_b200 import host
from visrag_b200.config import VisRAGConfig
from visrag_b200.tokenizer_stub import StubTokenizer

CFG = VisRAGConfig.tiny()


def test_plan_matches_reference_golden_geometry():
    z = np.load(f"{GOLDEN}/geometry_v1.npz")
    for W, H, sw, sh, gx, gy, pw, ph, npatch in z["cases"]:
        plan = host.plan_slices(int(W), int(H), CFG)
        assert plan.source_size == (sw, sh), (W, H)
        if gx == 0:
            assert plan.grid is None and npatch == 0
        else:
            assert plan.grid == (gx, gy) and plan.cell_size == (pw, ph) and plan.n_slices == 1 + npatch, (W, H)


def test_known_geometries():
    # 564x3040 page -> thumbnail 196x1036, grid [1,8], eight 546x364 slices
    p = host.plan_slices(564, 3040, CFG)
    assert p.source_size == (196, 1036) and p.grid == (1, 8) and p.cell_size == (546, 364)
    for s in (224, 448):
        p = host.plan_slices(s, s, CFG)
        assert p.source_size == (448, 448) and p.grid is None


@pytest.mark.parametrize("size", [(224, 224), (700, 900), (760, 141), (1200, 500), (449, 449)])
def test_render_is_pixel_exact_vs_oracle(size):
    img = synth_pages([size], 3)[0]
    src, patches, grid = O.slice_image(img, 9, 448, 14)
    want = [np.asarray(src)] + [np.asarray(p) for row in patches for p in row]
    got = host.render_slices(img, host.plan_slices(*img.size, CFG))
    assert len(got) == len(want)
    for g, w in zip(got, want):
        assert g.dtype == np.uint8 and g.shape == w.shape and np.array_equal(g, w)


def test_tokenisation_and_packing_match_oracle():
    tok = StubTokenizer(CFG.vocab)
    imgs = synth_pages([(224, 224), (700, 900)], 5) + [None]
    texts = ["", "caption text", "Represent this query for retrieving relevant documents: hello"]
    pb = host.prepare_batch(texts, imgs, tok, CFG, 2048)
    assert pb.n_items == 3 and pb.cu_seqlens[-1] == pb.token_src.shape[0] == pb.positions.shape[0]
    slice_rows = []
    for b, (t, im) in enumerate(zip(texts, imgs)):
        content, slices = O.prepare_context(t, im, tok, CFG.query_num)
        ids, bound = O.convert_to_tensors(tok, content, 2048)
        lo, hi = pb.cu_seqlens[b], pb.cu_seqlens[b + 1]
        assert hi - lo == len(ids)
        src = pb.token_src[lo:hi]
        assert np.array_equal(pb.positions[lo:hi], np.arange(len(ids)))
        is_img = np.zeros(len(ids), bool)
        for (s, e) in bound:
            is_img[s:e] = True
        assert np.array_equal(-(src[~is_img] + 1), ids[~is_img])
        assert (src[is_img] >= 0).all()
        for n, (s, e) in enumerate(bound):
            rows = src[s:e]
            assert np.array_equal(rows - rows[0], np.arange(64)) and rows[0] % 64 == 0
            slice_rows.append((rows[0] // 64, np.asarray(slices[n])))
    # every slice index maps to the right pixels inside its geometry group
    for idx, px in slice_rows:
        key = (px.shape[0], px.shape[1])
        j = idx - pb.group_row0[key]
        assert np.array_equal(pb.groups[key][j], px)
    assert pb.n_slices == len(slice_rows) == 1 + 5  # 224^2 page: 1 slice; 700x900: thumbnail + 2x2


def test_text_only_and_errors():
    tok = StubTokenizer(CFG.vocab)
    pb = host.prepare_batch(["a", "bcd"], [None, None], tok, CFG, 3)
    assert list(pb.seq_lens) == [2, 3] and pb.n_slices == 0 and not pb.groups
    with pytest.raises(ValueError):
        host.prepare_batch([""], [Image.new("RGB", (224, 224))], tok, CFG, 40)  # span cut by truncation
    with pytest.raises(NotImplementedError):
        host.prepare_batch([["chat"]], [None], tok, CFG)
