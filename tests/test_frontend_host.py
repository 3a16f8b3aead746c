"""Host side of the device image front-end, no GPU:
  * the oracle restatement of Pillow's 8-bit bicubic resampler is pinned against PIL itself (bit for bit),
  * the product's coefficient tables equal the oracle's,
  * prepare_batch(device_frontend=True) lays slices out exactly like the host path: executing its page jobs with the
    oracle resampler reproduces every slice the PIL path renders, at the same group / index."""
import numpy as np
import pytest
from PIL import Image

from oracle import pil_resample as PR
from tests.helpers import synth_pages
from visrag_b200 import frontend as F
from visrag_b200.config import VisRAGConfig
from visrag_b200.host import prepare_batch
from visrag_b200.tokenizer_stub import StubTokenizer

SHAPES = [(64, 48, 32, 32), (100, 75, 448, 336), (640, 480, 448, 336), (517, 301, 518, 301), (301, 517, 301, 520),
          (33, 400, 14, 434), (448, 448, 448, 448), (900, 700, 504, 392), (1344, 336, 896, 224), (50, 50, 700, 14),
          (1000, 30, 28, 28), (1, 1, 14, 14), (3, 2, 5, 7)]


@pytest.mark.parametrize("w,h,ow,oh", SHAPES)
def test_oracle_resampler_is_pillow_bit_for_bit(w, h, ow, oh):
    img = np.random.RandomState(w * 7 + h).randint(0, 256, (h, w, 3), dtype=np.uint8)
    want = np.asarray(Image.fromarray(img).resize((ow, oh), Image.Resampling.BICUBIC))
    assert np.array_equal(PR.resize_bicubic(img, ow, oh), want)


def test_extreme_pixels_saturate_like_pillow():
    """Bicubic overshoot on hard edges exercises clip8 on both sides."""
    img = np.zeros((40, 60, 3), dtype=np.uint8)
    img[:, 30:] = 255
    img[10:20] = 255
    for ow, oh in [(90, 70), (23, 17)]:
        want = np.asarray(Image.fromarray(img).resize((ow, oh), Image.Resampling.BICUBIC))
        assert np.array_equal(PR.resize_bicubic(img, ow, oh), want)


@pytest.mark.parametrize("n_in,n_out", [(448, 448), (1700, 448), (100, 448), (517, 518), (2200, 350), (7, 1000), (5000, 14)])
def test_product_tables_equal_oracle_tables(n_in, n_out):
    ksize, bounds, kk = F.resample_coeffs(n_in, n_out)
    oks, ob, okk = PR.precompute_coeffs(n_in, n_out)
    assert ksize == oks and np.array_equal(bounds, np.asarray(ob, dtype=np.int32))
    assert np.array_equal(kk, np.asarray(PR.normalize_coeffs_8bpc(okk), dtype=np.int32))
    assert kk.dtype == np.int32 and bounds.dtype == np.int32
    # every window stays inside the source axis and every row sums to (about) one in fixed point
    assert (bounds[:, 0] >= 0).all() and (bounds[:, 0] + bounds[:, 1] <= n_in).all()
    assert np.abs(kk.sum(1) - (1 << F.PRECISION_BITS)).max() <= ksize


def test_device_jobs_reproduce_the_host_layout():
    cfg = VisRAGConfig.tiny()
    tok = StubTokenizer(cfg.vocab)
    pages = synth_pages([(448, 448), (700, 900), (224, 224), (1100, 500), (700, 900), (640, 320)], 11)
    texts = [""] * len(pages) + ["a text query"]
    images = pages + [None]
    host = prepare_batch(texts, images, tok, cfg, 2048)
    dev = prepare_batch(texts, images, tok, cfg, 2048, device_frontend=True)
    for f in ("seq_lens", "cu_seqlens", "positions", "token_src"):
        assert np.array_equal(getattr(host, f), getattr(dev, f)), f
    assert host.group_row0 == dev.group_row0 and host.n_slices == dev.n_slices
    assert {k: len(v) for k, v in host.groups.items()} == {k: len(v) for k, v in dev.groups.items()}
    # every page is a device job (the 448x448 one is a plain copy there); pixels travel as Pillow's RGBX rows
    assert len(dev.jobs) == len(pages) and all(a is None for lst in dev.groups.values() for a in lst)
    assert all(j.pixels.dtype == np.uint8 and j.pixels.shape[2] in (3, 4) for j in dev.jobs)
    assert all(np.array_equal(j.pixels[..., :3], np.asarray(p)) for j, p in zip(dev.jobs, pages))
    # execute the jobs with the oracle resampler and compare slice by slice with the PIL-rendered groups
    filled = {k: [None] * len(v) for k, v in dev.groups.items()}
    for j in dev.jobs:
        p = j.plan
        key, idx = j.thumb
        rgb = np.ascontiguousarray(j.pixels[..., :3])
        filled[key][idx] = PR.resize_bicubic(rgb, *p.source_size)
        if p.grid is not None:
            ref = PR.resize_bicubic(rgb, *p.refine_size)
            cw, ch = p.cell_size
            key, idx = j.cells
            for cy in range(p.grid[1]):
                for cx in range(p.grid[0]):
                    filled[key][idx + cy * p.grid[0] + cx] = ref[cy * ch:(cy + 1) * ch, cx * cw:(cx + 1) * cw]
    for k, lst in host.groups.items():
        for i, a in enumerate(lst):
            got = filled[k][i] if dev.groups[k][i] is None else dev.groups[k][i]
            assert np.array_equal(a, got), (k, i)


def test_truncated_page_falls_back_to_host_rendering():
    """max_inp_length cuts the placeholder text short: fewer <image> spans than slices -> that page is rendered on the
    host (the reference consumes the first len(image_bound) slices), the others stay device jobs."""
    cfg = VisRAGConfig.tiny()
    tok = StubTokenizer(cfg.vocab)
    pages = synth_pages([(700, 900), (300, 200)], 3)
    full = prepare_batch(["", ""], pages, tok, cfg, 2048, device_frontend=True)
    cut = host_cut = None
    for limit in range(70, 400):  # a cut in the middle of a span raises (as the reference would); find a clean one
        try:
            host_cut = prepare_batch(["", ""], pages, tok, cfg, limit)
        except ValueError:
            continue
        if host_cut.n_slices < full.n_slices:
            cut = prepare_batch(["", ""], pages, tok, cfg, limit, device_frontend=True)
            break
    assert cut is not None
    assert len(full.jobs) == 2 and len(cut.jobs) == 1
    assert cut.n_slices == host_cut.n_slices and np.array_equal(cut.token_src, host_cut.token_src)
    for k, lst in host_cut.groups.items():
        for a, b in zip(lst, cut.groups[k]):
            assert b is None or np.array_equal(a, b)


def test_page_pixels_exports_pillow_rows_without_repacking():
    """RGBX zero-copy export when Pillow + pyarrow provide it, RGB otherwise; other modes are converted first."""
    rs = np.random.RandomState(1)
    arr = rs.randint(0, 256, (37, 53, 3), dtype=np.uint8)
    for im in (Image.fromarray(arr), Image.fromarray(arr).crop((3, 2, 40, 30)), Image.fromarray(arr[..., 0]), Image.fromarray(arr).convert("RGBA")):
        px = F.page_pixels(im)
        assert px.dtype == np.uint8 and px.shape[:2] == (im.size[1], im.size[0]) and px.shape[2] in (3, 4)
        assert np.array_equal(px[..., :3], np.asarray(im.convert("RGB")))


def test_page_pixels_on_decoded_files_and_multi_block_images():
    """Lazily decoded PNG/JPEG files export the same way; images large enough for Pillow to allocate them in several
    blocks cannot be exported as one buffer and must fall back to the packed RGB copy - with the same pixels."""
    import io

    rs = np.random.RandomState(2)
    arr = rs.randint(0, 256, (120, 200, 3), dtype=np.uint8)
    for fmt in ("PNG", "JPEG"):
        buf = io.BytesIO()
        Image.fromarray(arr).save(buf, format=fmt)
        buf.seek(0)
        im = Image.open(buf)
        px = F.page_pixels(im)
        assert np.array_equal(px[..., :3], np.asarray(Image.open(io.BytesIO(buf.getvalue())).convert("RGB")))
    big = rs.randint(0, 256, (4200, 3300, 3), dtype=np.uint8)
    px = F.page_pixels(Image.fromarray(big))
    assert px.shape[:2] == (4200, 3300) and np.array_equal(px[..., :3], big)
