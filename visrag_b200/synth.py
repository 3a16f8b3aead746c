"""Deterministic synthetic inputs (there is no network for datasets): noise pages and keyword queries.
numpy's legacy RandomState stream is stable across versions, so tests, goldens and benchmarks re-draw identical data."""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np

QUERY_PREFIX = "Represent this query for retrieving relevant documents: "  # reference eval.sh:45

_VOCAB = ["revenue", "table", "figure", "growth", "policy", "network", "energy", "chart", "model", "summary",
          "annual", "report", "risk", "market", "climate", "protein", "budget", "survey", "method", "result"]


def synth_pages(sizes: Sequence[Tuple[int, int]], seed: int):
    """uint8 RGB noise pages as PIL images; sizes are (width, height)."""
    from PIL import Image

    rs = np.random.RandomState(seed)
    return [Image.fromarray(rs.randint(0, 256, (int(h), int(w), 3), dtype=np.uint8)) for (w, h) in sizes]


def synth_queries(n: int, seed: int) -> List[str]:
    rs = np.random.RandomState(seed)
    return [QUERY_PREFIX + " ".join(_VOCAB[j] for j in rs.randint(0, len(_VOCAB), rs.randint(3, 12))) for _ in range(n)]


def synth_doc_pages(sizes: Sequence[Tuple[int, int]], seed: int):
    """Structured synthetic "document" pages as PIL RGB images; sizes are (width, height).

    i.i.d. noise pages of one size embed almost identically (the ViT averages the noise away), which makes ranking
    parity vacuous. These pages differ at low spatial frequencies the way scanned documents do: a tinted paper colour,
    a header bar, coloured figure blocks and rows of dark "glyph" runs of random widths, plus mild per-pixel noise.
    Pure integer drawing on a numpy canvas (no fonts, no PIL drawing code paths), so the pixels only depend on numpy's
    legacy RandomState stream."""
    from PIL import Image

    rs = np.random.RandomState(seed)
    pages = []
    for (w, h) in sizes:
        w, h = int(w), int(h)
        paper = rs.randint(200, 256, 3)
        img = np.empty((h, w, 3), dtype=np.int16)
        img[:] = paper
        # header bar
        hb = max(4, h // rs.randint(8, 20))
        img[:hb] = rs.randint(0, 200, 3)
        # figure blocks
        for _ in range(rs.randint(1, 5)):
            bw, bh = rs.randint(w // 8, w // 2 + 1), rs.randint(h // 10, h // 3 + 1)
            x0, y0 = rs.randint(0, w - bw + 1), rs.randint(hb, max(hb + 1, h - bh + 1))
            col = rs.randint(0, 256, 3)
            img[y0:y0 + bh, x0:x0 + bw] = col
            if rs.randint(2):  # striped fill (a "chart")
                step = rs.randint(6, 24)
                img[y0:y0 + bh:step, x0:x0 + bw] = 255 - col
        # text rows: dark runs of random widths on a line grid
        line_h = rs.randint(10, 28)
        margin = rs.randint(4, max(5, w // 10))
        ink = rs.randint(0, 90, 3)
        for y in range(hb + line_h, h - line_h, line_h * 2):
            if rs.rand() < 0.25:
                continue
            x = margin
            end = w - margin - rs.randint(0, w // 3 + 1)
            while x < end:
                run = rs.randint(3, 40)
                img[y:y + line_h - 3, x:min(x + run, end)] = ink
                x += run + rs.randint(3, 12)
        img += rs.randint(-6, 7, img.shape).astype(np.int16)
        pages.append(Image.fromarray(np.clip(img, 0, 255).astype(np.uint8)))
    return pages
