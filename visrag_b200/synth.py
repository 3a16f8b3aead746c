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
