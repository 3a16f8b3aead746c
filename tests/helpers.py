"""Shared test helpers: synthetic pages/queries (same generators the golden script used) and golden loading."""
import json
import os

import numpy as np
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
from visrag_b200.synth import QUERY_PREFIX, synth_pages  # noqa: E402,F401


def load_case(name):
    from visrag_b200.config import VisRAGConfig

    z = np.load(os.path.join(GOLDEN, f"{name}.npz"), allow_pickle=False)
    cfg = VisRAGConfig(**json.loads(str(z["config"])))
    pages = synth_pages(z["page_sizes"], int(z["page_seed"]))
    queries = [str(q) for q in z["queries"]]
    return cfg, int(z["weight_seed"]), pages, queries, z


def cosine_rows(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return (a * b).sum(1) / (np.linalg.norm(a, axis=1) * np.linalg.norm(b, axis=1))
