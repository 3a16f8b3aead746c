"""BASELINE configs[4] flavour on one GPU: mixed-resolution pages (sides in [336, 1344], 1-10 slices) through the pipelined
encode loop, host PIL front-end vs device front-end. Prints pages/s, slices/s and checks that both give identical
embeddings.   python tools/bench_mixed.py [--pages 256] [--batch 32]"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from visrag_b200 import inference as I  # noqa: E402
from visrag_b200.config import VisRAGConfig  # noqa: E402
from visrag_b200.host import plan_slices  # noqa: E402
from visrag_b200.modeling import DRModelForInference, VisRAGRetB200  # noqa: E402
from visrag_b200.tokenizer_stub import StubTokenizer  # noqa: E402
from visrag_b200.weights import random_state_dict_device  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pages", type=int, default=256)
    ap.add_argument("--batch", type=int, default=32)
    a = ap.parse_args()
    from PIL import Image

    cfg = VisRAGConfig.full()
    tok = StubTokenizer(cfg.vocab)
    lm = VisRAGRetB200(cfg, random_state_dict_device(cfg, 2024, "cuda:0"), "cuda:0")
    model = DRModelForInference(lm_q=lm, pooling="wmean", normalize=True)
    rs = np.random.RandomState(5)
    # a handful of distinct page formats (real corpora repeat a few scan sizes): log-uniform aspect in [1/3, 3], area in [336^2, 1344^2]
    formats = []
    for _ in range(8):
        area = rs.uniform(336 ** 2, 1344 ** 2)
        aspect = np.exp(rs.uniform(np.log(1 / 3), np.log(3)))
        w = int(np.clip(np.sqrt(area * aspect), 336, 1344))
        h = int(np.clip(np.sqrt(area / aspect), 336, 1344))
        formats.append((w, h))
    data, slices = [], 0
    for i in range(a.pages):
        w, h = formats[i % len(formats)]
        data.append({"id": f"p{i}", "text": "", "image": Image.fromarray(rs.randint(0, 256, (h, w, 3), dtype=np.uint8))})
        slices += plan_slices(w, h, cfg).n_slices
    kw = {"tokenizer": tok, "max_inp_length": 2048}
    print(f"{a.pages} pages, formats {formats}, {slices} slices ({slices / a.pages:.2f} per page), batch {a.batch}", file=sys.stderr, flush=True)
    res, line = {}, {"what": "BASELINE configs[4] flavour: mixed-resolution pages through inference.encode_stream, full-size model",
                     "pages": a.pages, "page_formats_wh": formats, "slices": slices, "slices_per_page": round(slices / a.pages, 2),
                     "pages_per_batch": a.batch, "slices_per_batch": round(slices / a.pages * a.batch, 1)}
    for name, dev in (("host PIL front-end", False), ("device front-end", True)):
        lm.engine.device_frontend = dev
        for _ in I.encode_stream(I._batches(data[: 2 * a.batch], a.batch), model, kw):
            pass
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = np.concatenate([arr for _, arr in I.encode_stream(I._batches(data, a.batch), model, kw)])
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        res[name] = out
        print(f"{name:20s}: {a.pages / dt:7.1f} pages/s  {slices / dt:7.1f} slices/s  ({dt * 1e3:.0f} ms)", file=sys.stderr, flush=True)
        line[name] = {"pages_per_s": round(a.pages / dt, 1), "slices_per_s": round(slices / dt, 1), "ms": round(dt * 1e3, 1)}
    same = np.array_equal(res["host PIL front-end"], res["device front-end"])
    line["embeddings_identical_between_front_ends"] = bool(same)
    line["graph_stats"] = dict(lm.engine.graph_stats)
    print(json.dumps(line), flush=True)
    return 0 if same else 1


if __name__ == "__main__":
    sys.exit(main())
