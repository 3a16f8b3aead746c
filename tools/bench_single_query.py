"""Where one query's latency goes (full-size model): wall time of model(query=...) through the CUDA-graph path, GPU time
of the replayed graph, and the per-kernel-class times of the same step launched eagerly (ops.profile).
  python tools/bench_single_query.py [--reps 20]"""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from visrag_b200 import ops  # noqa: E402
from visrag_b200.config import VisRAGConfig  # noqa: E402
from visrag_b200.modeling import DRModelForInference, VisRAGRetB200  # noqa: E402
from visrag_b200.synth import synth_queries  # noqa: E402
from visrag_b200.tokenizer_stub import StubTokenizer  # noqa: E402
from visrag_b200.weights import random_state_dict_device  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=20)
    a = ap.parse_args()
    cfg = VisRAGConfig.full()
    tok = StubTokenizer(cfg.vocab)
    lm = VisRAGRetB200(cfg, random_state_dict_device(cfg, 2024, "cuda:0"), "cuda:0")
    model = DRModelForInference(lm_q=lm, pooling="wmean", normalize=True)
    qs = synth_queries(a.reps + 4, 7)
    q1 = lambda t: {"id": ["q"], "text": [t], "image": [None]}
    for t in qs[:4]:
        model(query=q1(t), tokenizer=tok, max_inp_length=2048).q_reps.cpu()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for t in qs[4:]:
        model(query=q1(t), tokenizer=tok, max_inp_length=2048).q_reps.cpu()
    e1.record()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / a.reps * 1e3
    # host preparation alone
    t0 = time.perf_counter()
    pbs = [model.prepare(q1(t), tokenizer=tok, max_inp_length=2048) for t in qs[4:]]
    prep = (time.perf_counter() - t0) / a.reps * 1e3
    # the same device step launched eagerly with per-launch events
    lm.engine.cuda_graphs = False
    model.encode_prepared(pbs[0])
    ops.profile_begin()
    model.encode_prepared(pbs[0])
    prof = ops.profile_end()
    total = sum(v[1] for v in prof.values())
    rec = {"what": "one text query through DRModelForInference (CUDA-graph path)", "lm_tokens": int(pbs[0].seq_lens[0]),
           "wall_ms_per_query_incl_d2h": round(wall, 3), "cuda_event_ms_per_query": round(e0.elapsed_time(e1) / a.reps, 3),
           "host_prepare_ms": round(prep, 3), "graph_stats": dict(lm.engine.graph_stats),
           "eager_kernel_ms_sum": round(total, 3),
           "eager_classes": {k: [v[0], round(v[1], 3)] for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])}}
    print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
