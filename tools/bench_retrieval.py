"""Retrieval at BASELINE configs[3] scale on ONE GPU's shard (125 k pages x 10 k queries, top-10) and single-query
latency over a 1 M-page demo index. Prints one JSON line per measurement.
  python tools/bench_retrieval.py [--corpus 125000] [--queries 10000] [--big 1000000]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from visrag_b200 import retriever as R  # noqa: E402


def unit(n, d, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    out = torch.empty((n, d), dtype=torch.float32, device="cuda")
    for r0 in range(0, n, 65536):  # chunked: randn + normalise without a second full-size temporary
        x = torch.randn((min(65536, n - r0), d), device="cuda", generator=g)
        out[r0:r0 + x.shape[0]] = torch.nn.functional.normalize(x, dim=1)
    return out


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--corpus", type=int, default=125000)
    ap.add_argument("--queries", type=int, default=10000)
    ap.add_argument("--big", type=int, default=1000000)
    ap.add_argument("--dim", type=int, default=2304)
    ap.add_argument("--k", type=int, default=10)
    a = ap.parse_args()
    peaks = {}
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peaks = json.load(open(p))
    tf_peak = peaks.get("bf16_tflops_sustained") or peaks.get("bf16_tflops") or 1429.0
    hbm_peak = peaks.get("hbm_gbs") or 6500.0

    D = unit(a.corpus, a.dim, 1)
    Q = unit(a.queries, a.dim, 2)
    index = R.build_index(D)
    stats = {}
    ms, (s, i) = timed(lambda: R.score_topk(Q, index, a.k, stats=stats))
    sub = torch.randperm(a.queries, device="cuda")[:64]
    s_ref, i_ref = R.score_topk(Q[sub], index, a.k, force_exact=True)
    same = bool(torch.equal(i[sub], i_ref)) and float((s[sub] - s_ref).abs().max()) <= 2e-6
    tf = 2.0 * a.queries * a.corpus * a.dim / (ms / 1e3) / 1e12
    print(json.dumps({"what": "score+top-k, one shard", "queries": a.queries, "corpus": a.corpus, "k": a.k, "ms": round(ms, 3),
                      "queries_per_s": round(a.queries / (ms / 1e3), 1), "tflops_fp16_filter": round(tf, 1),
                      "frac_of_tensor_peak": round(tf / tf_peak, 3), "path": stats.get("path"), "flagged": stats.get("flagged"),
                      "equals_fp32_scan_on_64_queries": same}), flush=True)
    del D, Q, index, s, i
    torch.cuda.empty_cache()

    D = unit(a.big, a.dim, 3)
    index = R.build_index(D)
    for nq in (1, 8, 64):
        Q = unit(nq, a.dim, 10 + nq)
        ms, (s, i) = timed(lambda: R.score_topk(Q, index, a.k), reps=5)
        ref = torch.topk(Q @ D.T, a.k, dim=1)
        ok = bool(torch.equal(i, ref.indices)) or float((s - ref.values).abs().max()) <= 2e-6
        gbs = a.big * a.dim * 4 / (ms / 1e3) / 1e9
        print(json.dumps({"what": "few queries over a resident index (fp32 scan + chunked top-k)", "queries": nq, "corpus": a.big,
                          "ms": round(ms, 3), "index_GBps": round(gbs, 1), "frac_of_hbm_peak": round(gbs / hbm_peak, 3),
                          "matches_torch_fp32": ok}), flush=True)


if __name__ == "__main__":
    main()
