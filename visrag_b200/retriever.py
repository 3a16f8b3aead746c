"""Dense retrieval on the B200: drop-in for `src/openmatch/retriever/dense_retriever.py`.

Same call signatures and on-disk formats as the reference:
  * `_retrieve_one_shard(corpus_shard_path, encoded_queries_tensor, topk, device)` -> (scores, indices, lookup)
    (`dense_retriever.py:13-34`), shard file = `pickle((float32[n, d], List[str]))` (`inference/inference.py:126`);
  * `distributed_parallel_retrieve(args, topk)` -> {qid: {docid: score}} (`dense_retriever.py:37-97`).
Underneath, `torch.matmul` + `torch.topk` are replaced by the fused tcgen05 filter + exact fp32 rescoring kernels
(csrc/score.cu): the returned scores are fp32 dot products and the top-k equals the fp32 scan's
(order: score descending, then doc index ascending — `torch.topk` leaves tie order unspecified).

For a corpus that lives in HBM (index build + many query batches) use `CorpusIndex` / `score_topk` directly; the
multi-GPU form (`sharded_topk`) shards the corpus by page across ranks, takes the local top-k with global ids and
merges after ONE all-gather of [nq, k] (score, id) pairs.
"""
from __future__ import annotations

import glob
import os
import pickle
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from . import _lib as L

SMALL_PROBLEM = 1 << 22  # nq*nd below this: plain fp32 scan (the GEMM pipeline would not even fill)


@dataclass
class CorpusIndex:
    """One corpus shard resident in HBM: fp32 embeddings (exact rescoring), fp16 copy (tensor-core filter)."""
    emb: torch.Tensor          # [nd, d] fp32
    emb_f16: torch.Tensor      # [nd, d] fp16
    max_norm: torch.Tensor     # [1] fp32: max row L2 norm (error bound of the filter)
    lookup: Optional[List[str]] = None

    @property
    def nd(self) -> int:
        return self.emb.shape[0]


def _check_f32(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise ValueError(f"{name} must be a CUDA tensor (visrag_b200 has no CPU path)")
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def to_f16_rows(x: torch.Tensor, want_max_norm: bool = False):
    """fp32 [n,d] -> fp16 copy (+ max row norm) with the library kernel."""
    n, d = x.shape
    out = torch.empty((n, d), dtype=torch.float16, device=x.device)
    mx = torch.zeros(1, dtype=torch.float32, device=x.device) if want_max_norm else None
    L.check(L.lib().vr_f32_to_f16_rows(x.data_ptr(), n, d, out.data_ptr(), None, L.ptr(mx), L.stream_ptr()))
    return (out, mx) if want_max_norm else out


def build_index(emb, lookup: Optional[List[str]] = None, device: str = "cuda") -> CorpusIndex:
    """numpy / torch fp32 [nd, d] -> device-resident index (a tensor stays on the device it lives on).
    The tensor-core filter's exactness proof needs finite fp16 copies of Q and D (|x| < 65504). The L2-normalised
    embeddings of the encoder always are; for anything else the rescoring kernel sees a query or max document row norm
    that is not < 65504 (inf / NaN included), flags the query, and `score_topk` reruns it through the plain fp32 scan."""
    if isinstance(emb, np.ndarray):
        emb = torch.from_numpy(np.ascontiguousarray(emb, dtype=np.float32)).to(L.norm_device(device))
    emb = _check_f32(emb, "emb")
    if emb.shape[1] % 8 != 0:
        raise ValueError("embedding dim must be a multiple of 8")
    with L.on_device(emb.device):
        f16, mx = to_f16_rows(emb, want_max_norm=True)
    return CorpusIndex(emb, f16, mx, lookup)


def _exact_topk(q: torch.Tensor, index: CorpusIndex, k: int, id_offset: int):
    nq, d = q.shape
    nd = index.nd
    out_s = torch.empty((nq, k), dtype=torch.float32, device=q.device)
    out_i = torch.empty((nq, k), dtype=torch.int64, device=q.device)
    rows_per = max(1, min(nq, (1 << 28) // max(nd, 1)))  # <= 1 GiB of fp32 scratch
    scratch = torch.empty((rows_per, nd), dtype=torch.float32, device=q.device)
    lib = L.lib()
    for r0 in range(0, nq, rows_per):
        n = min(rows_per, nq - r0)
        L.check(lib.vr_score_exact(q[r0:].data_ptr(), n, index.emb.data_ptr(), nd, d, scratch.data_ptr(), L.stream_ptr()))
        chunks = min(1024, nd // 4096) if n <= 64 else 0  # few queries over a long index: spread each row over many SMs
        if chunks >= 2:
            ws_s = torch.empty((n, chunks, k), dtype=torch.float32, device=q.device)
            ws_i = torch.empty((n, chunks, k), dtype=torch.int64, device=q.device)
            L.check(lib.vr_topk_rows_chunked(scratch.data_ptr(), n, nd, k, id_offset, chunks, ws_s.data_ptr(), ws_i.data_ptr(),
                                             out_s[r0:].data_ptr(), out_i[r0:].data_ptr(), L.stream_ptr()))
        else:
            L.check(lib.vr_topk_rows(scratch.data_ptr(), None, n, nd, k, id_offset, out_s[r0:].data_ptr(),
                                     out_i[r0:].data_ptr(), L.stream_ptr()))
    return out_s, out_i


def score_topk(queries: torch.Tensor, index: CorpusIndex, k: int, id_offset: int = 0, force_exact: bool = False,
               stats: Optional[dict] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Exact fp32 top-k of `queries @ index.emb.T`: (scores [nq,k] f32, ids [nq,k] i64 = local index + id_offset).
    Rows are sorted by (score desc, id asc); if k > nd the tail is (-inf, -1)."""
    q = _check_f32(queries, "queries")
    if q.device != index.emb.device:
        raise ValueError(f"queries live on {q.device}, the index on {index.emb.device}")
    with L.on_device(q.device):
        return _score_topk(q, index, k, id_offset, force_exact, stats)


class _Stages:
    """Optional per-stage CUDA-event timing: pass stats={"stages": {}} and read stats["stages"] (ms) after a synchronize."""

    def __init__(self, stats: Optional[dict]):
        self.on = stats is not None and "stages" in stats
        self.stats = stats
        if self.on:
            self.last = torch.cuda.Event(enable_timing=True)
            self.last.record()
            stats.setdefault("_events", [])

    def mark(self, name: str) -> None:
        if self.on:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            self.stats["_events"].append((name, self.last, e))
            self.last = e


def resolve_stages(stats: dict) -> dict:
    """After torch.cuda.synchronize(): turn the recorded event pairs into stats["stages"][name] += ms."""
    for name, e0, e1 in stats.pop("_events", []):
        stats["stages"][name] = stats["stages"].get(name, 0.0) + e0.elapsed_time(e1)
    return stats["stages"]


def _score_topk(q: torch.Tensor, index: CorpusIndex, k: int, id_offset: int, force_exact: bool, stats: Optional[dict]):
    nq, d = q.shape
    nd = index.nd
    if nq == 0:
        return (torch.empty((0, k), dtype=torch.float32, device=q.device), torch.empty((0, k), dtype=torch.int64, device=q.device))
    if d != index.emb.shape[1]:
        raise ValueError("query / corpus dim mismatch")
    if force_exact or nq * nd <= SMALL_PROBLEM or nd < 256:
        if stats is not None:
            stats.update(path="exact", flagged=0)
        return _exact_topk(q, index, k, id_offset)
    lib = L.lib()
    ranges = lib.vr_score_ranges(nq, nd)
    lists = ranges * 2 * lib.vr_score_list_len()
    q16 = to_f16_rows(q)
    cand_s = torch.empty((nq, lists), dtype=torch.float32, device=q.device)
    cand_i = torch.empty((nq, lists), dtype=torch.int32, device=q.device)
    out_s = torch.empty((nq, k), dtype=torch.float32, device=q.device)
    out_i = torch.empty((nq, k), dtype=torch.int64, device=q.device)
    flags = torch.empty((nq,), dtype=torch.int32, device=q.device)
    sp = L.stream_ptr()
    ev = _Stages(stats)
    ev.mark("q_to_f16")
    L.check(lib.vr_score_filter(q16.data_ptr(), nq, index.emb_f16.data_ptr(), nd, d, ranges, cand_s.data_ptr(),
                                cand_i.data_ptr(), sp))
    ev.mark("filter")
    L.check(lib.vr_score_rescore(q.data_ptr(), nq, index.emb.data_ptr(), nd, d, ranges, cand_s.data_ptr(), cand_i.data_ptr(),
                                 index.max_norm.data_ptr(), k, id_offset, out_s.data_ptr(), out_i.data_ptr(),
                                 flags.data_ptr(), sp))
    ev.mark("rescore")
    bad = torch.nonzero(flags).flatten()  # host sync: the caller reads the result next anyway
    if stats is not None:
        stats.update(path="filter+rescore", flagged=int(bad.numel()), ranges=ranges)
    if bad.numel() > 0:
        s2, i2 = _exact_topk(q.index_select(0, bad), index, k, id_offset)
        out_s.index_copy_(0, bad, s2)
        out_i.index_copy_(0, bad, i2)
    return out_s, out_i


def merge_topk(scores: torch.Tensor, ids: torch.Tensor, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """[nq, m] candidate (score, id) pairs (id < 0 = empty) -> top-k by (score desc, id asc)."""
    scores = scores.contiguous().float()
    ids = ids.contiguous().to(torch.int64)
    nq, m = scores.shape
    out_s = torch.empty((nq, k), dtype=torch.float32, device=scores.device)
    out_i = torch.empty((nq, k), dtype=torch.int64, device=scores.device)
    with L.on_device(scores.device):
        L.check(L.lib().vr_topk_rows(scores.data_ptr(), ids.data_ptr(), nq, m, k, 0, out_s.data_ptr(), out_i.data_ptr(),
                                     L.stream_ptr()))
    return out_s, out_i


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous page range [lo, hi) owned by `rank` (corpus sharded by page)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_partials(scores: torch.Tensor, ids: torch.Tensor, group=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """The ONE collective of the retrieval path: all-gather every rank's [nq, k] (score, global id) pairs.
    Returns ([nq, world*k] scores, [nq, world*k] ids) on every rank. Score bits travel inside int64 so a single
    all_gather moves both arrays (12 -> 16 B per entry; the message is latency bound either way)."""
    import torch.distributed as dist

    world = dist.get_world_size(group)
    nq, k = scores.shape
    packed = torch.stack([scores.contiguous().view(torch.int32).to(torch.int64), ids.to(torch.int64)], dim=-1).contiguous()
    flat = torch.empty((world * nq, k, 2), dtype=torch.int64, device=packed.device)  # rank-major concatenation
    dist.all_gather_into_tensor(flat, packed, group=group)
    gathered = flat.view(world, nq, k, 2)
    gs = gathered[..., 0].to(torch.int32).view(torch.float32).permute(1, 0, 2).reshape(nq, world * k)
    gi = gathered[..., 1].permute(1, 0, 2).reshape(nq, world * k)
    return gs.contiguous(), gi.contiguous()


def sharded_topk(queries: torch.Tensor, index: CorpusIndex, k: int, id_offset: int, group=None, stats: Optional[dict] = None):
    """Corpus sharded by page across ranks (every rank holds the same queries): local exact top-k with GLOBAL ids,
    one all-gather of [nq, k] (score, id) pairs over NCCL/NVLink, k-way merge on every rank."""
    import torch.distributed as dist

    s, i = score_topk(queries, index, k, id_offset, stats=stats)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return s, i
    ev = _Stages(stats)
    gs, gi = gather_partials(s, i, group)
    ev.mark("all_gather_partials")
    out = merge_topk(gs, gi, k)
    ev.mark("merge")
    return out


def gather_queries(local: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """Queries sharded by rank for ENCODING (the reference's partition, `dense_retriever.py:48-50`): rank r encoded
    queries shard_range(n_total, r, world); one all-gather of the [ceil(n/world), d] fp32 blocks gives every rank all
    n_total embeddings in query order (10 k x 2304 fp32 = 92 MB: sub-millisecond over NVLink)."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    per = (n_total + world - 1) // world
    d = local.shape[1]
    block = torch.zeros((per, d), dtype=local.dtype, device=local.device)
    block[: local.shape[0]] = local
    flat = torch.empty((world * per, d), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(flat, block, group=group)
    parts = []
    for r in range(world):
        lo, hi = shard_range(n_total, r, world)
        parts.append(flat[r * per: r * per + (hi - lo)])
    return torch.cat(parts)


# ------------------------------------------------------------------------------------------------------
# Reference-signature drop-ins
# ------------------------------------------------------------------------------------------------------
def load_shard(path: str):
    """`pickle((float32[n,d], List[str]))` written by `inference.py:114-126`."""
    with open(path, "rb") as f:
        data = pickle.load(f)
    return np.asarray(data[0], dtype=np.float32), list(data[1])


def save_shard(path: str, emb: np.ndarray, lookup: List[str]) -> None:
    with open(path, "wb") as f:
        pickle.dump((np.asarray(emb, dtype=np.float32), list(lookup)), f, protocol=4)


def _retrieve_one_shard(corpus_shard_path: str, encoded_queries_tensor: torch.Tensor, topk: int, device: str):
    """Same contract as `dense_retriever.py:13-34`: indices index INTO the shard; lookup maps them to doc ids."""
    emb, lookup = load_shard(corpus_shard_path)
    index = build_index(emb, lookup, device=device)
    k = min(topk, index.nd)  # torch.topk would raise for k > n; a short shard simply yields fewer candidates
    scores, idx = score_topk(encoded_queries_tensor.to(device), index, k)
    return scores, idx, lookup


def distributed_parallel_retrieve(args, topk: int) -> Dict[str, Dict[str, float]]:
    """`dense_retriever.py:37-97`: this rank's query shards x ALL corpus shards -> {qid: {docid: score}}.
    The per-shard results are unioned exactly like the reference (<= k * n_shards docs per query); the element-wise
    `.item()` loop (`:88-92`) is replaced by one device->host copy per shard."""
    with torch.no_grad():
        q_parts = sorted(glob.glob(os.path.join(args.output_dir, f"embeddings.query.rank.{args.process_index}*")))
        encoded, query_lookup = [], []
        for part in q_parts:
            emb, ids = load_shard(part)
            if len(ids) == 0:
                continue
            encoded.append(emb)
            query_lookup.extend(ids)
        c_parts = sorted(glob.glob(os.path.join(args.output_dir, "embeddings.corpus.rank.*")))
        if len(c_parts) == 0:
            raise ValueError("No pre-computed document embeddings found")
        result: Dict[str, Dict[str, float]] = {qid: {} for qid in query_lookup}
        if not encoded:
            return result
        q = torch.from_numpy(np.concatenate(encoded)).to(args.device)
        for part in c_parts:
            scores, idx, lookup = _retrieve_one_shard(part, q, topk, args.device)
            s_host, i_host = scores.cpu().numpy(), idx.cpu().numpy()
            for r, qid in enumerate(query_lookup):
                row = result[qid]
                for j in range(s_host.shape[1]):
                    if i_host[r, j] >= 0:
                        row[lookup[int(i_host[r, j])]] = float(s_host[r, j])
        return result
