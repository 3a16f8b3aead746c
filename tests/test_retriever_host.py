"""Host-side retrieval logic that needs no GPU: shard pickle format, page sharding, refusal of CPU tensors,
and the world_size-2 gloo run of the partial-top-k exchange (the one collective of the path)."""
import os
import pickle

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import restated as O
from visrag_b200 import retriever as R


def test_shard_pickle_format_matches_reference(tmp_path):
    emb = np.random.RandomState(0).randn(5, 8).astype(np.float64)
    path = str(tmp_path / "embeddings.corpus.rank.0.0-5")
    R.save_shard(path, emb, [f"d{i}" for i in range(5)])
    with open(path, "rb") as f:
        data = pickle.load(f)      # exactly what `dense_retriever.py:19-23` unpickles
    assert isinstance(data, tuple) and data[0].dtype == np.float32 and data[0].shape == (5, 8) and data[1][4] == "d4"
    e2, ids = R.load_shard(path)
    assert np.array_equal(e2, emb.astype(np.float32)) and ids == data[1]


def test_shard_range_partitions_pages():
    for n in (0, 1, 7, 8, 1000003):
        for world in (1, 2, 3, 8):
            spans = [R.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_cpu_tensors_are_refused():
    with pytest.raises(ValueError, match="no CPU path"):
        R.build_index(torch.zeros(4, 8), device="cpu")


def _worker(rank, world, port, q_out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rs = np.random.RandomState(5)
    Q = rs.randn(6, 16).astype(np.float32)
    D = rs.randn(41, 16).astype(np.float32)
    lo, hi = R.shard_range(len(D), rank, world)
    s, i = O.score_topk(Q, D[lo:hi], 4)                       # local partial top-k (oracle stands in for the kernels)
    gs, gi = R.gather_partials(torch.from_numpy(s), torch.from_numpy(i + lo))
    ms, mi = O.merge_topk([(gs.numpy(), gi.numpy())], 4)
    want_s, want_i = O.score_topk(Q, D, 4)
    # query encode sharded by rank: each rank holds its slice of the query embeddings, one all-gather restores the order
    for nq in (6, 5, 1):
        qlo, qhi = R.shard_range(nq, rank, world)
        allq = R.gather_queries(torch.from_numpy(Q[:nq][qlo:qhi].copy()), nq)
        assert np.array_equal(allq.numpy(), Q[:nq]), (nq, rank)
    q_out.put((rank, bool(np.array_equal(mi, want_i) and np.allclose(ms, want_s)), tuple(gs.shape)))
    dist.destroy_process_group()


def test_partial_topk_all_gather_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
    assert sorted(r[0] for r in res) == [0, 1]
    assert all(r[1] for r in res) and all(r[2] == (6, 8) for r in res)
