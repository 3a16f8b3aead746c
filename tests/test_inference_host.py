"""Run files + metrics (drop-ins for src/openmatch/utils.py and driver/eval.py:272-304). CPU only."""
import math

import numpy as np

import pytest

from visrag_b200 import inference as I


def test_trec_round_trip_and_format(tmp_path):
    run = {"q1": {"dA": 0.5, "dB": 0.9, "dC": -0.1}, "q2": {"dZ": 1.25}}
    p = str(tmp_path / "sub" / "test.0.trec")
    I.save_as_trec(run, p)
    lines = open(p).read().splitlines()
    assert lines[0] == "q1\tQ0\tdB\t1\t0.9\tOpenMatch" and lines[2].split("\t")[2:4] == ["dC", "3"]
    assert I.load_from_trec(p) == run
    assert I.load_from_trec(p, as_list=True)["q1"][0] == ("dB", 0.9)
    assert list(I.load_from_trec(p, max_len_per_q=2)["q1"]) == ["dB", "dA"]
    (tmp_path / "three.trec").write_text("q1\td1\t0.3\n")
    assert I.load_from_trec(str(tmp_path / "three.trec")) == {"q1": {"d1": 0.3}}
    (tmp_path / "bad.trec").write_text("q1 d1\n")
    with pytest.raises(ValueError):
        I.load_from_trec(str(tmp_path / "bad.trec"))


def test_metrics_hand_computed():
    qrel = {"q1": {"d1": 1, "d2": 2, "d9": 0}, "q2": {"d5": 1}, "q3": {"d7": 1}}
    run = {"q1": {"d3": 0.9, "d2": 0.8, "d1": 0.1}, "q2": {"d4": 0.5, "d6": 0.4}}
    mrr = I.eval_mrr(qrel, run, 10)
    assert mrr["q1"] == 0.5 and mrr["q2"] == 0.0 and mrr["all"] == 0.25
    assert I.eval_mrr(qrel, run, 1)["q1"] == 0.0
    rec = I.recall_at_k(qrel, run, 2)
    assert rec["q1"] == 0.5 and rec["q2"] == 0.0 and rec["all"] == 0.25
    nd = I.ndcg_at_k(qrel, run, 10)
    dcg = 2 / math.log2(3) + 1 / math.log2(4)
    idcg = 2 / math.log2(2) + 1 / math.log2(3)
    assert abs(nd["q1"] - dcg / idcg) < 1e-12 and nd["q2"] == 0.0
    # trec_eval tie rule: equal scores -> higher doc id first
    assert I._trec_ranking({"a": 1.0, "b": 1.0, "c": 2.0}) == ["c", "b", "a"]


def test_collator_and_save_results(tmp_path):
    b = I.naive_collator([{"id": "1", "text": "a", "image": None}, {"id": "2", "text": "b", "image": None}])
    assert b == {"id": ["1", "2"], "text": ["a", "b"], "image": [None, None]}
    out = I.save_results(str(tmp_path), {"q": {"d": 1}}, {"q": {"d": 0.3, "e": 0.9}})
    assert out["recall_10"] == 1.0 and out["mrr_10"] == 0.5
    assert len(open(tmp_path / "test_result.log").read().splitlines()) == 3


def test_knowledge_base_files_match_the_demo_layout(tmp_path):
    """reps.npy + index2img_filename.txt exactly as visrag_pipeline/build_index.py:52-58 writes them."""
    from visrag_b200 import knowledge_base as KB

    reps = np.random.RandomState(0).randn(5, 8).astype(np.float32)
    names = [f"doc.pdf_{i}.png" for i in range(5)]
    KB.save_knowledge_base(str(tmp_path / "kb"), reps, names)
    assert np.array_equal(np.load(tmp_path / "kb" / "reps.npy"), reps) and np.load(tmp_path / "kb" / "reps.npy").dtype == np.float32
    assert (tmp_path / "kb" / "index2img_filename.txt").read_text() == "\n".join(names)   # no trailing newline
    with pytest.raises(ValueError):
        KB.save_knowledge_base(str(tmp_path / "kb2"), reps, names[:4])
    assert KB.DEMO_QUERY_PREFIX.endswith("document: ")


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the host-cores arm) must print ONE JSON line with the contract's keys; run here on
    the tiny model so that it takes seconds."""
    import json
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--model", "tiny", "--steps", "2",
                        "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=root)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "page-images encoded/sec" and d["unit"] == "pages/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] >= 1 and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "pages/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "configs[2]" in d["config"]["workload"]
