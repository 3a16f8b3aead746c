"""Demo knowledge-base layout and single-query retrieval.

Drop-in for the reference's demo pipeline:
  * `visrag_scripts/demo/visrag_pipeline/build_index.py:52-58` writes `reps.npy` (float32 [n, d], the L2-normalised page
    embeddings) and `index2img_filename.txt` ('\\n'-joined image basenames) into the knowledge-base directory;
  * `answer.py:26-35` (`retrieve`) reloads both files and re-uploads the embeddings for EVERY query, then
    `torch.matmul(query_rep, doc_reps.T)` + `torch.topk`.
Here the index is loaded once and stays resident in HBM (`KnowledgeBase`); a query is one fp32 scan of the index
(`vr_score_exact`, HBM bound: n*d*4 bytes) plus a two-level top-k spread over many SMs (`vr_topk_rows_chunked`).
Scores are the same fp32 dot products; ties are ordered by lower page index (torch.topk leaves tie order unspecified).
"""
from __future__ import annotations

import os
from typing import List, Sequence, Tuple

import numpy as np
import torch

from . import retriever

REPS_FILE = "reps.npy"
NAMES_FILE = "index2img_filename.txt"
# the demo's instruction (`answer.py:33`; singular "document", unlike the eval scripts' "documents")
DEMO_QUERY_PREFIX = "Represent this query for retrieving relevant document: "


def save_knowledge_base(path: str, reps, filenames: Sequence[str]) -> None:
    """Write the two files exactly as `build_index.py:52-58` does."""
    reps = np.asarray(reps.detach().cpu().numpy() if isinstance(reps, torch.Tensor) else reps, dtype=np.float32)
    if reps.ndim != 2 or reps.shape[0] != len(filenames):
        raise ValueError("reps must be [n, d] with one filename per row")
    if any("\n" in f for f in filenames):
        raise ValueError("filenames must not contain newlines")
    os.makedirs(path, exist_ok=True)
    np.save(os.path.join(path, REPS_FILE), reps)
    with open(os.path.join(path, NAMES_FILE), "w") as f:
        f.write("\n".join(filenames))


class KnowledgeBase:
    """A knowledge base resident on one GPU."""

    def __init__(self, path: str, device: str = "cuda"):
        self.path = path
        with open(os.path.join(path, NAMES_FILE), "r") as f:
            self.filenames: List[str] = f.read().split("\n")
        reps = np.load(os.path.join(path, REPS_FILE))
        if reps.ndim != 2 or reps.shape[0] != len(self.filenames):
            raise ValueError(f"{path}: reps.npy has {reps.shape} rows/dims but {len(self.filenames)} filenames")
        self.index = retriever.build_index(np.ascontiguousarray(reps, dtype=np.float32), self.filenames, device)

    def __len__(self) -> int:
        return self.index.nd

    def search(self, query_reps, topk: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """query_reps [nq, d] (tensor or ndarray, fp32) -> (scores [nq,k] f32, page indices [nq,k] i64) on the device."""
        q = query_reps if isinstance(query_reps, torch.Tensor) else torch.from_numpy(np.asarray(query_reps, dtype=np.float32))
        q = q.to(self.index.emb.device, torch.float32).reshape(-1, self.index.emb.shape[1]).contiguous()
        return retriever.score_topk(q, self.index, min(topk, len(self)))  # enters the index's device itself

    def retrieve(self, query_rep, topk: int) -> List[str]:
        """`answer.py: retrieve` after the query is encoded: paths of the top-k page images, best first."""
        _, ids = self.search(query_rep, topk)
        return [os.path.join(self.path, self.filenames[i]) for i in ids[0].tolist()]

    def retrieve_text(self, model, tokenizer, query: str, topk: int) -> List[str]:
        """Full `retrieve(knowledge_base_path, query, topk)`: instruction + query -> embedding (B2 wrapper) -> top-k."""
        out = model(query={"text": [DEMO_QUERY_PREFIX + query], "image": [None]}, tokenizer=tokenizer)
        return self.retrieve(out.q_reps, topk)
