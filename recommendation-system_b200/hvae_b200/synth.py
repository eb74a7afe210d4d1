"""Synthetic interaction data of the shapes BASELINE.json names.

Host-side numpy only; this is the input generator shared by the tests, the
bench and the oracle.  It mirrors what the reference's offline pipeline leaves
on disk for the hot path: a user x item CSR of positives
(/root/reference/src/ml/train.py:175-182), an L2-normalised item embedding
matrix (/root/reference/src/preprocessing/embeddings.py:62) and one held-out
test item per user (/root/reference/src/preprocessing/dataset.py:71-91).
"""

from __future__ import annotations

from dataclasses import dataclass

import numpy as np

# Shapes of BASELINE.json `configs` (C1..C4); hyper-parameters per SURVEY.md §8d.
CONFIGS = {
    "c1": dict(n_users=2072, n_items=890, emb_dim=384, latent_dim=128, hidden_dims=[512], dropout=0.3,
               beta=0.2, batch=512),
    "c2": dict(n_users=22363, n_items=12101, emb_dim=384, latent_dim=200, hidden_dims=[600], dropout=0.5,
               beta=0.2, batch=512),
    "c3": dict(n_users=1_000_000, n_items=200_000, emb_dim=768, latent_dim=200, hidden_dims=[600],
               dropout=0.5, beta=0.2, batch=4096),
    "c4": dict(n_users=1_000_000, n_items=1_000_000, emb_dim=768, latent_dim=200, hidden_dims=[600],
               dropout=0.5, beta=0.2, batch=4096),
}


@dataclass
class SynthData:
    indptr: np.ndarray      # int64 [U+1]
    indices: np.ndarray     # int32 [nnz], sorted within a row, no duplicates
    values: np.ndarray      # float32 [nnz], all 1.0
    n_users: int
    n_items: int
    test_items: np.ndarray  # int32 [U], never inside the user's row

    def scipy_csr(self):
        from scipy.sparse import csr_matrix

        return csr_matrix((self.values.astype(np.float64), self.indices, self.indptr),
                          shape=(self.n_users, self.n_items))


def make_interactions(n_users: int, n_items: int, seed: int = 0) -> SynthData:
    """Zipf-popular items, log-normal basket sizes, de-duplicated, CSR int32."""
    rng = np.random.default_rng(seed)
    cap = max(1, min(200, n_items - 1))
    counts = np.clip(np.rint(rng.lognormal(np.log(8.0), 0.6, n_users)), min(3, cap), cap).astype(np.int64)
    pop = 1.0 / (np.arange(n_items, dtype=np.float64) + 1.0)
    cdf = np.cumsum(pop)
    cdf /= cdf[-1]
    total = int(counts.sum())
    items = np.minimum(np.searchsorted(cdf, rng.random(total), side="right"), n_items - 1).astype(np.int64)
    users = np.repeat(np.arange(n_users, dtype=np.int64), counts)
    key = np.unique(users * n_items + items)          # sorts by (user, item) and drops duplicates
    users, items = key // n_items, (key % n_items).astype(np.int32)
    indptr = np.zeros(n_users + 1, dtype=np.int64)
    np.cumsum(np.bincount(users, minlength=n_users), out=indptr[1:])
    values = np.ones(items.shape[0], dtype=np.float32)

    # one held-out item per user, uniform over items outside the row
    rng_t = np.random.default_rng(seed + 2)
    test = rng_t.integers(0, n_items, n_users).astype(np.int64)
    seen_keys = key  # sorted
    for _ in range(64):
        tk = np.arange(n_users, dtype=np.int64) * n_items + test
        pos = np.searchsorted(seen_keys, tk)
        clash = (pos < seen_keys.shape[0]) & (seen_keys[np.minimum(pos, seen_keys.shape[0] - 1)] == tk)
        if not clash.any():
            break
        test[clash] = rng_t.integers(0, n_items, int(clash.sum()))
    return SynthData(indptr, items, values, n_users, n_items, test.astype(np.int32))


def make_item_embeddings(n_items: int, emb_dim: int, seed: int = 0) -> np.ndarray:
    """Row-normalised standard-normal matrix [N, d] float32 (SBERT stand-in)."""
    rng = np.random.default_rng(seed + 1)
    e = rng.standard_normal((n_items, emb_dim), dtype=np.float32)
    e /= np.linalg.norm(e, axis=1, keepdims=True).astype(np.float32)
    return e


def csr_rows(data: SynthData, rows: np.ndarray):
    """Slice a batch of users out of the CSR: (indptr int32 [B+1], indices int32, values f32)."""
    rows = np.asarray(rows, dtype=np.int64)
    starts, ends = data.indptr[rows], data.indptr[rows + 1]
    lens = ends - starts
    indptr = np.zeros(rows.shape[0] + 1, dtype=np.int32)
    np.cumsum(lens, out=indptr[1:])
    if lens.sum() == 0:
        return indptr, np.zeros(0, np.int32), np.zeros(0, np.float32)
    take = np.repeat(starts - indptr[:-1], lens) + np.arange(int(lens.sum()), dtype=np.int64)
    return indptr, data.indices[take], data.values[take]
