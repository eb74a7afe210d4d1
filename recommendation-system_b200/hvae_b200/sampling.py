"""99-negative evaluation protocol on the device (reference src/ml/evaluate.py:149-215, the `make evaluate` default).

Per (user, held-out item) row: 99 items the user has not interacted with are drawn on the host (NumPy, as the
reference does -- seedable here, the reference is unseeded), the 100 candidate scores come from one gather-dot
launch (u_b . E_c; no [B, N] score matrix), the rank of the held-out item among them gives Recall/NDCG/HR@K.
"""
from __future__ import annotations

import numpy as np
import torch

from ._cabi import p
from .engine import Batch, r4, r8


def sample_negatives(indptr, indices, n_items, users, tests, n_neg, rng):
    """[n, 1 + n_neg] int32: column 0 = test item, the rest distinct unseen items (fewer than n_neg available ->
    the row is padded by repeating the last candidate, which cannot outrank itself twice: see `valid`)."""
    users = np.asarray(users, dtype=np.int64)
    tests = np.asarray(tests, dtype=np.int64)
    n = users.shape[0]
    out = np.empty((n, 1 + n_neg), dtype=np.int32)
    out[:, 0] = tests
    valid = np.full(n, n_neg, dtype=np.int32)
    lens = np.diff(indptr)
    keys = (np.repeat(np.arange(len(lens), dtype=np.int64), lens) * n_items + indices.astype(np.int64))   # sorted (CSR order)
    draw = min(max(2 * n_neg, n_neg + 32), max(n_neg, n_items))
    todo = np.arange(n)
    filled = np.zeros(n, dtype=np.int64)
    for _ in range(8):
        if todo.size == 0:
            break
        cand = rng.integers(0, n_items, size=(todo.size, draw), dtype=np.int64)
        k = users[todo, None] * n_items + cand
        pos = np.searchsorted(keys, k)
        seen = (pos < keys.shape[0]) & (keys[np.minimum(pos, keys.shape[0] - 1)] == k)
        bad = seen | (cand == tests[todo, None])
        for j, row in enumerate(todo):                      # de-duplicate within the row, keep draw order
            c = cand[j][~bad[j]]
            _, first = np.unique(c, return_index=True)
            c = c[np.sort(first)]
            have = out[row, 1:1 + filled[row]]
            c = c[~np.isin(c, have)]
            take = min(n_neg - filled[row], c.shape[0])
            out[row, 1 + filled[row]:1 + filled[row] + take] = c[:take]
            filled[row] += take
        todo = todo[filled[todo] < n_neg]
    for row in todo:                                        # tiny catalogues: take everything that is available
        mask = np.ones(n_items, dtype=bool)
        mask[indices[indptr[users[row]]:indptr[users[row] + 1]]] = False
        mask[tests[row]] = False
        avail = np.where(mask)[0]
        pick = avail if avail.shape[0] <= n_neg else rng.choice(avail, n_neg, replace=False)
        out[row, 1:1 + pick.shape[0]] = pick
        filled[row] = pick.shape[0]
        valid[row] = pick.shape[0]
        out[row, 1 + pick.shape[0]:] = tests[row]           # padding never counts (masked by `valid` below)
    return out, valid


def candidate_ranks(ev, users, cand: np.ndarray):
    """0-based rank of candidate 0 among each row's candidates, computed on the device."""
    eng = ev.model.engine
    dev = ev.device
    n, C = cand.shape
    ranks = torch.empty(n, dtype=torch.int32, device=dev)
    cand_d = torch.from_numpy(np.ascontiguousarray(cand)).to(dev)
    users_d = torch.as_tensor(np.asarray(users, dtype=np.int32), device=dev)
    d = eng.lay.d
    with torch.no_grad():
        for s in range(0, n, ev.batch_users):
            rows = users_d[s:s + ev.batch_users].contiguous()
            B = rows.shape[0]
            u = eng.user_vectors(Batch(ev._csr, rows, B, 1))
            if eng.precision == "fp32":
                U, ldu, E, lde, bf = u, r4(d), eng.E, d, 0
            else:
                from . import tc
                U, ldu, E, lde, bf = tc.user_vectors_bf16(eng, u, B), r8(d), eng.E_bf16, r8(d), 1
            eng.lib.candidate_rank(p(U), ldu, p(E), lde, d, bf, cand_d.data_ptr() + 4 * s * C, C, B, None,
                                   ranks.data_ptr() + 4 * s, eng.stream)
    return ranks.cpu().numpy()


def evaluate_with_negatives(ev, users, tests, n_negatives, k_values, seed=None, negatives=None):
    """`negatives` [n, n_negatives] (optional): use these candidates instead of drawing them -- how a run of the reference,
    whose draws come from the unseeded np.random (src/ml/evaluate.py:169, tune.py:163), is replayed exactly."""
    if negatives is not None:
        negatives = np.asarray(negatives, dtype=np.int32)
        n_negatives = negatives.shape[1]
        cand = np.concatenate([np.asarray(tests, dtype=np.int32)[:, None], negatives], axis=1)
        valid = np.full(cand.shape[0], n_negatives, dtype=np.int32)
    else:
        rng = np.random.default_rng(seed)
        csr = ev._csr
        indptr = csr.indptr.cpu().numpy()
        indices = csr.indices.cpu().numpy()
        cand, valid = sample_negatives(indptr, indices, ev.n_items, users, tests, n_negatives, rng)
    ranks = candidate_ranks(ev, users, cand).astype(np.int64)
    short = valid < n_negatives
    if short.any():       # padded slots repeat the test item (score == test score -> counted as ahead): remove them
        ranks[short] -= (n_negatives - valid[short])
    out = {}
    for k in k_values:
        hit = ranks < k
        ndcg = np.where(hit, 1.0 / np.log2(ranks + 2.0), 0.0)
        out[k] = {"recall": float(hit.mean()), "ndcg": float(ndcg.mean()), "hit_ratio": float(hit.mean())}
    return out
