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
        # vectorised over the rows still short of candidates (a per-row Python loop here cost ~0.5 s per 22k users, i.e. most of
        # a grid-search configuration's wall time): draw, drop seen / test / already-taken / repeated items, keep draw order
        cand = rng.integers(0, n_items, size=(todo.size, draw), dtype=np.int64)
        k = users[todo, None] * n_items + cand
        pos = np.searchsorted(keys, k)
        seen = (pos < keys.shape[0]) & (keys[np.minimum(pos, keys.shape[0] - 1)] == k)
        bad = seen | (cand == tests[todo, None])
        have_n = filled[todo]
        if have_n.any():                                    # later rounds: not the ones taken in earlier rounds
            have = out[todo, 1:]
            for j in range(int(have_n.max())):
                bad |= (cand == have[:, j:j + 1]) & (j < have_n)[:, None]
        order = np.argsort(cand, axis=1, kind="stable")     # repeats within the draw: keep the first occurrence
        srt = np.take_along_axis(cand, order, axis=1)
        dup_sorted = np.zeros_like(bad)
        dup_sorted[:, 1:] = srt[:, 1:] == srt[:, :-1]
        dup = np.zeros_like(bad)
        np.put_along_axis(dup, order, dup_sorted, axis=1)
        bad |= dup
        first_good = np.argsort(bad, axis=1, kind="stable")                 # good candidates first, in draw order
        n_good = (~bad).sum(axis=1)
        take = np.minimum(n_neg - have_n, n_good)
        slots = np.arange(draw)[None, :]
        sel = slots < take[:, None]
        picked = np.take_along_axis(cand, first_good, axis=1)
        rows_rep = np.repeat(todo, take)
        cols_rep = (1 + have_n)[:, None] + slots
        out[rows_rep, cols_rep[sel]] = picked[sel]
        filled[todo] = have_n + take
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


def sample_negatives_device(csr, users, tests, n_neg, seed=None):
    """The same candidate sets drawn ON THE DEVICE (torch index ops: sort / searchsorted): -> (cand int32
    [n, 1 + n_neg] on the device, valid int32 [n] on the host).  Rows of catalogues too small to yield n_neg unseen items fall
    back to the host sampler above."""
    dev = csr.indptr.device
    N = csr.n_items
    users_h, tests_h = np.asarray(users, dtype=np.int64), np.asarray(tests, dtype=np.int64)
    users_d, tests_d = torch.as_tensor(users_h, device=dev), torch.as_tensor(tests_h, device=dev)
    n = users_d.shape[0]
    gen = torch.Generator(device=dev)
    gen.manual_seed(int(seed) if seed is not None else int(np.random.randint(0, 2 ** 31 - 1)))
    keys = getattr(csr, "_pair_keys", None)
    if keys is None:        # (user, item) pairs as sorted int64 keys; CSR order is already sorted
        lens = csr.indptr[1:] - csr.indptr[:-1]
        keys = torch.repeat_interleave(torch.arange(csr.n_users, device=dev, dtype=torch.int64), lens) * N + csr.indices.to(torch.int64)
        csr._pair_keys = keys
    out = torch.full((n, 1 + n_neg), -1, dtype=torch.int64, device=dev)
    out[:, 0] = tests_d
    filled = torch.zeros(n, dtype=torch.int64, device=dev)
    todo = torch.arange(n, device=dev)
    draw = min(n_neg + 32, max(n_neg, N))
    for _ in range(8):
        if todo.numel() == 0:
            break
        cand = torch.randint(0, N, (todo.numel(), draw), generator=gen, device=dev)
        k = users_d[todo, None] * N + cand
        pos = torch.searchsorted(keys, k.reshape(-1)).reshape(k.shape).clamp_(max=max(keys.numel() - 1, 0))
        bad = (keys[pos] == k) if keys.numel() else torch.zeros_like(k, dtype=torch.bool)
        bad |= cand == tests_d[todo, None]
        # repeats: within the draw and against what the row already holds -- stable sort, every later equal value is a repeat
        have = out[todo, 1:]                                         # -1 in free slots
        both = torch.cat([have, cand], dim=1)
        srt, order = torch.sort(both, dim=1, stable=True)
        dup_sorted = torch.zeros_like(both, dtype=torch.bool)
        dup_sorted[:, 1:] = srt[:, 1:] == srt[:, :-1]
        dup = torch.zeros_like(dup_sorted).scatter_(1, order, dup_sorted)
        bad |= dup[:, n_neg:]
        good = ~bad
        rank = torch.cumsum(good, dim=1)                              # 1-based position among the good candidates, draw order
        need = (n_neg - filled[todo])[:, None]
        sel = good & (rank <= need)
        rows = todo[:, None].expand_as(cand)[sel]
        cols = (filled[todo][:, None] + rank)[sel]                    # slot 1 + filled + (rank - 1)
        out[rows, cols] = cand[sel]
        filled[todo] = filled[todo] + sel.sum(dim=1)
        todo = todo[filled[todo] < n_neg]
    valid = np.full(n, n_neg, dtype=np.int32)
    cand32 = out.to(torch.int32)
    if todo.numel():          # tiny catalogues: the host sampler takes whatever is available for these rows
        idx = todo.cpu().numpy()
        hc, hv = sample_negatives(csr.indptr.cpu().numpy(), csr.indices.cpu().numpy(), N, users_h[idx], tests_h[idx], n_neg,
                                  np.random.default_rng(seed))
        cand32[todo] = torch.from_numpy(hc).to(dev)
        valid[idx] = hv
    return cand32.contiguous(), valid


def candidate_ranks(ev, users, cand):
    """0-based rank of candidate 0 among each row's candidates, computed on the device."""
    eng = ev.model.engine
    dev = ev.device
    n, C = cand.shape
    ranks = torch.empty(n, dtype=torch.int32, device=dev)
    cand_d = cand.contiguous() if isinstance(cand, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(cand)).to(dev)
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
    else:       # drawn on the device (the host sampler `sample_negatives` is the NumPy statement the CPU tests check)
        cand, valid = sample_negatives_device(ev._csr, users, tests, n_negatives, seed)
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
