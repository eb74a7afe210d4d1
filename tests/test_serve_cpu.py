"""CPU: host logic of the batched serving call (hvae_b200/serve.py) against a literal restatement of the reference's
/recommend handler (src/api/server.py:115-183) on a small dense score table -- response fields, ordering, seen-item
masking, -inf / unmapped-item filtering, unknown users, top_k limits, grouping by exclude_seen, micro-batching."""
import threading

import numpy as np
import pytest

from hvae_b200.serve import MAX_TOP_K, MicroBatcher, RecommendService, UnknownUser


def _world(U=9, N=40, seed=0):
    rng = np.random.default_rng(seed)
    scores = rng.permutation(U * N).reshape(U, N).astype(np.float32) / 7.0       # distinct scores: argsort order is unique
    seen = [np.sort(rng.choice(N, size=int(rng.integers(0, 6)), replace=False)) for _ in range(U)]
    seen[4] = np.arange(N - 3)                                                    # almost everything seen: fewer than top_k answers
    user_to_idx = {f"U{u}": u for u in range(U)}
    idx_to_item = {i: f"B{i}" for i in range(N) if i != 11}                       # one item without an id
    calls = []

    def topk_fn(users, K, exclude_seen):
        calls.append((len(users), K, bool(exclude_seen)))
        vals, idxs = [], []
        for u in users:
            s = scores[u].copy()
            if exclude_seen:
                s[seen[u]] = -np.inf
            order = np.lexsort((-np.arange(N), -s))[:K]                            # (score desc, index desc): the kernels' order
            vals.append(s[order]); idxs.append(order)
        return np.array(vals, np.float32), np.array(idxs, np.int32)

    return scores, seen, user_to_idx, idx_to_item, topk_fn, calls


def _reference_handler(scores, seen, user_to_idx, idx_to_item, user_id, top_k, exclude_seen):
    """src/api/server.py:142-178 on a precomputed score row."""
    u = user_to_idx[user_id]
    s = scores[u].copy()
    if exclude_seen:
        s[seen[u]] = -np.inf
    top = np.argsort(s)[::-1][:top_k]
    recs = [{"item_id": idx_to_item[i], "score": float(s[i])} for i in top.tolist() if i in idx_to_item and not np.isinf(s[i])]
    return {"user_id": user_id, "recommendations": recs, "total_items": scores.shape[1]}


def test_recommend_matches_reference_handler():
    scores, seen, u2i, i2it, topk_fn, calls = _world()
    svc = RecommendService(topk_fn, u2i, i2it, scores.shape[1])
    for user_id in u2i:
        for top_k, ex in ((10, True), (5, False), (40, True), (1, True)):
            assert svc.recommend(user_id, top_k, ex) == _reference_handler(scores, seen, u2i, i2it, user_id, top_k, ex)
    assert len(svc.recommend("U4", 10, True)["recommendations"]) == 3             # only three unseen items left


def test_recommend_many_groups_and_errors():
    scores, seen, u2i, i2it, topk_fn, calls = _world()
    svc = RecommendService(topk_fn, u2i, i2it, scores.shape[1], max_batch=3)
    reqs = [("U0", 10, True), ("nobody", 10, True), ("U1", 5, False), ("U2", 20, True), ("U3", 0, True), ("U5", 7, True),
            ("U6", 3, True), ("U7", MAX_TOP_K + 1, True), ("U8", 2, False)]
    out = svc.recommend_many(reqs)
    assert isinstance(out[1], UnknownUser) and isinstance(out[4], ValueError) and isinstance(out[7], ValueError)
    for n, (uid, k, ex) in enumerate(reqs):
        if not isinstance(out[n], Exception):
            assert out[n] == _reference_handler(scores, seen, u2i, i2it, uid, k, ex)
    # exclude_seen=True: U0,U2,U5 | U6 (max_batch 3) with K = the group's largest top_k; exclude_seen=False: U1,U8
    assert sorted(calls) == sorted([(3, 20, True), (1, 3, True), (2, 5, False)])
    with pytest.raises(UnknownUser):
        svc.recommend("nobody")


def test_microbatcher_shares_launches():
    scores, seen, u2i, i2it, topk_fn, calls = _world()
    svc = RecommendService(topk_fn, u2i, i2it, scores.shape[1])
    gate = threading.Event()
    inner = svc.topk_fn

    def slow(users, K, ex):              # the first launch blocks until every other request is queued
        gate.wait(5.0)
        return inner(users, K, ex)

    svc.topk_fn = slow
    with MicroBatcher(svc, max_batch=64, max_wait_ms=1.0) as mb:
        first = mb.submit("U0", 10, True)
        futs = [mb.submit(f"U{u % 9}", 5 + u % 3, True) for u in range(1, 30)]
        bad = mb.submit("nobody", 10, True)
        gate.set()
        assert first.result(10) == _reference_handler(scores, seen, u2i, i2it, "U0", 10, True)
        for u, f in zip(range(1, 30), futs):
            assert f.result(10) == _reference_handler(scores, seen, u2i, i2it, f"U{u % 9}", 5 + u % 3, True)
        with pytest.raises(UnknownUser):
            bad.result(10)
        assert mb.batches_served <= 3     # 31 requests in at most three launches (not 31)
    with pytest.raises(RuntimeError):
        mb.submit("U0")
