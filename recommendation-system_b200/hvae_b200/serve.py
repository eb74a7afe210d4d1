"""Batched form of the API's scoring call (reference: src/api/server.py:115-183).

The reference answers one `/recommend` request at a time: densify the user's row, `get_user_embedding`, `decode`, mask
the seen items with -inf, `argsort`, keep `top_k`, drop -inf entries and items without an id.  `RecommendService` gives
the same answer (same fields, same ordering under (score desc, index desc), same error for an unknown user), but a
whole group of requests becomes ONE encoder + scoring + top-K launch (`RecommendationEvaluator.topk_users`).
`MicroBatcher` collects concurrent requests for a few milliseconds so that independent callers share a launch.
The FastAPI plumbing itself (routes, pydantic models, CORS) is out of scope: a handler calls
`service.recommend(...)` or `await loop.run_in_executor(None, batcher.submit(...).result)`.
"""
from __future__ import annotations

import threading
import time
from concurrent.futures import Future
from typing import Callable, Iterable, Sequence

import numpy as np

MAX_TOP_K = 100          # the API's own limit (src/api/server.py: RecommendationRequest.top_k <= 100)


class UnknownUser(KeyError):
    """The user id is not in the training data (the reference answers HTTP 404, server.py:134-138)."""


def _response(user_id, idx: np.ndarray, val: np.ndarray, top_k: int, idx_to_item: dict, total_items: int) -> dict:
    """RecommendationResponse of the reference (server.py:164-178): -inf scores and unmapped items are dropped."""
    recs = []
    for item_idx, score in zip(idx[:top_k].tolist(), val[:top_k].tolist()):
        if item_idx in idx_to_item and not np.isinf(score):
            recs.append({"item_id": idx_to_item[item_idx], "score": float(score)})
    return {"user_id": user_id, "recommendations": recs, "total_items": total_items}


class RecommendService:
    """topk_fn(user_indices, K, exclude_seen) -> (values [n, K], indices [n, K]) array-likes; the production binding is
    `RecommendationEvaluator.topk_users` (`from_evaluator`)."""

    def __init__(self, topk_fn: Callable, user_to_idx: dict, idx_to_item: dict, n_items: int, max_batch: int = 4096):
        self.topk_fn, self.user_to_idx, self.idx_to_item = topk_fn, user_to_idx, idx_to_item
        self.n_items, self.max_batch = int(n_items), int(max_batch)

    @classmethod
    def from_evaluator(cls, evaluator, idx_to_item: dict, max_batch: int = 4096):
        def topk(users, K, exclude_seen):
            v, i = evaluator.topk_users(users, K, exclude_seen)
            return v.cpu().numpy(), i.cpu().numpy()
        return cls(topk, evaluator.user_to_idx, idx_to_item, evaluator.n_items, max_batch)

    def _check(self, user_id, top_k):
        if user_id not in self.user_to_idx:
            raise UnknownUser(f"User '{user_id}' not found in training data")
        if not 1 <= int(top_k) <= MAX_TOP_K:
            raise ValueError(f"top_k must be in [1, {MAX_TOP_K}], got {top_k}")

    def recommend(self, user_id, top_k: int = 10, exclude_seen: bool = True) -> dict:
        out = self.recommend_many([(user_id, top_k, exclude_seen)])[0]
        if isinstance(out, Exception):
            raise out
        return out

    def recommend_many(self, requests: Sequence[tuple]) -> list:
        """requests: (user_id, top_k, exclude_seen) triples.  One launch per exclude_seen group (K = the group's largest
        top_k).  Returns, in request order, the response dict or the exception that request would have raised."""
        out: list = [None] * len(requests)
        groups: dict[bool, list[int]] = {}
        for n, (user_id, top_k, exclude_seen) in enumerate(requests):
            try:
                self._check(user_id, top_k)
            except (UnknownUser, ValueError) as e:
                out[n] = e
                continue
            groups.setdefault(bool(exclude_seen), []).append(n)
        for exclude_seen, members in groups.items():
            for s in range(0, len(members), self.max_batch):
                part = members[s:s + self.max_batch]
                K = min(max(int(requests[n][1]) for n in part), self.n_items)
                users = np.array([self.user_to_idx[requests[n][0]] for n in part], dtype=np.int32)
                val, idx = self.topk_fn(users, K, exclude_seen)
                val, idx = np.asarray(val), np.asarray(idx)
                for r, n in enumerate(part):
                    out[n] = _response(requests[n][0], idx[r], val[r], int(requests[n][1]), self.idx_to_item, self.n_items)
        return out


class MicroBatcher:
    """Collects `submit()` calls from any number of threads and serves them with `service.recommend_many` in batches:
    a batch closes after `max_wait_ms` or at `max_batch` requests, whichever comes first."""

    def __init__(self, service: RecommendService, max_batch: int = 256, max_wait_ms: float = 2.0):
        self.service, self.max_batch, self.max_wait = service, int(max_batch), max_wait_ms / 1e3
        self._cv = threading.Condition()
        self._queue: list[tuple[tuple, Future]] = []
        self._closed = False
        self.batches_served = 0
        self._thread = threading.Thread(target=self._run, name="hvae-microbatcher", daemon=True)
        self._thread.start()

    def submit(self, user_id, top_k: int = 10, exclude_seen: bool = True) -> Future:
        fut: Future = Future()
        with self._cv:
            if self._closed:
                raise RuntimeError("MicroBatcher is closed")
            self._queue.append(((user_id, top_k, exclude_seen), fut))
            self._cv.notify()
        return fut

    def close(self):
        with self._cv:
            self._closed = True
            self._cv.notify()
        self._thread.join()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _take(self) -> list:
        with self._cv:
            while not self._queue and not self._closed:
                self._cv.wait()
            if not self._queue:
                return []
            deadline = time.monotonic() + self.max_wait
            while len(self._queue) < self.max_batch and not self._closed:
                left = deadline - time.monotonic()
                if left <= 0:
                    break
                self._cv.wait(left)
            batch, self._queue = self._queue[:self.max_batch], self._queue[self.max_batch:]
            return batch

    def _run(self):
        while True:
            batch = self._take()
            if not batch:
                return
            try:
                results: Iterable = self.service.recommend_many([req for req, _ in batch])
            except Exception as e:      # a failed launch fails every request of the batch, not the serving thread
                results = [e] * len(batch)
            self.batches_served += 1
            for (_, fut), res in zip(batch, results):
                if isinstance(res, Exception):
                    fut.set_exception(res)
                else:
                    fut.set_result(res)
