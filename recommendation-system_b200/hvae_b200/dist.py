"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL over NVLink/NVSwitch) -- SURVEY.md §8e.

The reference is single-process; these are the three ways its hot path shards on one 8xB200 box:

* training, data parallel over users (`DataParallel`): every rank holds the full model and the full
  interaction CSR (38 MB at 1M users).  A global batch is split evenly; each rank runs forward/backward on
  its users with every mean taken over the GLOBAL batch (src/ml/model.py:281,287).  Two collectives per step:
    - at the start, on a side stream: all-gather of the batch's user ids (2 KB), then the item-major transposition
      of the GLOBAL batch -- overlapped with the forward pass;
    - after backward: ONE all-gather of [small dense gradients (~2 MB) | d(pre-activation of layer 1) [B_local, h]]
      per rank.  The dense gradients are summed over ranks by our own kernel in a fixed order; every rank reduces
      the layer-1 weight gradient of the WHOLE batch locally, reading the gathered blocks in place.  The dense
      d(W1) [N, h] (480 MB at N=200k) never crosses NVLink.
  Gradient norm, clip coefficient and the fused Adam step are then identical on every rank.
* evaluation, item-sharded (`ItemShard` + `sharded_topk`): rank g scores items [lo_g, hi_g) only, keeps a
  local top-K of (score, global item id), all-gathers K candidates per user and merges G*K -> K.
* independent trainings (grid sweep): `assign_round_robin`.

Everything here works on CPU tensors with the gloo backend too (tests/test_dist_cpu.py); only the kernels
behind `Engine` need a GPU.
"""
from __future__ import annotations

import os

import numpy as np
import torch
import torch.distributed as dist


def split_even(n: int, world: int, rank: int):
    """[lo, hi) of rank's contiguous share of n units; the first n % world ranks get one extra."""
    q, r = divmod(n, world)
    lo = rank * q + min(rank, r)
    return lo, lo + q + (1 if rank < r else 0)


def assign_round_robin(n_units: int, world: int, rank: int):
    """Unit ids of rank for independent work (grid-search configurations, src/ml/tune.py:241)."""
    return list(range(rank, n_units, world))


class ItemShard:
    """Contiguous item range owned by a rank in item-sharded evaluation."""

    def __init__(self, n_items: int, world: int, rank: int):
        self.n_items, self.world, self.rank = n_items, world, rank
        self.lo, self.hi = split_even(n_items, world, rank)

    def ranges(self):
        return [split_even(self.n_items, self.world, r) for r in range(self.world)]


def gather_candidates(val: torch.Tensor, idx: torch.Tensor, group=None):
    """All-gather per-rank top-K candidates [B, K] -> ([B, G*K] values, [B, G*K] global ids), rank-major."""
    world = dist.get_world_size(group)
    B, K = val.shape
    gv = torch.empty(world * B, K, dtype=val.dtype, device=val.device)      # concatenation along dim 0 (gloo and nccl)
    gi = torch.empty(world * B, K, dtype=idx.dtype, device=idx.device)
    dist.all_gather_into_tensor(gv, val.contiguous(), group=group)
    dist.all_gather_into_tensor(gi, idx.contiguous(), group=group)
    gv, gi = gv.view(world, B, K), gi.view(world, B, K)
    return gv.permute(1, 0, 2).reshape(B, world * K).contiguous(), gi.permute(1, 0, 2).reshape(B, world * K).contiguous()


def merge_candidates_host(cval: np.ndarray, cidx: np.ndarray, K: int):
    """NumPy statement of hvae_topk_merge (total order: score desc, index desc; id < 0 = empty slot).
    Test infrastructure for the CPU (gloo) tests; the GPU path calls the kernel."""
    B = cval.shape[0]
    out_v = np.full((B, K), -np.inf, dtype=np.float32)
    out_i = np.full((B, K), -1, dtype=np.int32)
    for b in range(B):
        ok = cidx[b] >= 0
        v, i = cval[b][ok], cidx[b][ok]
        order = np.lexsort((-i.astype(np.int64), -v.astype(np.float64)))[:K]
        out_v[b, :len(order)], out_i[b, :len(order)] = v[order], i[order]
    return out_v, out_i


def sharded_topk(eng, batch, K: int, shard: ItemShard, exclude_seen: bool = True, group=None):
    """Item-sharded full ranking: local top-K over the rank's items, all-gather, merge."""
    from ._cabi import p
    v, i = eng.topk(batch, K, exclude_seen, shard.lo, shard.hi)
    if shard.world == 1:
        return v, i
    cv, ci = gather_candidates(v, i, group)
    out_v, out_i = torch.empty_like(v), torch.empty_like(i)
    eng.lib.topk_merge(p(cv), p(ci), batch.B, shard.world * K, K, p(out_v), p(out_i), eng.stream)
    return out_v, out_i


class ShardedEvaluator:
    """Item-sharded full-ranking evaluation on G ranks (BASELINE.json configs[3]; the reference loop is
    src/ml/evaluate.py:243-265, one user at a time).  Per tile of `tile` users, captured once as ONE CUDA graph and replayed:

      encoder + projection of tile/G users on each rank (user-sharded: the encoder is not replicated)
        -> all-gather of the bf16 user vectors (NCCL over NVLink; 6 MB per 4,096 users at d = 768)
        -> fused tcgen05 GEMM + seen mask + top-K over the rank's item shard, (value, id) written straight into one send block
        -> ONE all-gather of the packed [value | id] blocks
        -> merge of the G x K candidates per user read in place from the gathered blocks (hvae_topk_merge_groups).

    With world == 1 the two collectives drop out.  bf16 mode only (the fp32 path materialises scores: see sharded_topk)."""

    def __init__(self, eng, csr, K: int, group=None, tile: int = 4096, use_graph: bool = True, sharded: bool = True):
        from . import tc
        if eng.precision == "fp32" or K > tc.MAX_K_TC:
            raise RuntimeError("ShardedEvaluator runs the fused bf16 top-K kernel (K <= %d); use sharded_topk otherwise" % tc.MAX_K_TC)
        self.eng, self.csr, self.K, self.group = eng, csr, K, group
        # sharded=False: this rank scores the whole catalogue for its own users (user-sharded evaluation, or a single GPU)
        self.world = dist.get_world_size(group) if (sharded and dist.is_initialized()) else 1
        self.rank = dist.get_rank(group) if (sharded and dist.is_initialized()) else 0
        self.tile = (tile + self.world - 1) // self.world * self.world
        self.shard = ItemShard(eng.lay.N, self.world, self.rank)
        dev, T = eng.dev, self.tile
        ld8 = (eng.lay.d + 7) // 8 * 8
        self.rows = torch.zeros(T, dtype=torch.int32, device=dev)                 # static input of the captured tile
        self.ub_all = torch.zeros(T, ld8, dtype=torch.bfloat16, device=dev)
        self.ub_send = torch.zeros(T // self.world, ld8, dtype=torch.bfloat16, device=dev)
        self.send = torch.zeros(2, T, K, dtype=torch.int32, device=dev)           # [value bits | id]
        self.recv = torch.zeros(self.world, 2, T, K, dtype=torch.int32, device=dev)
        self.out_val = torch.zeros(T, K, dtype=torch.float32, device=dev)
        self.out_idx = torch.zeros(T, K, dtype=torch.int32, device=dev)
        self.use_graph, self._graph = use_graph, None

    def _tile(self):
        from . import tc
        from ._cabi import p
        from .engine import Batch
        eng, T, K, W = self.eng, self.tile, self.K, self.world
        bl = T // W
        mine = self.rows[self.rank * bl:(self.rank + 1) * bl]
        u = eng.user_vectors(Batch(self.csr, mine, bl, 1))
        d, ld8 = eng.lay.d, self.ub_all.shape[1]
        dst = self.ub_send if W > 1 else self.ub_all
        eng.lib.cast_bf16(p(u), bl, d, (d + 3) // 4 * 4, p(dst), ld8, eng.stream)
        if W > 1:
            dist.all_gather_into_tensor(self.ub_all, self.ub_send, group=self.group)
        full = Batch(self.csr, self.rows, T, 1)
        if W == 1:
            tc.topk_bf16_from_ub(eng, full, self.ub_all, K, True, 0, eng.lay.N, self.out_val, self.out_idx)
            return
        tc.topk_bf16_from_ub(eng, full, self.ub_all, K, True, self.shard.lo, self.shard.hi, self.send[0].view(torch.float32), self.send[1])
        dist.all_gather_into_tensor(self.recv, self.send, group=self.group)
        eng.lib.topk_merge_groups(p(self.recv), self.recv.data_ptr() + 4 * T * K, T, W, 2 * T * K, K, p(self.out_val), p(self.out_idx),
                                  eng.stream)

    def run_tile(self):
        if not self.use_graph:
            return self._tile()
        if self._graph is None:
            self._tile()                                   # allocates every workspace
            torch.cuda.synchronize(self.eng.dev)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._tile()
            self._graph = g
            return
        self._graph.replay()

    def topk(self, users: torch.Tensor, out_idx: torch.Tensor = None, out_val: torch.Tensor = None):
        """Top-K item ids [n, K] (int32, device) of `users` (device int32 ids; every rank passes the same list); the scores too
        when `out_val` [n, K] f32 is given."""
        n, T = users.shape[0], self.tile
        out = out_idx if out_idx is not None else torch.empty(n, self.K, dtype=torch.int32, device=self.eng.dev)
        with torch.no_grad():
            for s in range(0, n, T):
                m = min(T, n - s)
                self.rows[:m].copy_(users[s:s + m])
                if m < T:
                    self.rows[m:].copy_(users[:1].expand(T - m))      # pad slots repeat a valid user; their results are dropped
                self.run_tile()
                out[s:s + m].copy_(self.out_idx[:m])
                if out_val is not None:
                    out_val[s:s + m].copy_(self.out_val[:m])
        return out


class DataParallel:
    """Gradient exchange of data-parallel training; installed as `engine.dist`."""

    def __init__(self, group=None):
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        # gradient exchange: "nvl" = our one-shot all-gather by direct (multicast) stores into symmetric memory,
        # "nccl" = NCCL all-gather.  "auto" = nvl if symmetric memory can be set up on EVERY rank (decided once,
        # collectively, in prepare_exchange -- never inside a step), else nccl.
        self.exchange_mode = os.environ.get("HVAE_DP_EXCHANGE", "auto")
        self.exchange = None          # "nvl" | "nccl" once decided
        self._sym = None
        # HVAE_DP_EARLY_DENSE=1: the dense gradients from fc_mu on (fc_mu/fc_logvar, the projection MLP: ~all of the dense parameters) are
        # final well before the end of the backward pass; they are all-reduced by NCCL on a side stream under the encoder's backward
        # kernels and only the encoder's small tensors (ready last) travel with the dH1 rows in the step-end exchange.  Measured at N = 8
        # (profiles/r02_ncu_summary.md section 5): the exchange shrinks 0.208 -> 0.144 ms, the NCCL kernels beside the backward pass cost
        # the same again (3.243 vs 3.258 ms per step) -- off by default.
        self.early_dense = os.environ.get("HVAE_DP_EARLY_DENSE", "0") == "1"

    # -- batch splitting -------------------------------------------------------------------------------------
    def local_rows(self, global_rows):
        """This rank's contiguous slice of a global batch of user ids."""
        lo, hi = split_even(len(global_rows), self.world, self.rank)
        return global_rows[lo:hi]

    def b_max(self, b_global: int) -> int:
        return (b_global + self.world - 1) // self.world

    # -- collectives --------------------------------------------------------------------------------------------
    def gather_rows(self, rows_local: torch.Tensor, b_global: int):
        """All-gather the user ids of the global batch, padded per rank to b_max with -1 (= no user)."""
        bm = self.b_max(b_global)
        send = torch.full((bm,), -1, dtype=torch.int32, device=rows_local.device)
        send[:rows_local.shape[0]] = rows_local
        out = torch.empty(self.world * bm, dtype=torch.int32, device=rows_local.device)
        dist.all_gather_into_tensor(out, send, group=self.group)
        return out

    def gather_dpre(self, dpre_local: torch.Tensor, b_global: int):
        """All-gather d(pre-activation of layer 1) rows [B_local, ld] -> [world*b_max, ld] (pad rows zero)."""
        bm = self.b_max(b_global)
        ld = dpre_local.shape[1]
        if dpre_local.shape[0] == bm:
            send = dpre_local.contiguous()
        else:
            send = torch.zeros(bm, ld, dtype=dpre_local.dtype, device=dpre_local.device)
            send[:dpre_local.shape[0]] = dpre_local
        out = torch.empty(self.world * bm, ld, dtype=dpre_local.dtype, device=dpre_local.device)
        dist.all_gather_into_tensor(out, send, group=self.group)
        return out

    def reduce_dense(self, gd: torch.Tensor):
        dist.all_reduce(gd, op=dist.ReduceOp.SUM, group=self.group)

    def reduce_losses(self, acc: torch.Tensor):
        """Per-rank loss sums were scaled by 1/B_global, so the global loss is their sum (acc[3] = step count)."""
        t = acc[:3].clone()
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        acc[:3] = t

    # -- Engine hooks (engine.Engine.train_step) ----------------------------------------------------------------
    # Every buffer is a persistent workspace tensor, so both collectives are captured into the step's CUDA graph.
    def gather_batch(self, eng, batch):
        """User ids of the global batch (tiny all-gather).  Needs nothing from the forward pass: the engine issues it on a
        side stream at the start of the step, followed by the item-major transposition of the global batch."""
        from .engine import Batch
        rows = batch.rows
        if rows is None:
            raise RuntimeError("data-parallel steps need explicit user ids (batch.rows)")
        bm, B = self.b_max(eng.b_global), batch.B
        send_rows = eng.ws.get("dp_rows_send", (bm,), torch.int32)
        if B < bm:
            send_rows.fill_(-1)
        send_rows[:B].copy_(rows)
        rows_all = eng.ws.get("dp_rows_all", (self.world * bm,), torch.int32)
        dist.all_gather_into_tensor(rows_all, send_rows, group=self.group)
        cap = getattr(batch, "nnz_cap_global", None)
        if not cap:     # world * local bound is NOT a bound when other ranks hold denser users: insist on the real one
            raise RuntimeError("data-parallel steps need Batch.nnz_cap_global (an upper bound of the GLOBAL batch's nnz); "
                               "ShardedCSRLoader and VAETrainer.train_on_batch provide it")
        return Batch(batch.csr, rows_all, rows_all.shape[0], cap)

    def _symmetric(self, eng, n_floats):
        """Symmetric receive buffer [world * stride] + peer / multicast addresses (torch symmetric memory: CUDA VMM)."""
        if self._sym is not None and self._sym["n"] >= n_floats:
            return self._sym
        import torch.distributed._symmetric_memory as symm
        t = symm.empty(n_floats, dtype=torch.float32, device=eng.dev)
        hdl = symm.rendezvous(t, self.group if self.group is not None else dist.group.WORLD)
        mc = int(hdl.multicast_ptr or 0)                 # 0 when the fabric has no multicast (NVLS) support
        self._sym = dict(n=n_floats, buf=t, hdl=hdl, mc=mc, peers=int(hdl.buffer_ptrs_dev))
        eng.ws.generation += 1          # captured graphs must not keep pointers into an older buffer
        return self._sym

    def prepare_exchange(self, eng, n_floats):
        """Decide the gradient-exchange path ONCE and on all ranks together: try to set up the symmetric receive buffer,
        all-reduce a success flag, and fall back to NCCL only if some rank could not (a set-up error, never a step error;
        a rank falling back alone would leave the others in the symmetric-memory barrier)."""
        if self.exchange is not None and (self.exchange == "nccl" or (self._sym is not None and self._sym["n"] >= n_floats)):
            return self.exchange
        if self.exchange_mode == "nccl":
            self.exchange = "nccl"
            return self.exchange
        ok, err = 1, None
        try:
            self._symmetric(eng, n_floats)
        except Exception as e:          # set-up only (symmetric-memory allocation / rendezvous); reported below
            ok, err = 0, e
        flag = torch.tensor([ok], dtype=torch.int32, device=eng.dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        if int(flag.item()) == 1:
            self.exchange = "nvl"
        elif self.exchange_mode == "nvl":
            raise RuntimeError(f"HVAE_DP_EXCHANGE=nvl but symmetric memory is unavailable on some rank: {err}")
        else:
            self.exchange, self._sym = "nccl", None
        return self.exchange

    def dense_head(self, eng) -> int:
        """Floats of eng.gd exchanged at the end of the step (the rest was all-reduced early)."""
        return (eng.lay.slots["fc_mu.weight"].off - eng.lay.n_w1) if self.early_dense else eng.gd.numel()

    def reduce_dense_early(self, eng):
        """Called by Engine.backward once the weight / bias gradients of fc_mu..projection are enqueued (side streams 0 and 1): their
        all-reduce runs on its own stream beside the rest of the backward pass and is joined before the step-end exchange."""
        if not self.early_dense:
            return
        tail = eng.gd[self.dense_head(eng):]
        if not eng.concurrent:
            dist.all_reduce(tail, op=dist.ReduceOp.SUM, group=self.group)
            return
        with eng.side(6):
            for i in (0, 1):
                if i < len(eng._side):
                    torch.cuda.current_stream(eng.dev).wait_stream(eng._side[i])
            dist.all_reduce(tail, op=dist.ReduceOp.SUM, group=self.group)

    def exchange_grads(self, eng, batch, dpre0):
        n = self.world * (self.dense_head(eng) + self.b_max(eng.b_global) * dpre0.shape[1])
        if self.exchange is None or (self.exchange == "nvl" and self._sym["n"] < n):
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError("the gradient exchange must be set up before graph capture (run one eager step first)")
            self.prepare_exchange(eng, n)
        if self.exchange == "nvl":
            return self._exchange_grads_nvl(eng, batch, dpre0)
        return self._exchange_grads_nccl(eng, batch, dpre0)

    def _exchange_grads_nvl(self, eng, batch, dpre0):
        """Our own all-gather: every rank stores [dense gradients | dH1 rows] directly into all ranks' symmetric receive
        buffers (one NVSwitch-multicast store stream), bracketed by two symmetric-memory barriers; then the same fixed-order
        sum of the dense gradients and in-place consumption of the dH1 blocks as the NCCL path."""
        from ._cabi import p
        n_dense = self.dense_head(eng)
        bm, B, ld = self.b_max(eng.b_global), batch.B, dpre0.shape[1]
        stride = n_dense + bm * ld
        sym = self._symmetric(eng, self.world * stride)
        hdl, recv = sym["hdl"], sym["buf"]
        st = eng.stream
        hdl.barrier(channel=0)                       # every rank has consumed the previous step's blocks
        off = self.rank * stride
        eng.lib.nvl_push(p(eng.gd), n_dense, sym["mc"] or None, sym["peers"], self.world, off, st)
        eng.lib.nvl_push(p(dpre0), B * ld, sym["mc"] or None, sym["peers"], self.world, off + n_dense, st)
        hdl.barrier(channel=1)                       # all stores have landed everywhere
        cs = eng.ws.get("dp_colsum_ws", (n_dense + 64,))
        eng.lib.colsum(p(recv), stride, self.world, n_dense, p(eng.gd), p(cs), st)
        return recv.data_ptr() + 4 * n_dense, bm, stride

    def _exchange_grads_nccl(self, eng, batch, dpre0):
        """ONE all-gather per step: every rank sends [its dense gradients | its dH1 rows]; the dense gradients are then
        summed over ranks in a fixed order by our column-sum kernel (identical bits on every rank) and the layer-1
        weight-gradient kernel reads the gathered dH1 blocks in place.  -> (pointer to rank 0's dH1 block, rows per block,
        block stride in floats)."""
        from ._cabi import p
        n_dense = self.dense_head(eng)
        bm, B, ld = self.b_max(eng.b_global), batch.B, dpre0.shape[1]
        stride = n_dense + bm * ld
        ws = eng.ws
        send = ws.get("dp_send", (stride,))
        send[:n_dense].copy_(eng.gd[:n_dense])
        send[n_dense:n_dense + B * ld].copy_(dpre0[:B].reshape(-1))
        if B < bm:
            send[n_dense + B * ld:].zero_()
        recv = ws.get("dp_recv", (self.world * stride,))
        dist.all_gather_into_tensor(recv, send, group=self.group)
        cs = ws.get("dp_colsum_ws", (n_dense + 64,))
        eng.lib.colsum(p(recv), stride, self.world, n_dense, p(eng.gd), p(cs), eng.stream)
        return recv.data_ptr() + 4 * n_dense, bm, stride
