"""CPU, gloo, world_size 2: the host-side logic of the multi-GPU paths (hvae_b200/dist.py) -- batch splitting and
padding, the gradient exchange collectives of data-parallel training, candidate gather + merge of item-sharded
evaluation, round-robin placement of independent trainings.  The kernels themselves need a GPU (tests -m gpu)."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parents[1]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    for p in (str(ROOT), str(ROOT / "recommendation-system_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from hvae_b200 import dist as hd
        from hvae_b200.engine import Batch
        from hvae_b200.train import ShardedCSRLoader
        dp = hd.DataParallel()
        assert (dp.world, dp.rank) == (world, rank)

        # --- batch split + padded all-gather of user ids and d(pre-activation) rows
        rng = np.random.default_rng(0)
        b_global, ld = 11, 8                               # odd on purpose: ranks get 6 and 5 users
        rows_g = torch.from_numpy(rng.permutation(100)[:b_global].astype(np.int32))
        lo, hi = hd.split_even(b_global, world, rank)
        rows_l = dp.local_rows(rows_g)
        assert torch.equal(rows_l, rows_g[lo:hi])
        rows_all = dp.gather_rows(rows_l, b_global)
        bm = dp.b_max(b_global)
        assert rows_all.shape[0] == world * bm
        dpre_g = torch.from_numpy(rng.standard_normal((b_global, ld)).astype(np.float32))
        dpre_all = dp.gather_dpre(dpre_g[lo:hi], b_global)
        # every real user appears once, at the slot the gathered gradient rows use; pads are -1 / zero rows
        got_rows, got_dpre = [], []
        for r in range(world):
            l2, h2 = hd.split_even(b_global, world, r)
            seg = rows_all[r * bm:(r + 1) * bm]
            assert torch.equal(seg[:h2 - l2], rows_g[l2:h2]) and torch.all(seg[h2 - l2:] == -1)
            assert torch.equal(dpre_all[r * bm:r * bm + (h2 - l2)], dpre_g[l2:h2])
            assert torch.all(dpre_all[r * bm + (h2 - l2):(r + 1) * bm] == 0)
            got_rows.append(seg[:h2 - l2]); got_dpre.append(dpre_all[r * bm:r * bm + (h2 - l2)])
        assert torch.equal(torch.cat(got_rows), rows_g) and torch.equal(torch.cat(got_dpre), dpre_g)

        # --- dense-gradient all-reduce and loss reduction: sum of per-rank partials == global value
        g_parts = torch.from_numpy(rng.standard_normal((world, 37)).astype(np.float32))
        gd = g_parts[rank].clone()
        dp.reduce_dense(gd)
        np.testing.assert_allclose(gd.numpy(), g_parts.sum(0).numpy(), rtol=1e-6)
        acc = torch.tensor([1.0 + rank, 2.0 + rank, 3.0 + rank, 7.0])
        dp.reduce_losses(acc)
        assert acc.tolist() == [3.0, 5.0, 7.0, 7.0]

        # --- item-sharded top-K: local top-K per shard -> all-gather -> merge == global top-K
        B, N, K = 9, 101, 7
        S = torch.from_numpy(rng.standard_normal((B, N)).astype(np.float32))
        S[:, ::5] = torch.round(S[:, ::5] * 2) / 2           # ties across shards
        shard = hd.ItemShard(N, world, rank)
        assert shard.ranges()[0][0] == 0 and shard.ranges()[-1][1] == N
        loc = S[:, shard.lo:shard.hi].numpy()
        lv = np.full((B, K), -np.inf, np.float32); li = np.full((B, K), -1, np.int32)
        for b in range(B):
            order = np.lexsort((-np.arange(shard.lo, shard.hi), -loc[b].astype(np.float64)))[:K]
            lv[b, :len(order)], li[b, :len(order)] = loc[b][order], order + shard.lo
        cv, ci = hd.gather_candidates(torch.from_numpy(lv), torch.from_numpy(li))
        assert cv.shape == (B, world * K)
        mv, mi = hd.merge_candidates_host(cv.numpy(), ci.numpy(), K)
        for b in range(B):
            ref = np.argsort(S[b].numpy(), kind="stable")[::-1][:K]      # (score desc, index desc)
            assert np.array_equal(mi[b], ref), (b, mi[b], ref)
            assert np.array_equal(mv[b], S[b].numpy()[ref])

        # --- sharded loader: same global batches on every rank, disjoint slices, global sizes attached
        from scipy.sparse import random as sprand
        m = sprand(50, 30, density=0.2, format="csr", random_state=1)
        torch.manual_seed(4)
        ld2 = ShardedCSRLoader(m, list(range(47)), 16, True, device="cpu", dp=dp)
        mine = [(b.rows.clone(), b.b_global, b.nnz_cap_global) for b in ld2]
        gathered = [None] * world
        dist.all_gather_object(gathered, [(r.tolist(), bg, cg) for r, bg, cg in mine])
        seen = []
        for step in range(len(mine)):
            rows_step = sum((gathered[r][step][0] for r in range(world)), [])
            assert len(rows_step) == gathered[0][step][1] == gathered[1][step][1]
            nnz = int(np.diff(m.indptr)[rows_step].sum())
            assert nnz <= gathered[0][step][2]
            seen += rows_step
        assert sorted(seen) == list(range(47))

        # --- the packed (value | id) candidate blocks of ShardedEvaluator: one all-gather of [2, B, K] per rank, merged in place
        #     with group stride 2*B*K (what hvae_topk_merge_groups reads) == the merge of the concatenated candidates
        send = torch.empty(2, B, K, dtype=torch.int32)
        send[0] = torch.from_numpy(lv).view(torch.int32)
        send[1] = torch.from_numpy(li)
        recv = torch.empty(world * 2, B, K, dtype=torch.int32)    # == [rank][2][B][K]
        dist.all_gather_into_tensor(recv, send)
        flat = recv.reshape(-1).numpy()
        gv = np.stack([np.concatenate([flat[g * 2 * B * K + b * K:g * 2 * B * K + b * K + K].view(np.float32) for g in range(world)])
                       for b in range(B)])
        gi = np.stack([np.concatenate([flat[g * 2 * B * K + B * K + b * K:g * 2 * B * K + B * K + b * K + K] for g in range(world)])
                       for b in range(B)])
        mv2, mi2 = hd.merge_candidates_host(gv, gi, K)
        assert np.array_equal(mi2, mi) and np.array_equal(mv2, mv)

        # --- the gradient-exchange path is decided once and collectively: no symmetric memory on CPU -> every rank says "nccl"
        class _Eng:
            dev = torch.device("cpu")
            class ws:
                generation = 0
        assert dp.prepare_exchange(_Eng(), 1024) == "nccl" and dp.exchange == "nccl"
        forced = hd.DataParallel()
        forced.exchange_mode = "nvl"
        try:
            forced.prepare_exchange(_Eng(), 1024)
            raise AssertionError("HVAE_DP_EXCHANGE=nvl must not fall back silently")
        except RuntimeError as e:
            assert "symmetric memory" in str(e)

        # --- independent trainings (grid sweep): every configuration placed exactly once
        mine_cfg = hd.assign_round_robin(16, world, rank)
        allc = [None] * world
        dist.all_gather_object(allc, mine_cfg)
        assert sorted(sum(allc, [])) == list(range(16))
        (Path(out_dir) / f"ok{rank}").write_text("ok")
    finally:
        dist.destroy_process_group()


def test_dist_host_logic_world2(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))
