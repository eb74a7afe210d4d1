#!/usr/bin/env python
"""Full-catalogue evaluation at the C4 shape (BASELINE.json configs[3]: users scored against 1M items, d=768), item-sharded:
every rank scores its contiguous item shard with the fused tcgen05 GEMM + seen-mask + top-K kernel, the K candidates per
user are all-gathered (NCCL) and merged, Recall/NDCG/HR are reduced on the device.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 \
        tools/bench_eval_sharded.py [--items 1000000 --users 32768 --d 768 --K 20]
Weights and E are random (drawn on the device: the reference's CPU initialisation of a 1M x 600 layer is not the thing
measured); interactions are the synthetic generator of hvae_b200.synth.  Prints one JSON line on rank 0."""
import argparse
import json
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "recommendation-system_b200")]
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--items", type=int, default=1_000_000)
    ap.add_argument("--users", type=int, default=32768)
    ap.add_argument("--d", type=int, default=768)
    ap.add_argument("--K", type=int, default=20)
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from hvae_b200 import dist as hd
    from hvae_b200._cabi import p
    from hvae_b200.engine import Batch, DeviceCSR, Engine, Layout
    from hvae_b200.synth import make_interactions
    N, U, d, K = a.items, a.users, a.d, a.K
    lay = Layout(N, d, 200, [600])
    g = torch.Generator(device=dev).manual_seed(0)           # same weights on every rank
    arena = torch.randn(lay.n_params, device=dev, generator=g) * 0.02
    lay.view(arena, "encoder.1.weight").fill_(1.0)
    E = torch.nn.functional.normalize(torch.randn(N, d, device=dev, generator=g), dim=1)
    eng = Engine(lay, arena, E, 0.5, "bf16")
    data = make_interactions(U, N, 0)
    csr = DeviceCSR.from_arrays(data.indptr, data.indices, None, N, dev)
    shard = hd.ItemShard(N, world, rank)
    users = torch.arange(U, dtype=torch.int32, device=dev)
    rel_ptr = torch.arange(U + 1, dtype=torch.int64, device=dev)
    rel_idx = torch.from_numpy(data.test_items.astype(np.int32)).to(dev)
    from hvae_b200.evaluate import _metric_tables
    kvals = [5, 10, 20]
    disc, idcg = _metric_tables(K)
    t = lambda x, dt: torch.as_tensor(np.asarray(x), dtype=dt, device=dev)
    kv, dd, ii = t(kvals, torch.int32), t(disc, torch.float64), t(idcg, torch.float64)
    out = torch.zeros(len(kvals) * 3 + 1, dtype=torch.float64, device=dev)
    wsd = torch.empty(148 * (len(kvals) * 3 + 1), dtype=torch.float64, device=dev)
    topk_all = torch.empty(U, K, dtype=torch.int32, device=dev)
    mask = torch.empty(U, 4, dtype=torch.int32, device=dev)

    def run():
        with torch.no_grad():
            for s in range(0, U, a.batch):
                rows = users[s:s + a.batch]
                b = Batch(csr, rows, rows.shape[0], 1)
                v, i = hd.sharded_topk(eng, b, K, shard)
                topk_all[s:s + rows.shape[0]] = i
            eng.lib.hit_mask(p(topk_all), U, K, p(rel_ptr), p(rel_idx), p(mask), eng.stream)
            eng.lib.metrics_reduce(p(mask), p(rel_ptr), U, p(kv), len(kvals), p(dd), p(ii), p(wsd), p(out), eng.stream)

    run()
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.reps):
        run()
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / a.reps
    if world > 1:
        tt = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt[0])
    # cross-check against the un-sharded path on the first batch (same arithmetic per item -> identical ids)
    with torch.no_grad():
        rows = users[:min(U, 1024)]
        v1, i1 = eng.topk(Batch(csr, rows, rows.shape[0], 1), K)
    same = bool(torch.equal(i1, topk_all[:rows.shape[0]]))
    if rank == 0:
        o = out.cpu().numpy()
        pk = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {"bf16_tflops_sustained": 1400.0}
        flops = 2.0 * U * N * d
        print(json.dumps({"metric": "eval_topk_users_per_sec", "value": U / (ms * 1e-3), "unit": "users/s", "n_gpus": world, "users": U, "items": N,
                          "d": d, "K": K, "ms": ms, "sharding": "items" if world > 1 else "none", "items_per_rank": shard.hi - shard.lo,
                          "aggregate_tflops": flops / (ms * 1e-3) / 1e12, "frac_of_sustained_peak_per_gpu": flops / (ms * 1e-3) / 1e12 / world / pk["bf16_tflops_sustained"],
                          "ids_equal_unsharded": same, "ndcg@10": float(o[1 * 3 + 1] / max(o[-1], 1)), "timing": "CUDA events, max over ranks, encoder + scoring + all-gather + merge + metrics"}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    sys.exit(0 if same else 1)


if __name__ == "__main__":
    main()
