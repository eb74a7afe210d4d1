#!/usr/bin/env python
"""Multi-GPU parity check (run under torchrun, one rank per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py
1. data-parallel training step (NCCL all-reduce of dense grads + all-gather of dH1 / user ids) == the single-GPU step
   on the same global batch and noise, for fp32 and bf16 modes;
2. item-sharded evaluation (local top-K, all-gather, merge) == single-GPU top-K (bit-exact ids in fp32 mode).
Prints one JSON line per check on rank 0; exits non-zero on mismatch."""
import json
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "recommendation-system_b200"), str(ROOT / "tests")]
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from golden_util import Case
    from hvae_b200 import dist as hd
    from hvae_b200.engine import Batch, DeviceCSR
    from hvae_b200.model import create_hybrid_vae
    from hvae_b200.train import VAETrainer
    ok = True
    c = Case("tiny_two_hidden")
    u8 = lambda t: None if t is None else t.to(torch.uint8).to(dev).contiguous()
    for precision, tol in (("fp32", 2e-6), ("bf16", 2e-3)):
        def build():
            m = create_hybrid_vae(**c.model_kwargs(), precision=precision)
            m.load_state_dict(c.state("init"))
            return m.to(dev)
        single, multi = build(), build()
        tr1 = VAETrainer(single, dev, use_cuda_graph=False)
        tr2 = VAETrainer(multi, dev, use_cuda_graph=False)
        dp = tr2.enable_data_parallel()
        csr = DeviceCSR.from_scipy(c.csr, dev)
        single.train(); multi.train()
        worst = 0.0
        for s in range(c.steps):
            rows = c.rows(s)
            n = c.noise(s)
            noise = dict(masks=[u8(m) for m in n["masks"]], eps=n["eps"].to(dev).contiguous(), pmask=u8(n["pmask"]))
            rows_d = torch.tensor(rows, dtype=torch.int32, device=dev)
            tr1.train_step(csr.batch(rows_d, rows), noise)
            ref = np.array(tr1.last_losses())
            lo, hi = hd.split_even(len(rows), world, rank)
            nl = dict(masks=[m[lo:hi].contiguous() for m in noise["masks"]], eps=noise["eps"][lo:hi].contiguous(),
                      pmask=None if noise["pmask"] is None else noise["pmask"][lo:hi].contiguous())
            cap_g = int(np.diff(c.csr.indptr)[rows].sum())
            b = Batch(csr, rows_d[lo:hi].contiguous(), hi - lo, max(1, int(np.diff(c.csr.indptr)[rows[lo:hi]].sum())), b_global=len(rows),
                      nnz_cap_global=cap_g)
            tr2.train_step(b, nl, b_global=len(rows))
            part = multi.engine.loss_out.clone()
            dist.all_reduce(part)
            got = part.cpu().numpy()
            worst = max(worst, float(np.max(np.abs(got - ref) / np.abs(ref))))
        sd1, sd2 = single.state_dict(), multi.state_dict()
        pdiff = max(float((sd1[k] - sd2[k]).abs().max() / (sd1[k].abs().max() + 1e-12)) for k in sd1)
        good = worst < tol and pdiff < (1e-5 if precision == "fp32" else 5e-2)
        ok &= good
        if rank == 0:
            print(json.dumps({"check": "dp_train_step", "precision": precision, "world": world, "max_rel_loss_diff": worst,
                              "max_rel_param_diff": pdiff, "ok": good}), flush=True)
        # item-sharded evaluation
        single.eval()
        eng = single.engine
        users = torch.arange(c.n_users, dtype=torch.int32, device=dev)
        batch = Batch(csr, users, c.n_users, 1)
        K = 10
        v1, i1 = eng.topk(batch, K)
        shard = hd.ItemShard(c.n_items, world, rank)
        v2, i2 = hd.sharded_topk(eng, batch, K, shard)
        same = bool(torch.equal(i1, i2)) if precision == "fp32" else float((i1 == i2).float().mean()) > 0.97
        vclose = bool(torch.allclose(v1, v2, rtol=1e-5, atol=1e-6)) if precision == "fp32" else True
        ok &= same and vclose
        if rank == 0:
            print(json.dumps({"check": "item_sharded_topk", "precision": precision, "world": world, "ids_equal": same,
                              "values_close": vclose, "shard": [shard.lo, shard.hi]}), flush=True)
    flag = torch.tensor([0 if ok else 1], device=dev)
    dist.all_reduce(flag)
    dist.destroy_process_group()
    sys.exit(int(flag.item() != 0))


if __name__ == "__main__":
    main()
