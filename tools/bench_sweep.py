#!/usr/bin/env python
"""Grid-search sweep at the C2 shape (BASELINE.json configs[4]): 16 (latent_dim, hidden_dims, dropout, beta) configurations,
10 epochs each with early stopping (patience 3), validation NDCG@10 under the 99-negative protocol -- the reference's
run_grid_search loop (src/ml/tune.py:187-322) with the configurations placed round-robin on the ranks.

    python tools/bench_sweep.py                                   # 1 GPU: the 16 trainings run one after the other
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29530 tools/bench_sweep.py
Prints one JSON line on rank 0 (wall clock of the whole sweep, max over ranks)."""
import argparse
import json
import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "recommendation-system_b200")]
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--epochs", type=int, default=10)
    ap.add_argument("--cpu-reference-steps", type=int, default=0, help="time this many reference (oracle) steps of one configuration on the host")
    a = ap.parse_args()
    world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from hvae_b200.synth import CONFIGS, make_interactions, make_item_embeddings
    from hvae_b200.tune import grid_search_core
    c = CONFIGS["c2"]
    data = make_interactions(c["n_users"], c["n_items"], 0)
    E = make_item_embeddings(c["n_items"], c["emb_dim"], 0)
    train = data.scipy_csr()
    from scipy.sparse import csr_matrix
    U = c["n_users"]
    val = csr_matrix((np.ones(U), (np.arange(U), data.test_items)), shape=train.shape)      # 1-hot validation rows (train.py:232,255)
    space = {"latent_dim": [64, 128], "hidden_dims": [[256], [512]], "dropout": [0.3, 0.5], "beta": [0.1, 0.2], "learning_rate": [1e-3]}
    torch.manual_seed(0)
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    out = grid_search_core(train, val, list(range(U)), list(range(U)), (np.arange(U), data.test_items.astype(np.int64)), E, space,
                           epochs_per_config=a.epochs, patience=3, batch_size=512, use_annealing=True, device=dev, seed=0)
    torch.cuda.synchronize(dev)
    dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t[0])
    if rank == 0:
        ok = [r for r in out["all_results"] if "error" not in r]
        epochs_run = sum(min(a.epochs, r["best_epoch"] + 3) for r in ok)
        line = {"metric": "grid_sweep_seconds", "value": dt, "unit": "s", "n_gpus": world, "configs": len(out["all_results"]), "failed": len(out["all_results"]) - len(ok),
                "epochs_per_config": a.epochs, "users": U, "items": c["n_items"], "best_config": out["best_config"], "best_ndcg@10": out["best_metric"],
                "configs_per_second": len(ok) / dt, "approx_epochs_trained": epochs_run, "placement": "round-robin, one process per GPU, no data-path collective"}
        if a.cpu_reference_steps:
            from oracle import hvae_oracle as orc
            torch.set_num_threads(os.cpu_count() or 1)
            m = orc.OracleVAE(c["n_items"], E, 128, [512], 0.3, 0.2)
            n, secs, _ = orc.cpu_train_users(m, train, np.arange(a.cpu_reference_steps * 512), 512)
            steps_per_cfg = a.epochs * ((U + 511) // 512)
            line["cpu_reference"] = {"kind": "port", "cores": torch.get_num_threads(), "seconds_per_step": secs / a.cpu_reference_steps,
                                     "extrapolated_sweep_seconds_training_only": secs / a.cpu_reference_steps * steps_per_cfg * len(out["all_results"]),
                                     "sample": f"{a.cpu_reference_steps} steps of one configuration (latent 128, hidden [512])"}
        print(json.dumps(line, default=str), flush=True)
    if world > 1:
        torch.cuda.synchronize(dev)
        sys.stdout.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
