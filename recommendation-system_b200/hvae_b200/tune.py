"""Drop-in grid search: DEFAULT_SEARCH_SPACE, train_single_config, evaluate_config_on_val, run_grid_search
(reference src/ml/tune.py).

The reference walks `itertools.product(search_space)` sequentially (tune.py:241).  Configurations are independent
trainings, so under torch.distributed they are placed round-robin on the ranks (one process per GPU, no collective on
the data path); the per-configuration result dicts are gathered on every rank and rank 0 writes the same
`grid_search_results.json` (tune.py:297-310).  Each training is the fused / CUDA-graphed step of VAETrainer, the
validation NDCG@10 is the 99-negative protocol on the device (sampling.py).
"""
from __future__ import annotations

import itertools
import json
import logging
from datetime import datetime
from pathlib import Path
from typing import Any

import numpy as np
import torch

from . import data as _data
from . import sampling
from .evaluate import RecommendationEvaluator
from .model import create_hybrid_vae
from .train import CSRLoader, VAETrainer

logger = logging.getLogger(__name__)

DEFAULT_SEARCH_SPACE = {                       # src/ml/tune.py:33-39
    "latent_dim": [32, 64, 128],
    "hidden_dims": [[256], [512], [256, 128]],
    "dropout": [0.3, 0.5],
    "beta": [0.1, 0.2, 0.3],
    "learning_rate": [1e-3, 5e-4],
}


def train_single_config(model, train_loader, val_loader, device, learning_rate: float, epochs: int = 10,
                        patience: int = 3) -> tuple[float, int]:
    """(best validation loss, its epoch) with early stopping -- src/ml/tune.py:63-118."""
    trainer = VAETrainer(model, device, lr=learning_rate)
    best_val_loss, best_epoch, patience_counter = float("inf"), 0, 0
    for epoch in range(epochs):
        trainer.train_epoch(train_loader)
        val_loss = trainer.validate(val_loader)["total_loss"]
        if val_loss < best_val_loss:
            best_val_loss, best_epoch, patience_counter = val_loss, epoch + 1, 0
        else:
            patience_counter += 1
            if patience_counter >= patience:
                break
    return best_val_loss, best_epoch


def evaluate_pairs_with_negatives(model, matrix, users, items, device, n_negatives=99, k_values=None, seed=None, negatives=None):
    """99-negative metrics for (user index, held-out item index) pairs; scores from `matrix` rows."""
    k_values = k_values or [10]
    ev = RecommendationEvaluator(model, matrix, {}, {}, device)
    res = sampling.evaluate_with_negatives(ev, np.asarray(users), np.asarray(items), n_negatives, list(k_values), seed, negatives)
    return {f"{metric}@{k}": res[k][metric] for k in k_values for metric in ("recall", "ndcg", "hit_ratio")}


def evaluate_config_on_val(model, train_matrix, val_df, user_to_idx: dict, item_to_idx: dict, device, n_negatives: int = 99,
                           k_values: list[int] | None = None, seed=None, negatives=None) -> dict[str, float]:
    """src/ml/tune.py:121-184: one row per (user, held-out validation item), input = the user's train row.
    `negatives` [kept rows, n_negatives]: replay given draws (the reference's come from the unseeded np.random)."""
    users, items = [], []
    for user_id, item_id in zip(val_df["user_id"].values, val_df["asin"].values):
        if user_id in user_to_idx and item_id in item_to_idx:
            users.append(user_to_idx[user_id])
            items.append(item_to_idx[item_id])
    k_values = k_values or [10]
    if not users:
        return {f"{metric}@{k}": 0.0 for k in k_values for metric in ("recall", "ndcg", "hit_ratio")}
    return evaluate_pairs_with_negatives(model, train_matrix, users, items, device, n_negatives, k_values, seed, negatives)


def _dist():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist, dist.get_world_size(), dist.get_rank()
    return None, 1, 0


def grid_search_core(train_matrix, val_matrix, train_users, val_users, val_pairs, embeddings, search_space, epochs_per_config=10,
                     patience=3, batch_size=512, use_annealing=True, device="cuda", precision=None, seed=None) -> dict[str, Any]:
    """The loop of run_grid_search on in-memory inputs.  val_pairs = (user indices, item indices) of the validation rows.
    Under torch.distributed every rank trains configurations rank, rank+world, ... and all ranks return all results."""
    from .dist import assign_round_robin
    device = torch.device(device)
    n_items = train_matrix.shape[1]
    param_names = list(search_space.keys())
    all_configs = list(itertools.product(*search_space.values()))
    dist, world, rank = _dist()
    mine = assign_round_robin(len(all_configs), world, rank)
    train_loader = CSRLoader(train_matrix, train_users, batch_size, True, device)
    val_loader = CSRLoader(val_matrix, val_users, batch_size, False, device)
    local = {}
    for i in mine:
        cfg = dict(zip(param_names, all_configs[i]))
        try:
            model = create_hybrid_vae(n_items=n_items, item_embeddings=embeddings, latent_dim=cfg.get("latent_dim", 64),
                                      hidden_dims=cfg.get("hidden_dims", [256]), dropout=cfg.get("dropout", 0.5),
                                      beta=cfg.get("beta", 0.2), use_annealing=use_annealing,
                                      anneal_steps=len(train_loader) * epochs_per_config // 2, precision=precision)
            val_loss, best_epoch = train_single_config(model, train_loader, val_loader, device, cfg.get("learning_rate", 1e-3),
                                                       epochs_per_config, patience)
            metrics = evaluate_pairs_with_negatives(model, train_matrix, val_pairs[0], val_pairs[1], device, 99, [10], seed)
            local[i] = {"config": cfg, "val_loss": val_loss, "best_epoch": best_epoch, **metrics}
            logger.info("[%d/%d] %s: val %.4f ndcg@10 %.4f", i + 1, len(all_configs), cfg, val_loss, metrics["ndcg@10"])
            del model
        except Exception as e:                       # src/ml/tune.py:288-290: a failing configuration is recorded, not fatal
            logger.error("[%d/%d] %s failed: %s", i + 1, len(all_configs), cfg, e)
            local[i] = {"config": cfg, "error": str(e)}
    if dist is not None and world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, local)
        merged = {}
        for g in gathered:
            merged.update(g)
    else:
        merged = local
    results = [merged[i] for i in range(len(all_configs))]
    best_config, best_metric = None, -float("inf")
    for r in results:                                # first strictly better wins, in product order (tune.py:281-282)
        if "error" not in r and r["ndcg@10"] > best_metric:
            best_metric, best_config = r["ndcg@10"], r["config"]
    return {"best_config": best_config, "best_metric": best_metric, "all_results": results}


def run_grid_search(data_dir: str, embeddings_path: str, output_dir: str, search_space: dict[str, list] | None = None,
                    epochs_per_config: int = 10, patience: int = 3, batch_size: int = 512, use_annealing: bool = True,
                    device: str | None = None, precision: str | None = None) -> dict[str, Any]:
    """src/ml/tune.py:187-322: same inputs on disk, same grid_search_results.json."""
    search_space = search_space or DEFAULT_SEARCH_SPACE
    dev = torch.device(device) if device else torch.device("cuda")
    output_path = Path(output_dir)
    output_path.mkdir(parents=True, exist_ok=True)
    full_matrix, train_df, val_df, mappings = _data.load_training_data(data_dir)
    user_to_idx, item_to_idx = mappings["user_to_idx"], mappings["item_to_idx"]
    train_matrix = _data.build_matrix(train_df, user_to_idx, item_to_idx, full_matrix.shape)
    val_matrix = _data.build_matrix(val_df, user_to_idx, item_to_idx, full_matrix.shape)
    emb_path = Path(embeddings_path)
    embeddings, _, _ = _data.load_embeddings(embeddings_path, str(emb_path.with_name(f"{emb_path.stem}_mappings.pkl")))
    users, items = [], []
    for user_id, item_id in zip(val_df["user_id"].values, val_df["asin"].values):
        if user_id in user_to_idx and item_id in item_to_idx:
            users.append(user_to_idx[user_id])
            items.append(item_to_idx[item_id])
    out = grid_search_core(train_matrix, val_matrix, _data.get_user_indices_from_df(train_df, user_to_idx),
                           _data.get_user_indices_from_df(val_df, user_to_idx), (users, items), embeddings, search_space,
                           epochs_per_config, patience, batch_size, use_annealing, dev, precision)
    _, _, rank = _dist()
    if rank == 0:
        with open(output_path / "grid_search_results.json", "w") as f:
            json.dump({"search_space": {k: [str(v) for v in vals] for k, vals in search_space.items()},
                       "best_config": out["best_config"], "best_ndcg@10": out["best_metric"], "all_results": out["all_results"],
                       "timestamp": datetime.now().isoformat()}, f, indent=2, default=str)
        logger.info("Grid search complete: best %s (NDCG@10 %.4f)", out["best_config"], out["best_metric"])
    return out
