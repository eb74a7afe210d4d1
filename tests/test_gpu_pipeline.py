"""GPU: file-level drop-in of the reference's entry points on a synthetic data directory laid out as the reference's
offline pipeline writes it (src/preprocessing/dataset.py:137-179, embeddings.py:93-131): train_hybrid_vae
(src/ml/train.py:201-342), evaluate_recommendation_model (src/ml/evaluate.py:294-340, both protocols) and
run_grid_search (src/ml/tune.py:187-322).  Mirrors the reference's tests/test_e2e_pipeline.py at file granularity."""
import json
import pickle

import numpy as np
import pandas as pd
import pytest
import torch
from scipy.sparse import csr_matrix

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def data_dir(tmp_path_factory):
    from hvae_b200.synth import make_interactions, make_item_embeddings
    root = tmp_path_factory.mktemp("pipeline")
    U, N, d = 300, 200, 32
    s = make_interactions(U, N, 7)
    users = [f"U{u:05d}" for u in range(U)]
    items = [f"B{i:06d}" for i in range(N)]
    rows = []
    rng = np.random.default_rng(0)
    val_rows, test_rows = [], []
    for u in range(U):
        its = s.indices[s.indptr[u]:s.indptr[u + 1]]
        its = rng.permutation(its)
        test_rows.append((users[u], items[its[0]], 1))            # leave-one-out: last -> test, one -> val, rest -> train
        val_rows.append((users[u], items[its[1]], 1))
        rows += [(users[u], items[i], 1) for i in its[2:]]
        rows += [(users[u], items[int(rng.integers(0, N))], 0) for _ in range(2)]   # sampled negatives rows (binary_rating 0)
    cols = ["user_id", "asin", "binary_rating"]
    pd.DataFrame(rows, columns=cols).to_csv(root / "train.csv", index=False)
    pd.DataFrame(val_rows, columns=cols).to_csv(root / "val.csv", index=False)
    pd.DataFrame(test_rows, columns=cols).to_csv(root / "test.csv", index=False)
    full = csr_matrix((np.ones(len(s.indices)), s.indices, s.indptr), shape=(U, N))
    with open(root / "interaction_matrix.pkl", "wb") as f:
        pickle.dump(full, f)
    u2i, i2i = {u: k for k, u in enumerate(users)}, {it: k for k, it in enumerate(items)}
    with open(root / "mappings.pkl", "wb") as f:
        pickle.dump({"user_to_idx": u2i, "item_to_idx": i2i, "idx_to_user": {k: u for u, k in u2i.items()},
                     "idx_to_item": {k: it for it, k in i2i.items()}}, f)
    np.save(root / "item_embeddings.npy", make_item_embeddings(N, d, 7))
    with open(root / "item_embeddings_mappings.pkl", "wb") as f:
        pickle.dump({"item_to_idx": i2i, "idx_to_item": {k: it for it, k in i2i.items()}}, f)
    return root


def test_train_evaluate_tune_from_files(data_dir, tmp_path):
    from hvae_b200.evaluate import evaluate_recommendation_model
    from hvae_b200.train import train_hybrid_vae
    from hvae_b200.tune import run_grid_search
    out = tmp_path / "models"
    torch.manual_seed(0)
    train_hybrid_vae(str(data_dir), str(data_dir / "item_embeddings.npy"), str(out), latent_dim=16, hidden_dims=[48], batch_size=64,
                     epochs=3, beta=0.2, dropout=0.3, use_annealing=True, patience=5, device="cuda")
    assert (out / "best_model.pth").exists() and (out / "checkpoint_epoch_3.pth").exists()
    hist = json.loads((out / "training_history.json").read_text())
    assert len(hist["train_losses"]) == 3 and all(np.isfinite(hist["train_losses"])) and hist["train_losses"][-1] < hist["train_losses"][0]
    ck = torch.load(out / "best_model.pth", map_location="cpu", weights_only=False)
    assert ck["model_config"]["latent_dim"] == 16 and "encoder.0.weight" in ck["model_state_dict"]
    assert tuple(ck["model_state_dict"]["encoder.0.weight"].shape) == (48, 200)
    full = evaluate_recommendation_model(str(out / "best_model.pth"), str(data_dir), str(data_dir / "item_embeddings.npy"),
                                         k_values=[5, 10], device="cuda", n_negatives=None)
    neg = evaluate_recommendation_model(str(out / "best_model.pth"), str(data_dir), str(data_dir / "item_embeddings.npy"),
                                        k_values=[5, 10], device="cuda", n_negatives=99)
    for res in (full, neg):
        assert set(res) == {5, 10}
        for k in (5, 10):
            assert 0.0 <= res[k]["ndcg"] <= res[k]["hit_ratio"] <= 1.0 and res[k]["recall"] == res[k]["hit_ratio"]
    assert neg[10]["hit_ratio"] >= full[10]["hit_ratio"]          # 100 candidates are easier than 200 items
    space = {"latent_dim": [8, 16], "hidden_dims": [[32]], "dropout": [0.3], "beta": [0.2], "learning_rate": [1e-3]}
    gs = run_grid_search(str(data_dir), str(data_dir / "item_embeddings.npy"), str(tmp_path / "tune"), search_space=space,
                         epochs_per_config=2, patience=2, batch_size=64, device="cuda")
    assert len(gs["all_results"]) == 2 and all("error" not in r for r in gs["all_results"])
    assert gs["best_config"] in [r["config"] for r in gs["all_results"]]
    assert gs["best_metric"] == max(r["ndcg@10"] for r in gs["all_results"])
    saved = json.loads((tmp_path / "tune" / "grid_search_results.json").read_text())
    assert {"search_space", "best_config", "best_ndcg@10", "all_results", "timestamp"} <= set(saved)
    assert {"config", "val_loss", "best_epoch", "recall@10", "ndcg@10", "hit_ratio@10"} <= set(saved["all_results"][0])
