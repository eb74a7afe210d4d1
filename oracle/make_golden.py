"""Freeze golden vectors from the REFERENCE ITSELF (run in the build container only).

    python oracle/make_golden.py            # writes tests/golden/*.npz

Imports the unmodified reference classes from /root/reference (with a stub for the
absent `sentence_transformers` package, which the hot path never calls), drives them on
small seeded inputs, and stores inputs + outputs as .npz.  The GPU box has no
/root/reference; tests there read only the committed .npz files.

What is recorded, per case:
  * the reference's initial state_dict (so init parity of the oracle can be checked),
  * the CSR batches, the noise tensors the reference consumed (recovered by replaying its
    draw order from the same generator state -- see oracle/hvae_oracle.py:draw_noise),
  * per-step loss / recon / kl / pre-clip grad norm from the reference's own
    VAETrainer._compute_loss + backward + clip_grad_norm_ + Adam (src/ml/train.py:71-96),
  * all gradients of step 0, the state_dict after the last step,
  * validate() numbers (src/ml/train.py:105-124), full-ranking top-K and metrics from the
    reference's RecommendationEvaluator (src/ml/evaluate.py:137-147,217-265).
"""

from __future__ import annotations

import os
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
REF = Path("/root/reference")
OUT = ROOT / "tests" / "golden"
sys.path[:0] = [str(ROOT), str(ROOT / "recommendation-system_b200")]


def _import_reference():
    os.environ["PYTHONDONTWRITEBYTECODE"] = "1"
    sys.dont_write_bytecode = True
    stub = Path(tempfile.mkdtemp(prefix="st_stub_")) / "sentence_transformers"
    stub.mkdir(parents=True)
    (stub / "__init__.py").write_text("class SentenceTransformer:\n    pass\n")
    sys.path[:0] = [str(stub.parent), str(REF), str(REF / "src")]
    cwd = os.getcwd()
    os.chdir(tempfile.mkdtemp(prefix="ref_cwd_"))  # src.config mkdirs relative dirs
    try:
        import ml.evaluate as ref_eval
        import ml.model as ref_model
        import ml.train as ref_train
    finally:
        os.chdir(cwd)
    return ref_model, ref_train, ref_eval


def _flat(prefix, sd):
    return {f"{prefix}/{k}": v.detach().cpu().numpy().copy() for k, v in sd.items()}


def _csr_case(n_users, n_items, seed, dup=False, empty_row=None):
    from hvae_b200.synth import make_interactions

    d = make_interactions(n_users, n_items, seed)
    vals = d.values.copy()
    if dup:  # csr_matrix sums duplicate (user,item) pairs -> values 2.0, 3.0 (train.py:182)
        vals[::7] = 2.0
        vals[::31] = 3.0
    indptr, indices = d.indptr.copy(), d.indices.copy()
    if empty_row is not None:
        s, e = indptr[empty_row], indptr[empty_row + 1]
        indices = np.concatenate([indices[:s], indices[e:]])
        vals = np.concatenate([vals[:s], vals[e:]])
        indptr[empty_row + 1:] -= (e - s)
    return indptr, indices, vals, d.test_items


def train_case(name, ref_model, ref_train, *, n_users, n_items, d, latent, hidden, dropout, beta, batch,
               steps, seed, annealing=False, dup=False, empty_row=None, weight_decay=0.0):
    import torch
    from scipy.sparse import csr_matrix

    from oracle import hvae_oracle as orc
    from hvae_b200.synth import make_item_embeddings

    indptr, indices, vals, test_items = _csr_case(n_users, n_items, seed, dup, empty_row)
    csr = csr_matrix((vals.astype(np.float64), indices, indptr), shape=(n_users, n_items))
    E = make_item_embeddings(n_items, d, seed)

    torch.manual_seed(seed)
    kw = dict(n_items=n_items, item_embeddings=E, latent_dim=latent, hidden_dims=hidden, dropout=dropout,
              beta=beta)
    model = ref_model.create_hybrid_vae(use_annealing=annealing, anneal_steps=4, **kw) if annealing \
        else ref_model.create_hybrid_vae(**kw)
    out = dict(n_users=n_users, n_items=n_items, d=d, latent=latent, hidden=np.array(hidden), dropout=dropout,
               beta=beta, batch=batch, steps=steps, seed=seed, annealing=int(annealing),
               weight_decay=weight_decay, indptr=indptr, indices=indices, values=vals, test_items=test_items)
    out.update(_flat("init", model.state_dict()))

    trainer = ref_train.VAETrainer(model, torch.device("cpu"), lr=1e-3, weight_decay=weight_decay)
    # a shadow oracle model, only to size the noise replay
    shadow = orc.OracleVAE(**kw)
    rows_per_step, stats = [], []
    model.train()
    for s in range(steps):
        rows = np.arange(s * batch, (s + 1) * batch) % n_users
        if s == steps - 1:
            rows = rows[: max(1, len(rows) - 3)]  # ragged last batch
        rows_per_step.append(rows)
        x = torch.from_numpy(np.asarray(csr[rows].toarray(), dtype=np.float32))
        state = torch.get_rng_state()
        noise = orc.draw_noise(shadow, len(rows))          # replay ...
        torch.set_rng_state(state)                         # ... then let the reference consume the same draws
        trainer.optimizer.zero_grad()
        loss, recon, kl = trainer._compute_loss(x)         # reference code path incl. annealing
        loss.backward()
        gnorm = torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm=5.0)
        if s == 0:
            # grads are post-clip here; store the clip coefficient's inputs too
            out.update({f"grad0/{k}": p.grad.detach().numpy().copy() for k, p in model.named_parameters()})
        trainer.optimizer.step()
        stats.append([loss.item(), recon.item(), kl.item(), float(gnorm)])
        out[f"rows/{s}"] = rows
        for i, m in enumerate(noise["masks"]):
            out[f"noise/{s}/mask{i}"] = m.numpy().astype(np.uint8)
        out[f"noise/{s}/eps"] = noise["eps"].numpy()
        if noise["pmask"] is not None:
            out[f"noise/{s}/pmask"] = noise["pmask"].numpy().astype(np.uint8)
    out["stats"] = np.array(stats, dtype=np.float64)
    out.update(_flat("final", model.state_dict()))
    osd = trainer.optimizer.state_dict()["state"]
    names = [k for k, _ in model.named_parameters()]
    for i, k in enumerate(names):
        out[f"adam_m/{k}"] = osd[i]["exp_avg"].numpy()
        out[f"adam_v/{k}"] = osd[i]["exp_avg_sq"].numpy()

    # validate(): eval mode, fixed model.beta (train.py:105-124) over two loader batches
    class _DS(torch.utils.data.Dataset):
        def __len__(self):
            return min(n_users, 2 * batch)

        def __getitem__(self, i):
            return torch.FloatTensor(csr[i].toarray().flatten())

    val = trainer.validate(torch.utils.data.DataLoader(_DS(), batch_size=batch, shuffle=False))
    out["validate"] = np.array([val["total_loss"], val["recon_loss"], val["kl_loss"]], dtype=np.float64)

    # forward() in eval mode on the first 8 rows: scores, mu, logvar (model.py:202-221)
    model.eval()
    with torch.no_grad():
        x8 = torch.from_numpy(np.asarray(csr[:8].toarray(), dtype=np.float32))
        s8, mu8, lv8 = model(x8)
    out["fwd8/scores"], out["fwd8/mu"], out["fwd8/logvar"] = s8.numpy(), mu8.numpy(), lv8.numpy()
    np.savez_compressed(OUT / f"{name}.npz", **out)
    print(name, "steps", stats)
    return model, csr, test_items


def eval_case(name, ref_eval, model, csr, test_items, k_values=(5, 10, 20), n_rec_users=6):
    import pandas as pd
    import torch

    n_users, n_items = csr.shape
    user_to_idx = {f"u{i:07d}": i for i in range(n_users)}
    item_to_idx = {f"i{i:07d}": i for i in range(n_items)}
    ev = ref_eval.RecommendationEvaluator(model, csr, user_to_idx, item_to_idx, torch.device("cpu"))
    test_df = pd.DataFrame({"user_id": [f"u{i:07d}" for i in range(n_users)],
                            "asin": [f"i{int(t):07d}" for t in test_items]})
    res = ev.evaluate_dataset(test_df, list(k_values))
    kmax = max(k_values)
    tops = np.stack([ev.get_user_recommendations(u, top_k=kmax)[0] for u in range(n_users)])
    out = dict(k_values=np.array(k_values), topk=tops.astype(np.int32),
               metrics=np.array([[res[k][m] for m in ("recall", "ndcg", "hit_ratio")] for k in k_values]))
    for u in range(n_rec_users):
        idx, sc = ev.get_user_recommendations(u, top_k=100)
        out[f"rec100/{u}/idx"], out[f"rec100/{u}/score"] = idx.astype(np.int32), sc
        idx, sc = ev.get_user_recommendations(u, top_k=10, exclude_seen=False)
        out[f"rec10_all/{u}/idx"], out[f"rec10_all/{u}/score"] = idx.astype(np.int32), sc
    np.savez_compressed(OUT / f"{name}.npz", **out)
    print(name, res)


def c1_case(ref_model, ref_eval):
    """C1-shaped (2,072 x 890, d=384, L=128, h=[512]) eval of the untrained reference model under
    torch.manual_seed(0); inputs regenerate from seeds, only outputs + weight checksums are stored."""
    import torch
    from hvae_b200.synth import CONFIGS, make_interactions, make_item_embeddings

    c = CONFIGS["c1"]
    data = make_interactions(c["n_users"], c["n_items"], 0)
    E = make_item_embeddings(c["n_items"], c["emb_dim"], 0)
    torch.manual_seed(0)
    model = ref_model.HybridVAE(n_items=c["n_items"], item_embeddings=E, latent_dim=c["latent_dim"],
                                hidden_dims=c["hidden_dims"], dropout=c["dropout"], beta=c["beta"])
    csr = data.scipy_csr()
    import pandas as pd

    users = np.arange(400)
    user_to_idx = {f"u{i:07d}": i for i in range(c["n_users"])}
    item_to_idx = {f"i{i:07d}": i for i in range(c["n_items"])}
    ev = ref_eval.RecommendationEvaluator(model, csr, user_to_idx, item_to_idx, torch.device("cpu"))
    test_df = pd.DataFrame({"user_id": [f"u{i:07d}" for i in users],
                            "asin": [f"i{int(data.test_items[i]):07d}" for i in users]})
    res = ev.evaluate_dataset(test_df, [5, 10, 20])
    tops = np.stack([ev.get_user_recommendations(int(u), top_k=20)[0] for u in users])
    sums = {k: np.float64(v.double().sum().item()) for k, v in model.state_dict().items()}
    np.savez_compressed(OUT / "c1_eval.npz", users=users, topk=tops.astype(np.int32),
                        metrics=np.array([[res[k][m] for m in ("recall", "ndcg", "hit_ratio")] for k in (5, 10, 20)]),
                        **{f"sum/{k}": v for k, v in sums.items()})
    print("c1_eval", res)


def main():
    OUT.mkdir(parents=True, exist_ok=True)
    ref_model, ref_train, ref_eval = _import_reference()
    import torch

    torch.set_num_threads(1)  # fixed reduction order for the frozen numbers
    m, csr, t = train_case("tiny_two_hidden", ref_model, ref_train, n_users=150, n_items=300, d=32, latent=16,
                           hidden=[48, 24], dropout=0.5, beta=0.2, batch=64, steps=3, seed=7)
    eval_case("tiny_two_hidden_eval", ref_eval, m, csr, t)
    m, csr, t = train_case("tiny_identity", ref_model, ref_train, n_users=90, n_items=211, d=24, latent=24,
                           hidden=[40], dropout=0.0, beta=0.35, batch=32, steps=3, seed=11, dup=True,
                           empty_row=5, weight_decay=0.01)
    eval_case("tiny_identity_eval", ref_eval, m, csr, t, k_values=(1, 5, 50))
    train_case("tiny_annealed", ref_model, ref_train, n_users=130, n_items=257, d=40, latent=12, hidden=[36],
               dropout=0.3, beta=0.2, batch=48, steps=6, seed=3, annealing=True)
    c1_case(ref_model, ref_eval)


if __name__ == "__main__":
    main()
