"""Golden vectors from the REFERENCE ITSELF at the shapes bench.py times (run in the build container only).

    python oracle/make_golden_big.py        # writes tests/golden/{c2_train,d768_train,randn_forward,tune_val}.npz

Complements oracle/make_golden.py (tiny cases, everything stored) with cases whose INPUTS regenerate from seeds
(hvae_b200.synth + torch.manual_seed -> the reference's own initialisation) and whose OUTPUTS are stored as per-step scalars
plus per-tensor checksums / samples, so that the fixtures stay small:

  c2_train      BASELINE.json configs[1] shape (22,363 x 12,101, d=384, latent 200, hidden [600], dropout 0.5): two optimisation
                steps of the reference's VAETrainer (src/ml/train.py:71-96) with the noise it consumed, validate() on two batches.
  d768_train    512 users x 3,000 items, d=768, latent 200, hidden [600]: the d > 384 scoring kernels end to end
                (src/ml/model.py:198,281 + train.py:88-92), plus eval-mode scores of 8 users and the evaluator's top-20 of 64.
  randn_forward the reference's own unit-test input (tests/test_unit.py:153-172): dense torch.randn rows, negative values included.
  tune_val      evaluate_config_on_val (src/ml/tune.py:121-184) with the 99 negatives np.random.choice drew, captured per row.
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "recommendation-system_b200"), str(ROOT / "oracle")]
OUT = ROOT / "tests" / "golden"
N_SAMPLE = 48


def sample_positions(shape, seed):
    """Fixed pseudo-random flat positions inside a tensor of `shape` (shared by the generator and the tests)."""
    n = int(np.prod(shape))
    return np.random.default_rng(seed).integers(0, n, size=min(N_SAMPLE, n))


def tensor_digest(prefix, sd, out, seed=12345):
    """sum, sum of |.|, and N_SAMPLE sampled entries of every tensor of a state_dict-like mapping."""
    for i, (k, v) in enumerate(sd.items()):
        a = v.detach().cpu().double().numpy() if hasattr(v, "detach") else np.asarray(v, dtype=np.float64)
        out[f"{prefix}/{k}/sum"] = np.float64(a.sum())
        out[f"{prefix}/{k}/abs"] = np.float64(np.abs(a).sum())
        out[f"{prefix}/{k}/sample"] = a.reshape(-1)[sample_positions(a.shape, seed + i)]


def big_train_case(name, ref_model, ref_train, ref_eval, *, n_users, n_items, d, latent, hidden, dropout, beta, batches, seed,
                   store_eval=False):
    import torch
    from make_golden import _flat  # noqa: F401  (same helpers / reference import)
    from oracle import hvae_oracle as orc
    from hvae_b200.synth import make_interactions, make_item_embeddings

    data = make_interactions(n_users, n_items, seed)
    csr = data.scipy_csr()
    E = make_item_embeddings(n_items, d, seed)
    torch.manual_seed(seed)
    kw = dict(n_items=n_items, item_embeddings=E, latent_dim=latent, hidden_dims=hidden, dropout=dropout, beta=beta)
    model = ref_model.create_hybrid_vae(**kw)
    out = dict(n_users=n_users, n_items=n_items, d=d, latent=latent, hidden=np.array(hidden), dropout=dropout, beta=beta, seed=seed,
               steps=len(batches))
    tensor_digest("init", {k: v for k, v in model.state_dict().items() if k != "item_embeddings"}, out)
    trainer = ref_train.VAETrainer(model, torch.device("cpu"), lr=1e-3)
    shadow = orc.OracleVAE(**kw)
    order = np.random.default_rng(seed + 100).permutation(n_users)
    pos, stats = 0, []
    model.train()
    for s, B in enumerate(batches):
        rows = order[pos:pos + B]
        pos += B
        x = torch.from_numpy(np.asarray(csr[rows].toarray(), dtype=np.float32))
        state = torch.get_rng_state()
        noise = orc.draw_noise(shadow, len(rows))          # replay the reference's draw order ...
        torch.set_rng_state(state)                         # ... then let the reference consume the same draws
        trainer.optimizer.zero_grad()
        loss, recon, kl = trainer._compute_loss(x)
        loss.backward()
        gnorm = torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm=5.0)
        trainer.optimizer.step()
        stats.append([loss.item(), recon.item(), kl.item(), float(gnorm)])
        out[f"rows/{s}"] = rows.astype(np.int32)
        for i, m in enumerate(noise["masks"]):
            out[f"noise/{s}/mask{i}"] = np.packbits(m.numpy().astype(np.uint8), axis=1)
        out[f"noise/{s}/eps"] = noise["eps"].numpy()
        out[f"noise/{s}/pmask"] = np.packbits(noise["pmask"].numpy().astype(np.uint8), axis=1)
    out["stats"] = np.array(stats, dtype=np.float64)
    # the same steps in float64 (oracle module tree in double, same noise): at these sizes the reference's own fp32 grad norm
    # (torch's foreach norm over a 7M-element tensor) is ~3e-5 off the exact value, more than the kernels under test are
    torch.manual_seed(seed)
    o64 = orc.OracleVAE(**kw).double()
    opt64 = orc.make_adam(o64, 1e-3, 0.0)
    stats64 = []
    for s in range(len(batches)):
        rows = out[f"rows/{s}"]
        x64 = torch.from_numpy(np.asarray(csr[rows].toarray(), dtype=np.float64))
        n64 = dict(masks=[torch.from_numpy(np.unpackbits(out[f"noise/{s}/mask{i}"], axis=1)[:, :h].astype(np.float64))
                          for i, h in enumerate(hidden)], eps=torch.from_numpy(out[f"noise/{s}/eps"]).double(),
                   pmask=torch.from_numpy(np.unpackbits(out[f"noise/{s}/pmask"], axis=1)[:, :d].astype(np.float64)))
        stats64.append(list(orc.train_step(o64, opt64, x64, n64, beta)))
    out["stats64"] = np.array(stats64, dtype=np.float64)
    tensor_digest("final", {k: v for k, v in model.state_dict().items() if k != "item_embeddings"}, out)
    osd = trainer.optimizer.state_dict()["state"]
    names = [k for k, _ in model.named_parameters()]
    tensor_digest("adam_m", {k: osd[i]["exp_avg"] for i, k in enumerate(names)}, out)
    tensor_digest("adam_v", {k: osd[i]["exp_avg_sq"] for i, k in enumerate(names)}, out)

    vb = batches[0]

    class _DS(torch.utils.data.Dataset):
        def __len__(self):
            return min(n_users, 2 * vb)

        def __getitem__(self, i):
            return torch.FloatTensor(csr[i].toarray().flatten())

    val = trainer.validate(torch.utils.data.DataLoader(_DS(), batch_size=vb, shuffle=False))
    out["validate"] = np.array([val["total_loss"], val["recon_loss"], val["kl_loss"]], dtype=np.float64)
    if store_eval:
        model.eval()
        with torch.no_grad():
            x8 = torch.from_numpy(np.asarray(csr[:8].toarray(), dtype=np.float32))
            s8, mu8, lv8 = model(x8)
        out["fwd8/scores"], out["fwd8/mu"], out["fwd8/logvar"] = s8.numpy(), mu8.numpy(), lv8.numpy()
        u2i = {f"u{i:07d}": i for i in range(n_users)}
        i2i = {f"i{i:07d}": i for i in range(n_items)}
        ev = ref_eval.RecommendationEvaluator(model, csr, u2i, i2i, torch.device("cpu"))
        out["top20"] = np.stack([ev.get_user_recommendations(u, top_k=20)[0] for u in range(64)]).astype(np.int32)
        out["top20_scores"] = np.stack([ev.get_user_recommendations(u, top_k=20)[1] for u in range(64)]).astype(np.float32)
    np.savez_compressed(OUT / f"{name}.npz", **out)
    print(name, stats, out["validate"])


def randn_forward_case(ref_model):
    """tests/test_unit.py:60-74,153-198 of the reference: 50 items, 384-d mock embeddings (np.random.seed(42)), latent 64,
    hidden [128]; dense randn input.  Eval-mode outputs are deterministic and stored; shapes hold in train mode too."""
    import torch
    np.random.seed(42)
    emb = np.random.randn(50, 384).astype(np.float32)
    emb = emb / np.linalg.norm(emb, axis=1, keepdims=True)
    torch.manual_seed(0)
    model = ref_model.HybridVAE(n_items=50, item_embeddings=emb, latent_dim=64, hidden_dims=[128])
    x = torch.randn(4, 50)
    model.eval()
    with torch.no_grad():
        scores, mu, logvar = model(x)
        z = model.get_user_embedding(x)
        dec = model.decode(z)
        loss, recon, kl = ref_model.vae_loss_function(scores, x, mu, logvar, beta=0.2)
    out = dict(emb=emb, x=x.numpy(), scores=scores.numpy(), mu=mu.numpy(), logvar=logvar.numpy(), user_embedding=z.numpy(),
               decode=dec.numpy(), loss=np.array([loss.item(), recon.item(), kl.item()]))
    del out["emb"]                                     # regenerates from np.random.seed(42); the weights from torch.manual_seed(0)
    tensor_digest("init", {k: v for k, v in model.state_dict().items() if k != "item_embeddings"}, out)
    np.savez_compressed(OUT / "randn_forward.npz", **out)
    print("randn_forward", out["loss"])


def tune_val_case(ref_model):
    """evaluate_config_on_val of the reference (src/ml/tune.py:121-184) on a small seeded model; the negatives it draws with the
    unseeded np.random.choice are captured (np.random.seed(123) first) so that the same candidates can be replayed."""
    import pandas as pd
    import torch
    import ml.tune as ref_tune
    from hvae_b200.synth import make_interactions, make_item_embeddings

    n_users, n_items, d = 120, 260, 32
    data = make_interactions(n_users, n_items, 21)
    csr = data.scipy_csr()
    E = make_item_embeddings(n_items, d, 21)
    torch.manual_seed(21)
    model = ref_model.create_hybrid_vae(n_items=n_items, item_embeddings=E, latent_dim=16, hidden_dims=[40], dropout=0.3, beta=0.2)
    u2i = {f"u{i:07d}": i for i in range(n_users)}
    i2i = {f"i{i:07d}": i for i in range(n_items)}
    rows = list(range(0, n_users, 1))
    val_df = pd.DataFrame({"user_id": [f"u{i:07d}" for i in rows] + ["unknown_user"],
                           "asin": [f"i{int(data.test_items[i]):07d}" for i in rows] + ["i0000001"]})
    drawn = []
    real_choice = np.random.choice

    def recording_choice(a, size=None, replace=True, p=None):
        r = real_choice(a, size, replace, p)
        drawn.append(np.asarray(r).copy())
        return r

    np.random.seed(123)
    np.random.choice = recording_choice
    try:
        res = ref_tune.evaluate_config_on_val(model, csr, val_df, u2i, i2i, torch.device("cpu"), n_negatives=99, k_values=[5, 10])
    finally:
        np.random.choice = real_choice
    out = dict(n_users=n_users, n_items=n_items, d=d, seed=21, negatives=np.stack(drawn).astype(np.int32),
               keys=np.array(sorted(res)), values=np.array([res[k] for k in sorted(res)], dtype=np.float64))
    out.update({f"init/{k}": v.numpy().copy() for k, v in model.state_dict().items() if k != "item_embeddings"})
    np.savez_compressed(OUT / "tune_val.npz", **out)
    print("tune_val", res)


def main():
    from make_golden import _import_reference
    OUT.mkdir(parents=True, exist_ok=True)
    ref_model, ref_train, ref_eval = _import_reference()
    import torch
    torch.set_num_threads(8)
    from hvae_b200.synth import CONFIGS
    c = CONFIGS["c2"]
    big_train_case("c2_train", ref_model, ref_train, ref_eval, n_users=c["n_users"], n_items=c["n_items"], d=c["emb_dim"],
                   latent=c["latent_dim"], hidden=c["hidden_dims"], dropout=c["dropout"], beta=c["beta"], batches=[512, 509], seed=0)
    big_train_case("d768_train", ref_model, ref_train, ref_eval, n_users=512, n_items=3000, d=768, latent=200, hidden=[600],
                   dropout=0.5, beta=0.2, batches=[300, 212], seed=5, store_eval=True)
    randn_forward_case(ref_model)
    tune_val_case(ref_model)


if __name__ == "__main__":
    main()
