"""Golden vectors of the reference's Mult-VAE baseline (src/ml/baseline.py:126-231), run in the build container only.

    python oracle/make_golden_multvae.py     # writes tests/golden/multvae.npz

Drives the unmodified reference MultVAE through three optimisation steps of MultVAERecommender.fit's loop body
(baseline.py:196-204) on a small seeded matrix with duplicate-summed values, records the noise it consumed (replayed from the
same generator state: input dropout over the dense row, the two hidden dropouts, eps), the per-step losses, the final
state_dict and Adam moments, and eval-mode predictions of six users (the recommender's predict(), baseline.py:224-231)."""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "recommendation-system_b200"), str(ROOT / "oracle")]
OUT = ROOT / "tests" / "golden"


def main():
    from make_golden import _csr_case, _flat, _import_reference
    _import_reference()
    import ml.baseline as ref_base
    import torch
    from scipy.sparse import csr_matrix
    from oracle import hvae_oracle as orc
    torch.set_num_threads(1)
    n_users, n_items, h, L, pdrop, beta, batch, steps, seed = 96, 300, 48, 16, 0.5, 0.2, 40, 3, 13
    indptr, indices, vals, test_items = _csr_case(n_users, n_items, seed, dup=True)
    csr = csr_matrix((vals.astype(np.float64), indices, indptr), shape=(n_users, n_items))
    torch.manual_seed(seed)
    model = ref_base.MultVAE(n_items, h, L, dropout=pdrop)
    out = dict(n_users=n_users, n_items=n_items, hidden=h, latent=L, dropout=pdrop, beta=beta, batch=batch, steps=steps, seed=seed,
               indptr=indptr, indices=indices, values=vals)
    out.update(_flat("init", model.state_dict()))
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    shadow = orc.OracleMultVAE(n_items, h, L, pdrop)
    model.train()
    stats = []
    for s in range(steps):
        rows = np.arange(s * batch, (s + 1) * batch) % n_users
        if s == steps - 1:
            rows = rows[:-5]
        x = torch.from_numpy(np.asarray(csr[rows].toarray(), dtype=np.float32))
        state = torch.get_rng_state()
        noise = orc.draw_multvae_noise(shadow, len(rows), n_items, h, L)
        torch.set_rng_state(state)
        opt.zero_grad()                                                       # baseline.py:196-204, verbatim semantics
        recon, mu, logvar = model(x)
        recon_loss = -torch.mean(torch.sum(torch.nn.functional.log_softmax(recon, dim=1) * x, dim=1))
        kl_loss = -0.5 * torch.mean(torch.sum(1 + logvar - mu.pow(2) - logvar.exp(), dim=1))
        loss = recon_loss + beta * kl_loss
        loss.backward()
        opt.step()
        stats.append([loss.item(), recon_loss.item(), kl_loss.item()])
        out[f"rows/{s}"] = rows
        # input dropout: only the flags at the row's non-zeros matter; stored in batch (CSR) order
        keep = noise["in_mask"].numpy().astype(np.uint8)
        sub = csr[rows]
        out[f"noise/{s}/keep"] = keep[np.repeat(np.arange(len(rows)), np.diff(sub.indptr)), sub.indices]
        out[f"noise/{s}/mask0"] = noise["masks"][0].numpy().astype(np.uint8)
        out[f"noise/{s}/mask1"] = noise["masks"][1].numpy().astype(np.uint8)
        out[f"noise/{s}/eps"] = noise["eps"].numpy()
    out["stats"] = np.array(stats, dtype=np.float64)
    out.update(_flat("final", model.state_dict()))
    osd = opt.state_dict()["state"]
    for i, (k, _) in enumerate(model.named_parameters()):
        out[f"adam_m/{k}"] = osd[i]["exp_avg"].numpy()
        out[f"adam_v/{k}"] = osd[i]["exp_avg_sq"].numpy()
    model.eval()
    with torch.no_grad():
        x6 = torch.from_numpy(np.asarray(csr[:6].toarray(), dtype=np.float32))
        s6, mu6, lv6 = model(x6)
    out["pred6/scores"], out["pred6/mu"], out["pred6/logvar"] = s6.numpy(), mu6.numpy(), lv6.numpy()
    np.savez_compressed(OUT / "multvae.npz", **out)
    print("multvae", stats)


if __name__ == "__main__":
    main()
