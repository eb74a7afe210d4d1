"""CPU oracle for the HybridVAE hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A restatement, on PyTorch-CPU / NumPy, of what the reference computes on its
training / validation / full-ranking path.  Only `tests/`,
`__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of
`bench.py` may import this module; nothing under `recommendation-system_b200/`
does.

Parity status: PINNED.  `oracle/make_golden.py` runs the *reference's own*
classes (imported from /root/reference in the build container) on seeded
inputs and freezes the results under `tests/golden/`; `tests/test_oracle_golden.py`
checks every function here against those files.

Each function cites the reference lines it restates (paths relative to
/root/reference).
"""

from __future__ import annotations

import time

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


# --------------------------------------------------------------------------------------
# model (src/ml/model.py:35-136)
# --------------------------------------------------------------------------------------
class OracleVAE(nn.Module):
    """Same module tree, hence the same state_dict keys and the same RNG draw order at
    construction, as the reference HybridVAE (src/ml/model.py:57-101)."""

    def __init__(self, n_items, item_embeddings, latent_dim=200, hidden_dims=None, dropout=0.5, beta=0.2):
        super().__init__()
        hidden_dims = [600, 200] if hidden_dims is None else list(hidden_dims)  # model.py:65-66
        self.n_items, self.latent_dim, self.dropout, self.beta = n_items, latent_dim, dropout, beta
        self.hidden_dims = hidden_dims
        self.embedding_dim = item_embeddings.shape[1]
        self.register_buffer("item_embeddings", torch.tensor(np.asarray(item_embeddings), dtype=torch.float32))
        blocks, fan_in = [], n_items
        for h in hidden_dims:                                                  # model.py:111-120
            blocks += [nn.Linear(fan_in, h), nn.LayerNorm(h), nn.GELU(), nn.Dropout(dropout)]
            fan_in = h
        self.encoder = nn.Sequential(*blocks)
        self.fc_mu = nn.Linear(fan_in, latent_dim)                             # model.py:126-127
        self.fc_logvar = nn.Linear(fan_in, latent_dim)
        if latent_dim != self.embedding_dim:                                   # model.py:89-98
            self.projection_layer = nn.Sequential(
                nn.Linear(latent_dim, self.embedding_dim), nn.GELU(), nn.Dropout(dropout),
                nn.Linear(self.embedding_dim, self.embedding_dim))
        else:
            self.projection_layer = nn.Identity()
        for m in self.modules():                                               # model.py:129-136
            if isinstance(m, nn.Linear):
                nn.init.kaiming_normal_(m.weight, nonlinearity="relu")
                nn.init.constant_(m.bias, 0.0)

    # -- deterministic-noise forward: the same arithmetic as model.py:138-221 with the three
    #    random tensors (hidden dropout masks, eps, projection dropout mask) passed in.
    def encode_with(self, x, masks=None):
        h, li = x, 0
        for i in range(0, len(self.encoder), 4):
            h = self.encoder[i](h)
            h = self.encoder[i + 1](h)
            h = self.encoder[i + 2](h)
            if masks is not None:
                h = h * masks[li] * (1.0 / (1.0 - self.dropout))
            li += 1
        return self.fc_mu(h), self.fc_logvar(h)

    def project_with(self, z, mask=None):
        if isinstance(self.projection_layer, nn.Identity):
            return z
        t = self.projection_layer[1](self.projection_layer[0](z))
        if mask is not None:
            t = t * mask * (1.0 / (1.0 - self.dropout))
        return self.projection_layer[3](t)

    def forward_with(self, x, noise=None):
        """noise = None (eval: z = mu, no dropout) or dict(masks=[...], eps=..., pmask=...)."""
        mu, logvar = self.encode_with(x, None if noise is None else noise["masks"])
        z = mu if noise is None else mu + noise["eps"] * torch.exp(0.5 * logvar)   # model.py:168-179
        u = self.project_with(z, None if noise is None else noise.get("pmask"))
        return u @ self.item_embeddings.t(), mu, logvar                            # model.py:198

    def forward_ref(self, x):
        """Stock forward drawing from torch's global generator, as the reference does
        (src/ml/model.py:202-221)."""
        h = self.encoder(x)
        mu, logvar = self.fc_mu(h), self.fc_logvar(h)
        z = mu + torch.randn_like(mu) * torch.exp(0.5 * logvar) if self.training else mu
        return self.projection_layer(z) @ self.item_embeddings.t(), mu, logvar

    forward = forward_ref


def draw_noise(model: OracleVAE, batch: int):
    """Replays the reference's per-step draw order on torch's CPU generator: one Bernoulli
    keep-mask per hidden layer (model.py:117), eps (model.py:173), the projection dropout mask
    (model.py:93).  With the same torch.manual_seed this reproduces, bit for bit, the tensors
    the reference module consumes (checked in tests/test_oracle_golden.py)."""
    p = model.dropout
    masks = []
    for h in model.hidden_dims:
        if p > 0:
            masks.append(torch.native_dropout(torch.ones(batch, h), p, True)[1].float())
        else:
            masks.append(torch.ones(batch, h))
    eps = torch.randn(batch, model.latent_dim)
    pmask = None
    if not isinstance(model.projection_layer, nn.Identity):
        if p > 0:
            pmask = torch.native_dropout(torch.ones(batch, model.embedding_dim), p, True)[1].float()
        else:
            pmask = torch.ones(batch, model.embedding_dim)
    return dict(masks=masks, eps=eps, pmask=pmask)


def loss_terms(scores, x, mu, logvar, beta):
    """src/ml/model.py:281-290: multinomial NLL + beta * KL, both averaged over the batch."""
    recon = -torch.mean(torch.sum(x * F.log_softmax(scores, dim=-1), dim=-1))
    kl = -0.5 * torch.sum(1 + logvar - mu.pow(2) - logvar.exp()) / x.size(0)
    return recon + beta * kl, recon, kl


def annealed_beta(step, anneal_steps, beta_max, beta_min=0.0):
    """src/ml/model.py:312-319."""
    if step >= anneal_steps:
        return beta_max
    return beta_min + (step / anneal_steps) * (beta_max - beta_min)


def make_adam(model, lr=1e-3, weight_decay=0.0):
    """src/ml/train.py:63."""
    return torch.optim.Adam(model.parameters(), lr=lr, weight_decay=weight_decay)


def train_step(model, opt, x, noise, beta, max_norm=5.0):
    """One optimisation step, src/ml/train.py:88-96 (zero_grad, forward, loss, backward,
    clip_grad_norm_(5.0), Adam step)."""
    model.train()
    opt.zero_grad()
    scores, mu, logvar = model.forward_with(x, noise)
    loss, recon, kl = loss_terms(scores, x, mu, logvar, beta)
    loss.backward()
    gnorm = torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm=max_norm)
    opt.step()
    return loss.item(), recon.item(), kl.item(), float(gnorm)


def validate_batch(model, x, beta):
    """src/ml/train.py:107-117 for one batch."""
    model.eval()
    with torch.no_grad():
        scores, mu, logvar = model.forward_with(x, None)
        loss, recon, kl = loss_terms(scores, x, mu, logvar, beta)
    return float(loss), float(recon), float(kl)


# --------------------------------------------------------------------------------------
# metrics (src/ml/evaluate.py:32-54)
# --------------------------------------------------------------------------------------
def recall_at_k(rec, rel, k):
    if len(rel) == 0:
        return 0.0
    return len(np.intersect1d(rec[:k], rel)) / len(rel)


def ndcg_at_k(rec, rel, k):
    if len(rel) == 0:
        return 0.0
    dcg = sum(1.0 / np.log2(i + 2) for i, it in enumerate(rec[:k]) if it in rel)
    idcg = sum(1.0 / np.log2(i + 2) for i in range(min(len(rel), k)))
    return dcg / idcg if idcg > 0 else 0.0


def hit_ratio_at_k(rec, rel, k):
    if len(rel) == 0:
        return 0.0
    return 1.0 if len(np.intersect1d(rec[:k], rel)) > 0 else 0.0


def densify(csr, rows):
    return torch.from_numpy(np.asarray(csr[rows].toarray(), dtype=np.float32))


def user_scores(model, csr, u):
    """src/ml/evaluate.py:125-135."""
    with torch.no_grad():
        x = torch.from_numpy(csr[u].toarray().flatten().astype(np.float32)).unsqueeze(0)
        mu, _ = model.encode_with(x, None)
        return (model.project_with(mu) @ model.item_embeddings.t()).squeeze(0).numpy()


def topk_desc_index_desc(scores, k):
    """Top-k under the total order (score desc, index desc): the order a *stable* ascending
    argsort followed by [::-1] yields (src/ml/evaluate.py:146).  The reference's default
    argsort is unstable, so on exact ties its order is undefined; away from ties both agree."""
    order = np.argsort(scores, kind="stable")[::-1]
    return order[:k]


def recommend(model, csr, u, top_k=100, exclude_seen=True):
    """src/ml/evaluate.py:137-147."""
    s = user_scores(model, csr, u).copy()
    if exclude_seen:
        s[csr[u].nonzero()[1]] = -np.inf
    idx = topk_desc_index_desc(s, top_k)
    return idx, s[idx]


def full_ranking_eval(model, csr, test_items, k_values=(5, 10, 20), users=None):
    """src/ml/evaluate.py:217-265 with one test item per user; returns ({k:{...}}, topk [U,maxk])."""
    model.eval()
    users = range(csr.shape[0]) if users is None else users
    kmax = max(k_values)
    acc = {k: {"recall": [], "ndcg": [], "hit_ratio": []} for k in k_values}
    tops = []
    for u in users:
        rec, _ = recommend(model, csr, u, top_k=kmax)
        tops.append(rec)
        rel = np.array([test_items[u]])
        for k in k_values:
            acc[k]["recall"].append(recall_at_k(rec, rel, k))
            acc[k]["ndcg"].append(ndcg_at_k(rec, rel, k))
            acc[k]["hit_ratio"].append(hit_ratio_at_k(rec, rel, k))
    out = {k: {m: float(np.mean(v)) if v else 0.0 for m, v in acc[k].items()} for k in k_values}
    return out, np.stack(tops) if tops else np.zeros((0, kmax), np.int64)


def negative_sampling_rank(scores, test_item, negatives):
    """src/ml/evaluate.py:172-176: rank [test]+negatives by score, descending."""
    cand = np.concatenate([[test_item], negatives])
    return cand[np.argsort(scores[cand], kind="stable")[::-1]]


# --------------------------------------------------------------------------------------
# reference-shaped CPU loops, used only as the timed CPU baseline (bench.py)
# --------------------------------------------------------------------------------------
class _RowDataset(torch.utils.data.Dataset):
    """src/ml/train.py:35-47: one dense fp32 row per __getitem__."""

    def __init__(self, csr, rows):
        self.csr, self.rows = csr, list(rows)

    def __len__(self):
        return len(self.rows)

    def __getitem__(self, i):
        return torch.FloatTensor(self.csr[self.rows[i]].toarray().flatten())


def cpu_train_users(model, csr, rows, batch_size, lr=1e-3, beta=None, shuffle=True):
    """The reference's train_epoch (src/ml/train.py:81-103) over `rows`: DataLoader densify,
    stock dropout/randn from the global generator, Adam.  Returns (users, seconds, mean loss)."""
    opt = make_adam(model, lr)
    beta = model.beta if beta is None else beta
    loader = torch.utils.data.DataLoader(_RowDataset(csr, rows), batch_size=batch_size, shuffle=shuffle,
                                         num_workers=0)
    model.train()
    tot, t0 = 0.0, time.perf_counter()
    for x in loader:
        opt.zero_grad()
        s, mu, lv = model.forward_ref(x)
        loss, _, _ = loss_terms(s, x, mu, lv, beta)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm=5.0)
        opt.step()
        tot += loss.item()
    dt = time.perf_counter() - t0
    return len(rows), dt, tot / max(1, len(loader))


def cpu_eval_users(model, csr, test_items, users, k_values=(5, 10, 20)):
    """The reference's evaluate_dataset loop (src/ml/evaluate.py:243-265) over `users`."""
    t0 = time.perf_counter()
    res, _ = full_ranking_eval(model, csr, test_items, k_values, users)
    return len(list(users)), time.perf_counter() - t0, res


# --------------------------------------------------------------------------------------
# Mult-VAE baseline (src/ml/baseline.py:126-206)
# --------------------------------------------------------------------------------------
class OracleMultVAE(nn.Module):
    """Same module tree (state_dict keys, default nn.Linear initialisation, RNG draw order) as the reference's
    MultVAE (src/ml/baseline.py:129-149), with a forward that takes its noise as tensors."""

    def __init__(self, n_items, hidden_dim=600, latent_dim=200, dropout=0.5):
        super().__init__()
        self.p = dropout
        self.encoder = nn.Sequential(nn.Linear(n_items, hidden_dim), nn.Tanh(), nn.Dropout(dropout),
                                     nn.Linear(hidden_dim, hidden_dim), nn.Tanh(), nn.Dropout(dropout))
        self.mu_layer = nn.Linear(hidden_dim, latent_dim)
        self.logvar_layer = nn.Linear(hidden_dim, latent_dim)
        self.decoder = nn.Sequential(nn.Linear(latent_dim, hidden_dim), nn.Tanh(), nn.Linear(hidden_dim, n_items))
        self.dropout = nn.Dropout(dropout)

    def forward_with(self, x, noise=None):
        """noise = None (eval) or dict(in_mask [B,N] {0,1}, masks=[[B,h],[B,h]], eps [B,L]) -- baseline.py:150-160."""
        s = 1.0 / (1.0 - self.p) if noise is not None else 1.0
        xin = F.normalize(x, p=2, dim=1)
        if noise is not None:
            xin = xin * noise["in_mask"] * s
        h = torch.tanh(self.encoder[0](xin))
        if noise is not None:
            h = h * noise["masks"][0] * s
        h = torch.tanh(self.encoder[3](h))
        if noise is not None:
            h = h * noise["masks"][1] * s
        mu, logvar = self.mu_layer(h), self.logvar_layer(h)
        z = mu if noise is None else mu + torch.exp(0.5 * logvar) * noise["eps"]
        return self.decoder(z), mu, logvar


def draw_multvae_noise(model: OracleMultVAE, batch: int, n_items: int, hidden_dim: int, latent_dim: int):
    """The reference's per-step draw order on torch's CPU generator: input dropout over the dense [B, N] row
    (baseline.py:151), the two encoder dropouts (:139,:142), eps (:155)."""
    keep = lambda *shape: torch.native_dropout(torch.ones(*shape), model.p, True)[1].float()
    return dict(in_mask=keep(batch, n_items), masks=[keep(batch, hidden_dim), keep(batch, hidden_dim)],
                eps=torch.randn(batch, latent_dim))


def multvae_train_step(model, opt, x, noise, beta):
    """baseline.py:196-204: zero_grad, forward, multinomial NLL + beta * KL, backward, Adam step (no gradient clipping)."""
    model.train()
    opt.zero_grad()
    recon, mu, logvar = model.forward_with(x, noise)
    recon_loss = -torch.mean(torch.sum(F.log_softmax(recon, dim=1) * x, dim=1))
    kl_loss = -0.5 * torch.mean(torch.sum(1 + logvar - mu.pow(2) - logvar.exp(), dim=1))
    loss = recon_loss + beta * kl_loss
    loss.backward()
    opt.step()
    return loss.item(), recon_loss.item(), kl_loss.item()
