"""Drop-in HybridVAE / AnnealedVAE / vae_loss_function / create_hybrid_vae.

Mirrors the public surface of the reference's src/ml/model.py (constructor arguments, attributes, method
names, state_dict keys and shapes, SURVEY.md §8 a1-a8, b) while every FLOP runs in the hand-written sm_100a
kernels behind include/hvae_b200.h.  Parameters live in one flat fp32 arena (engine.Layout); the
reference-shaped tensors exist only at the state_dict()/load_state_dict() boundary.
"""
from __future__ import annotations

import logging
import os
from collections import OrderedDict
from typing import Optional, Tuple

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from .engine import Batch, DeviceCSR, Engine, Layout, r4
from ._cabi import p

logger = logging.getLogger(__name__)


def _default_precision():
    return os.environ.get("HVAE_B200_PRECISION", "bf16")


class HybridVAE(nn.Module):
    """Same constructor and methods as the reference class (src/ml/model.py:27-256).

    Extra keyword `precision`: "bf16" (tensor-core scoring, default) or "fp32" (exact mode: true fp32
    accumulation everywhere, bit-comparable top-K).  The item embeddings are a frozen buffer; a trainable
    E (`freeze_embeddings=False`) is outside the accelerated path and is rejected.
    """

    def __init__(self, n_items: int, item_embeddings: np.ndarray, latent_dim: int = 200,
                 hidden_dims: Optional[list] = None, dropout: float = 0.5, beta: float = 0.2,
                 freeze_embeddings: bool = True, precision: Optional[str] = None):
        super().__init__()
        if not freeze_embeddings:
            raise NotImplementedError("hvae_b200 keeps the item embeddings frozen (no gradient for E); "
                                      "freeze_embeddings=False is outside the accelerated hot path")
        self.n_items, self.latent_dim, self.dropout, self.beta = n_items, latent_dim, dropout, beta
        if hidden_dims is None:
            hidden_dims = [600, 200]                       # src/ml/model.py:65-66
        self.hidden_dims = hidden_dims
        item_embeddings = np.asarray(item_embeddings)
        if item_embeddings.shape[0] != n_items:
            raise ValueError(f"item_embeddings has {item_embeddings.shape[0]} rows, expected n_items={n_items}")
        self.embedding_dim = item_embeddings.shape[1]
        self.precision = precision or _default_precision()
        if self.precision not in ("bf16", "fp32"):
            raise ValueError("precision must be 'bf16' or 'fp32'")
        self.register_buffer("item_embeddings", torch.tensor(item_embeddings, dtype=torch.float32))
        self.layout = Layout(n_items, self.embedding_dim, latent_dim, hidden_dims)
        arena = torch.zeros(self.layout.n_params, dtype=torch.float32)
        self._reference_init(arena)
        self.arena = nn.Parameter(arena)
        self._eng = None
        logger.info("Initializing HybridVAE (hvae_b200): items=%d latent=%d emb=%d hidden=%s beta=%s precision=%s",
                    n_items, latent_dim, self.embedding_dim, hidden_dims, beta, self.precision)

    # -- initialisation --------------------------------------------------------------------------------------
    def _reference_init(self, arena):
        """Kaiming-normal weights / zero biases / unit LayerNorm, drawing from torch's global CPU generator in
        the order the reference's constructor does (default nn.Linear init first, then _init_weights,
        src/ml/model.py:103-136), so that the same torch.manual_seed gives the same initial weights."""
        lay, h = self.layout, self.hidden_dims
        dims = [(self.n_items, h[0], "encoder.0")]
        dims += [(h[i - 1], h[i], f"encoder.{4 * i}") for i in range(1, len(h))]
        dims += [(h[-1], self.latent_dim, "fc_mu"), (h[-1], self.latent_dim, "fc_logvar")]
        if not lay.identity_proj:
            dims += [(self.latent_dim, self.embedding_dim, "projection_layer.0"),
                     (self.embedding_dim, self.embedding_dim, "projection_layer.3")]
        mods = [(nn.Linear(i, o), name) for i, o, name in dims]     # consumes the default-init draws
        with torch.no_grad():
            for lin, name in mods:
                nn.init.kaiming_normal_(lin.weight, nonlinearity="relu")
                lay.view(arena, name + ".weight").copy_(lin.weight)
                lay.view(arena, name + ".bias").zero_()
            for i in range(len(h)):
                lay.view(arena, f"encoder.{4 * i + 1}.weight").fill_(1.0)
                lay.view(arena, f"encoder.{4 * i + 1}.bias").zero_()

    # -- engine ------------------------------------------------------------------------------------------------
    @property
    def engine(self) -> Engine:
        a = self.arena.data
        if self._eng is None or self._eng.arena.data_ptr() != a.data_ptr() or self._eng.precision != self.precision:
            self._eng = Engine(self.layout, a, self.item_embeddings, self.dropout, self.precision)
        return self._eng

    def num_parameters(self) -> int:
        """Trainable parameter count in the reference's terms (padding excluded)."""
        return sum(int(np.prod(self.layout.view(self.arena, k).shape)) for k in self.layout.reference_keys())

    # -- state dict in the reference layout --------------------------------------------------
    def state_dict(self, *args, destination=None, prefix="", keep_vars=False):
        sd = OrderedDict() if destination is None else destination
        sd[prefix + "item_embeddings"] = self.item_embeddings.detach().clone()
        for k, v in self.layout.export(self.arena.data).items():
            sd[prefix + k] = v
        return sd

    def load_state_dict(self, state_dict, strict: bool = True, assign: bool = False):
        expected = ["item_embeddings"] + self.layout.reference_keys()
        missing = [k for k in expected if k not in state_dict]
        unexpected = [k for k in state_dict if k not in expected]
        if strict and (missing or unexpected):
            raise RuntimeError(f"Error(s) in loading state_dict for HybridVAE: missing keys {missing}, "
                               f"unexpected keys {unexpected}")
        if missing:
            raise RuntimeError(f"missing keys {missing}")
        with torch.no_grad():
            if tuple(state_dict["item_embeddings"].shape) != tuple(self.item_embeddings.shape):
                raise RuntimeError("size mismatch for item_embeddings")
            self.item_embeddings.copy_(state_dict["item_embeddings"])
        self.layout.load(self.arena.data, state_dict)
        if self._eng is not None:
            self._eng.invalidate_embeddings()
        return torch.nn.modules.module._IncompatibleKeys(missing, unexpected)

    # -- inputs ---------------------------------------------------------------------------------------------
    def _as_batch(self, x) -> Batch:
        if isinstance(x, Batch):
            return x
        if isinstance(x, DeviceCSR):
            return x.full_batch()
        if not isinstance(x, torch.Tensor):
            raise TypeError(f"unsupported input type {type(x)}")
        if not x.is_cuda:
            raise RuntimeError("hvae_b200 runs on CUDA devices only (there is no CPU fallback)")
        if x.layout == torch.sparse_csr:
            csr = DeviceCSR(x.crow_indices().to(torch.int64), x.col_indices().to(torch.int32),
                            x.values().to(torch.float32), x.shape[0], x.shape[1],
                            np.diff(x.crow_indices().cpu().numpy()))
            return csr.full_batch()
        if x.dim() == 1:
            x = x.unsqueeze(0)
        if x.shape[1] != self.n_items:
            raise RuntimeError(f"input has {x.shape[1]} columns, model has n_items={self.n_items}")
        return DeviceCSR.from_dense(x.float()).full_batch()

    def _draw_noise(self, B, encoder_only=False):
        """Training-mode noise from torch's generator on the model's device, in the reference's draw order
        (hidden dropout masks, eps, projection dropout mask; src/ml/model.py:117,173,93)."""
        dev, pdrop = self.arena.device, self.dropout
        masks = []
        for h in self.hidden_dims:
            masks.append(torch.native_dropout(torch.ones(B, h, device=dev), pdrop, True)[1].to(torch.uint8)
                         if pdrop > 0 else None)
        if encoder_only:
            return dict(masks=masks, eps=None, pmask=None)
        eps = torch.randn(B, self.latent_dim, device=dev)
        pmask = None
        if not self.layout.identity_proj and pdrop > 0:
            pmask = torch.native_dropout(torch.ones(B, self.embedding_dim, device=dev), pdrop, True)[1].to(torch.uint8)
        return dict(masks=masks, eps=eps, pmask=pmask)

    # -- reference methods ------------------------------------------------------------------------------------
    def encode(self, x) -> Tuple[torch.Tensor, torch.Tensor]:
        """(mu, logvar) -- src/ml/model.py:138-155.  Dropout is active in train() mode, as in the reference."""
        b = self._as_batch(x)
        eng, L = self.engine, self.latent_dim
        masks = self._draw_noise(b.B, encoder_only=True)["masks"] if self.training else None
        with torch.no_grad():
            ml = eng.encode(b, masks)
            return ml[:, :L].clone(), ml[:, L:2 * L].clone()

    def reparameterize(self, mu: torch.Tensor, logvar: torch.Tensor) -> torch.Tensor:
        """src/ml/model.py:157-179."""
        if self.training:
            return mu + torch.randn_like(mu) * torch.exp(0.5 * logvar)
        return mu

    def decode(self, z: torch.Tensor) -> torch.Tensor:
        """Item scores [B, N] = projection(z) E^T -- src/ml/model.py:181-200 (materialised, fp32)."""
        if not z.is_cuda:
            raise RuntimeError("hvae_b200 runs on CUDA devices only (there is no CPU fallback)")
        eng, lay = self.engine, self.layout
        squeeze = z.dim() == 1
        z2 = (z.unsqueeze(0) if squeeze else z).float().contiguous()
        B, L = z2.shape
        with torch.no_grad():
            ml = eng.ws.get("ml", (B, r4(2 * L)))
            ml.zero_()
            ml[:, :L] = z2
            pmask = None
            if self.training and not lay.identity_proj and self.dropout > 0:
                pmask = torch.native_dropout(torch.ones(B, self.embedding_dim, device=z.device), self.dropout, True)[1].to(torch.uint8)
            u = eng.latent_and_project(B, ml, None, pmask, want_kl=False)
            s = eng.scores_dense(u, B)
        return s[0] if squeeze else s

    def forward(self, x) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """(scores [B,N], mu, logvar) -- src/ml/model.py:202-221.  Differentiable w.r.t. the parameters."""
        b = self._as_batch(x)
        noise = self._draw_noise(b.B) if self.training else None
        if torch.is_grad_enabled() and self.arena.requires_grad:
            return _ForwardFn.apply(self.arena, self, b, noise)
        with torch.no_grad():
            return _forward_impl(self, b, noise)

    def get_user_embedding(self, x) -> torch.Tensor:
        """mu -- src/ml/model.py:223-234."""
        return self.encode(x)[0]

    def recommend(self, user_embedding: torch.Tensor, top_k: int = 10) -> Tuple[torch.Tensor, torch.Tensor]:
        """(top indices, top scores) -- src/ml/model.py:236-256."""
        with torch.no_grad():
            scores = self.decode(user_embedding)
            s2 = (scores.unsqueeze(0) if scores.dim() == 1 else scores).contiguous().clone()
            B = s2.shape[0]
            eng = self.engine
            val = torch.empty(B, top_k, dtype=torch.float32, device=s2.device)
            idx = torch.empty(B, top_k, dtype=torch.int32, device=s2.device)
            zero_ptr = torch.zeros(B + 1, dtype=torch.int64, device=s2.device)
            nc = int(eng.lib.mask_topk_chunks(B, s2.shape[1]))
            cv = torch.empty(B, nc * top_k, dtype=torch.float32, device=s2.device)
            ci = torch.empty(B, nc * top_k, dtype=torch.int32, device=s2.device)
            eng.lib.mask_topk(p(s2), s2.shape[1], B, s2.shape[1], 0, p(zero_ptr), None, None, 0, top_k, p(cv), p(ci), p(val), p(idx),
                              eng.stream)
        return idx.long(), val


def _forward_impl(model: HybridVAE, b: Batch, noise):
    eng, L = model.engine, model.latent_dim
    ml = eng.encode(b, None if noise is None else noise["masks"])
    u = eng.latent_and_project(b.B, ml, None if noise is None else noise["eps"], None if noise is None else noise.get("pmask"),
                               want_kl=False)
    scores = eng.scores_dense(u, b.B)
    return scores, ml[:, :L].clone(), ml[:, L:2 * L].clone()


_SAVED = ["ml", "z", "q", "t"]


class _ForwardFn(torch.autograd.Function):
    """Autograd bridge for the reference-style loop (model(x) -> vae_loss_function -> loss.backward()):
    the backward runs the same hand-written kernels as the fused trainer, with a dense d(W1)."""

    @staticmethod
    def forward(ctx, arena, model, b, noise):
        eng = model.engine
        out = _forward_impl(model, b, noise)
        names = [n for n in _SAVED if n in eng.ws.buf] + [f"{k}{i}" for i in range(len(model.hidden_dims))
                                                           for k in ("pre", "act", "mean", "rstd")]
        ctx.saved = {n: eng.ws.buf[n].clone() for n in names}
        ctx.model, ctx.b, ctx.noise = model, b, noise
        return out

    @staticmethod
    def backward(ctx, dscores, dmu, dlogvar):
        model, b, noise = ctx.model, ctx.b, ctx.noise
        eng, lay = model.engine, model.layout
        eng.ensure_optimizer()
        for n, t in ctx.saved.items():
            eng.ws.buf[n][:t.numel()].copy_(t)
        B, L, d = b.B, lay.L, lay.d
        ldd = r4(d)
        dU = eng.ws.get("dU", (B, ldd))
        dU.zero_()
        if dscores is not None:
            ds = dscores.contiguous().float()
            eng.gemm(B, d, lay.N, p(ds), lay.N, 1, p(eng.E), d, 1, p(dU), ldd)
        ext = torch.zeros(B, 2 * L, device=dU.device)
        if dmu is not None:
            ext[:, :L] = dmu
        if dlogvar is not None:
            ext[:, L:] = dlogvar
        ld1 = r4(lay.hidden[0])
        dense_w1 = torch.zeros(lay.N, ld1, device=dU.device)
        ml = eng.ws.get("ml", (B, r4(2 * L)))
        eng.backward(b, noise, ml, None, None, dense_w1=dense_w1, ext_dml=ext, du_override=dU)
        grad = torch.cat([dense_w1.view(-1), eng.gd])
        return grad, None, None, None


def vae_loss_function(recon_x: torch.Tensor, x: torch.Tensor, mu: torch.Tensor, logvar: torch.Tensor,
                      beta: float = 0.2) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """(total, recon, kl) on materialised scores -- src/ml/model.py:259-292.  Compatibility entry point for
    callers that already hold recon_x; the trainer uses the fused kernels and never materialises recon_x."""
    if x.layout != torch.strided:
        x = x.to_dense()
    recon_loss = -torch.mean(torch.sum(x * F.log_softmax(recon_x, dim=-1), dim=-1))
    kl_loss = -0.5 * torch.sum(1 + logvar - mu.pow(2) - logvar.exp()) / x.size(0)
    return recon_loss + beta * kl_loss, recon_loss, kl_loss


class AnnealedVAE(HybridVAE):
    """Linear KL-weight warm-up -- src/ml/model.py:295-334."""

    def __init__(self, *args, **kwargs):
        self.beta_min = kwargs.pop("beta_min", 0.0)
        self.beta_max = kwargs.pop("beta_max", kwargs.get("beta", 0.2))
        self.anneal_steps = kwargs.pop("anneal_steps", 10000)
        super().__init__(*args, **kwargs)
        self.current_step = 0

    def get_current_beta(self) -> float:
        if self.current_step >= self.anneal_steps:
            return self.beta_max
        return self.beta_min + (self.current_step / self.anneal_steps) * (self.beta_max - self.beta_min)

    def step_annealing(self):
        self.current_step += 1

    def compute_loss(self, recon_x, x, mu, logvar):
        return vae_loss_function(recon_x, x, mu, logvar, self.get_current_beta())


def create_hybrid_vae(n_items: int, item_embeddings: np.ndarray, latent_dim: int = 200, hidden_dims: Optional[list] = None,
                      dropout: float = 0.5, beta: float = 0.2, use_annealing: bool = False, freeze_embeddings: bool = True,
                      **annealing_kwargs) -> HybridVAE:
    """Factory -- src/ml/model.py:337-385 (annealing kwargs are dropped when use_annealing is False, as there)."""
    precision = annealing_kwargs.pop("precision", None)
    common = dict(n_items=n_items, item_embeddings=item_embeddings, latent_dim=latent_dim, hidden_dims=hidden_dims,
                  dropout=dropout, beta=beta, freeze_embeddings=freeze_embeddings, precision=precision)
    if use_annealing:
        return AnnealedVAE(**common, **annealing_kwargs)
    return HybridVAE(**common)
