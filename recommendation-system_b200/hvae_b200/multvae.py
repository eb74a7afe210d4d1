"""Drop-in Mult-VAE baseline: MultVAE + MultVAERecommender (reference src/ml/baseline.py:126-231) on the HybridVAE kernels.

The reference model is  x -> F.normalize -> Dropout -> Linear(N,h) -> Tanh -> Dropout -> Linear(h,h) -> Tanh -> Dropout ->
(mu, logvar) -> z -> Linear(L,h) -> Tanh -> Linear(h,N)  with a multinomial NLL + beta*KL loss and plain Adam (no clipping).
Here:
  * the first Linear on the normalised, dropped-out SPARSE row is the CSR gather-sum kernel with per-entry values
    (hvae_mv_input_values + hvae_gather_ln_fwd in its plain-linear mode); its weight gradient is the item-major reduction;
  * the hidden stack runs on the MLP GEMMs + hvae_tanh_drop_fwd/bwd, mu/logvar/z/KL on hvae_reparam_kl / hvae_latent_bwd;
  * the output layer is TRAINABLE here (unlike HybridVAE's frozen E), so the scores go through the materialised fp32 path
    (GEMM -> row log-sum-exp -> softmax scale -> two GEMMs for dT and dW); its bias is one more weight column against a
    constant-one activation column, so every kernel sees a bias-free [N, h+1] matrix;
  * one fused Adam over the flat parameter arena.
state_dict() uses the reference's keys and shapes (encoder.0.weight [h,N], ..., decoder.2.weight [N,h], decoder.2.bias [N]).
CUDA only, no fallback.
"""
from __future__ import annotations

import logging
from collections import OrderedDict
from types import SimpleNamespace

import numpy as np
import torch
import torch.nn as nn

from . import _cabi
from ._cabi import STATE_OFF, STATE_WORDS, p
from .engine import Batch, DeviceCSR, Engine, Workspace, r4

logger = logging.getLogger(__name__)


class _Slots:
    """Flat fp32 arena: [W1^T [N, ld(h)] | b1 | W2 | b2 | W_mu;W_logvar | b_ml | Wd0 | bd0 | W_out|b_out [N, ld(h+1)]]."""

    def __init__(self, N, h, L):
        self.N, self.h, self.L = N, h, L
        self.ldh, self.ldL, self.ld2L, self.lda = r4(h), r4(L), r4(2 * L), r4(h + 1)
        off = 0

        def take(n):
            nonlocal off
            o = off
            off += n
            return o
        self.w1 = take(N * self.ldh)
        self.n_w1 = off
        self.b1 = take(self.ldh)
        self.w2 = take(h * self.ldh)
        self.b2 = take(self.ldh)
        self.wml = take(2 * L * self.ldh)
        self.bml = take(self.ld2L)
        self.wd0 = take(h * self.ldL)
        self.bd0 = take(self.ldh)
        self.waug = take(N * self.lda)
        self.n_params = off
        self.n_dense = off - self.n_w1


class MultVAE(nn.Module):
    """src/ml/baseline.py:126-160: same constructor, methods and state_dict layout; forward() returns (logits, mu, logvar)."""

    def __init__(self, n_items: int, hidden_dim: int = 600, latent_dim: int = 200, dropout: float = 0.5):
        super().__init__()
        self.n_items, self.hidden_dim, self.latent_dim, self.p_drop = n_items, hidden_dim, latent_dim, float(dropout)
        self.sl = _Slots(n_items, hidden_dim, latent_dim)
        arena = torch.zeros(self.sl.n_params, dtype=torch.float32)
        # default nn.Linear initialisation, drawn in the reference constructor's order (baseline.py:134-149)
        mods = [nn.Linear(n_items, hidden_dim), nn.Linear(hidden_dim, hidden_dim), nn.Linear(hidden_dim, latent_dim),
                nn.Linear(hidden_dim, latent_dim), nn.Linear(latent_dim, hidden_dim), nn.Linear(hidden_dim, n_items)]
        sd = {}
        for name, m in zip(self.KEYS, mods):
            sd[name + ".weight"], sd[name + ".bias"] = m.weight.detach(), m.bias.detach()
        self.arena = nn.Parameter(arena)
        self._load(sd)
        self._rt = None

    KEYS = ["encoder.0", "encoder.3", "mu_layer", "logvar_layer", "decoder.0", "decoder.2"]

    # -- reference-shaped views of the arena --------------------------------------------------------------------
    def _views(self, a):
        s, h, L, N = self.sl, self.hidden_dim, self.latent_dim, self.n_items
        v = OrderedDict()
        v["encoder.0.weight"] = a[s.w1:s.w1 + N * s.ldh].view(N, s.ldh)[:, :h].t()
        v["encoder.0.bias"] = a[s.b1:s.b1 + h]
        v["encoder.3.weight"] = a[s.w2:s.w2 + h * s.ldh].view(h, s.ldh)[:, :h]
        v["encoder.3.bias"] = a[s.b2:s.b2 + h]
        wml = a[s.wml:s.wml + 2 * L * s.ldh].view(2 * L, s.ldh)
        v["mu_layer.weight"], v["mu_layer.bias"] = wml[:L, :h], a[s.bml:s.bml + L]
        v["logvar_layer.weight"], v["logvar_layer.bias"] = wml[L:, :h], a[s.bml + L:s.bml + 2 * L]
        v["decoder.0.weight"] = a[s.wd0:s.wd0 + h * s.ldL].view(h, s.ldL)[:, :L]
        v["decoder.0.bias"] = a[s.bd0:s.bd0 + h]
        waug = a[s.waug:s.waug + N * s.lda].view(N, s.lda)
        v["decoder.2.weight"], v["decoder.2.bias"] = waug[:, :h], waug[:, h]
        return v

    def _load(self, sd):
        with torch.no_grad():
            for k, dst in self._views(self.arena.data).items():
                if tuple(sd[k].shape) != tuple(dst.shape):
                    raise RuntimeError(f"size mismatch for {k}: {tuple(sd[k].shape)} vs {tuple(dst.shape)}")
                dst.copy_(sd[k])

    def state_dict(self, *args, destination=None, prefix="", keep_vars=False):
        sd = OrderedDict() if destination is None else destination
        for k, v in self._views(self.arena.data).items():
            sd[prefix + k] = v.detach().clone().contiguous()
        return sd

    def load_state_dict(self, state_dict, strict: bool = True, assign: bool = False):
        expected = list(self._views(self.arena.data))
        missing = [k for k in expected if k not in state_dict]
        unexpected = [k for k in state_dict if k not in expected]
        if missing or (strict and unexpected):
            raise RuntimeError(f"Error(s) in loading state_dict for MultVAE: missing keys {missing}, unexpected keys {unexpected}")
        self._load(state_dict)
        return torch.nn.modules.module._IncompatibleKeys(missing, unexpected)

    # -- runtime (device buffers, kernel sequencing) ------------------------------------------------------------------
    @property
    def rt(self):
        a = self.arena.data
        if not a.is_cuda:
            raise RuntimeError("hvae_b200.MultVAE runs on CUDA devices only (there is no CPU fallback); move it with .to('cuda')")
        if self._rt is None or self._rt.arena.data_ptr() != a.data_ptr():
            self._rt = _Runtime(self, a)
        return self._rt

    def _batch(self, x) -> Batch:
        if isinstance(x, Batch):
            return x
        if isinstance(x, DeviceCSR):
            return x.full_batch()
        if not x.is_cuda:
            raise RuntimeError("hvae_b200 runs on CUDA devices only (there is no CPU fallback)")
        if x.dim() == 1:
            x = x.unsqueeze(0)
        if x.layout == torch.sparse_csr:
            x = x.to_dense()
        if x.shape[1] != self.n_items:
            raise RuntimeError(f"input has {x.shape[1]} columns, model has n_items={self.n_items}")
        return DeviceCSR.from_dense(x.float()).full_batch()

    def encode(self, x):
        b = self._batch(x)
        with torch.no_grad():
            ml = self.rt.encode(b, self.rt.draw_noise(b) if self.training else None)
        L = self.latent_dim
        return ml[:, :L].clone(), ml[:, L:2 * L].clone()

    def reparameterize(self, mu, logvar):
        if self.training:
            std = torch.exp(0.5 * logvar)
            return mu + std * torch.randn_like(std)
        return mu

    def forward(self, x):
        """(logits [B, N], mu, logvar) -- baseline.py:157-160 (inference / scoring path; training goes through train_step)."""
        b = self._batch(x)
        with torch.no_grad():
            rt = self.rt
            noise = rt.draw_noise(b) if self.training else None
            ml = rt.encode(b, noise)
            t = rt.decode_hidden(b.B, ml, None if noise is None else noise["eps"])
            S = torch.empty(b.B, self.n_items, dtype=torch.float32, device=ml.device)
            rt.scores(t, b.B, S)
        L = self.latent_dim
        return S, ml[:, :L].clone(), ml[:, L:2 * L].clone()


class _Runtime:
    ADAM_B1, ADAM_B2, ADAM_EPS = 0.9, 0.999, 1e-8

    def __init__(self, model: MultVAE, arena):
        self.m, self.arena, self.sl = model, arena, model.sl
        self.dev = arena.device
        self.lib = _cabi.lib()
        self.ws = Workspace(self.dev)
        self.state = torch.zeros(STATE_WORDS, dtype=torch.float32, device=self.dev)
        self.loss_out = torch.zeros(3, dtype=torch.float32, device=self.dev)
        self.acc = torch.zeros(4, dtype=torch.float32, device=self.dev)
        self.mom = self.var = self.gd = self.slot_of_item = None
        self.keep_scale = 1.0 / (1.0 - model.p_drop) if model.p_drop < 1.0 else 0.0
        # the batch-transposition / layer-1 gradient helpers of the HybridVAE engine, on this model's shapes
        self._eng = Engine.__new__(Engine)
        self._eng.lay = SimpleNamespace(N=model.n_items, hidden=[model.hidden_dim])
        self._eng.ws, self._eng.lib, self._eng.dev, self._eng.prof = self.ws, self.lib, self.dev, None
        self._eng.slot_of_item = None

    @property
    def stream(self):
        return torch.cuda.current_stream(self.dev).cuda_stream

    def P(self, off):
        return self.arena.data_ptr() + 4 * off

    def G(self, off):
        return self.gd.data_ptr() + 4 * (off - self.sl.n_w1)

    def ensure_optimizer(self):
        if self.mom is None:
            self.mom, self.var = torch.zeros_like(self.arena), torch.zeros_like(self.arena)
            self.gd = torch.zeros(self.sl.n_dense, dtype=torch.float32, device=self.dev)
            self.slot_of_item = torch.full((self.sl.N,), -1, dtype=torch.int32, device=self.dev)
            self._eng.slot_of_item = self.slot_of_item

    def gemm(self, M, N, K, A, a_rs, a_cs, Bp, b_rs, b_cs, C, ldc, bias=None, alpha=1.0):
        self.lib.gemm_f32(M, N, K, A, a_rs, a_cs, Bp, b_rs, b_cs, C, ldc, bias, alpha, self.stream)

    # -- noise ------------------------------------------------------------------------------------------------------
    def batch_offsets(self, b: Batch):
        boff = self.ws.get("boff", (b.B + 1,), torch.int32)
        self.lib.batch_offsets(p(b.csr.indptr), p(b.rows), b.B, p(boff), self.stream)
        return boff

    def draw_noise(self, b: Batch):
        """Input-dropout keep flags per batch entry, the two hidden dropout masks, eps -- the reference's draw order
        (baseline.py:151,139,142,155), from torch's generator on the device."""
        pd, h, L = self.m.p_drop, self.m.hidden_dim, self.m.latent_dim
        nnz = int(b.nnz_cap)
        keep = lambda n: (torch.rand(n, device=self.dev) >= pd).to(torch.uint8)
        return dict(keep=keep(nnz) if pd > 0 else None, masks=[keep(b.B * h).view(b.B, h) if pd > 0 else None for _ in range(2)],
                    eps=torch.randn(b.B, L, device=self.dev))

    # -- forward ---------------------------------------------------------------------------------------------------
    def encode(self, b: Batch, noise):
        sl, ws, lib, st = self.sl, self.ws, self.lib, self.stream
        B, h, L = b.B, sl.h, sl.L
        csr = b.csr
        vals = ws.get("mv_vals", (int(csr.indices.shape[0]),))
        keep = None if noise is None else noise.get("keep")
        boff = self.batch_offsets(b) if keep is not None else None
        lib.mv_input_values(p(csr.indptr), p(csr.values), p(b.rows), B, p(keep), p(boff), self.keep_scale, p(vals), st)
        self._in_csr = DeviceCSR(csr.indptr, csr.indices, vals, csr.n_users, csr.n_items, csr.host_lengths)
        h1 = ws.get("h1pre", (B, sl.ldh))
        lib.gather_ln_fwd(p(csr.indptr), p(csr.indices), p(vals), p(b.rows), B, self.P(sl.w1), sl.ldh, h, self.P(sl.b1), None, None, None,
                          1.0, None, None, None, p(h1), st)
        m1 = None if noise is None else noise["masks"][0]
        m2 = None if noise is None else noise["masks"][1]
        a1, a2, h2 = ws.get("a1", (B, sl.ldh)), ws.get("a2", (B, sl.ldh)), ws.get("h2pre", (B, sl.ldh))
        lib.tanh_drop_fwd(p(h1), p(m1), self.keep_scale, B, h, sl.ldh, p(a1), sl.ldh, -1, st)
        self.gemm(B, h, h, p(a1), sl.ldh, 1, self.P(sl.w2), 1, sl.ldh, p(h2), sl.ldh, self.P(sl.b2))
        lib.tanh_drop_fwd(p(h2), p(m2), self.keep_scale, B, h, sl.ldh, p(a2), sl.ldh, -1, st)
        ml = ws.get("ml", (B, sl.ld2L))
        self.gemm(B, 2 * L, h, p(a2), sl.ldh, 1, self.P(sl.wml), 1, sl.ldh, p(ml), sl.ld2L, self.P(sl.bml))
        return ml

    def decode_hidden(self, B, ml, eps, want_kl=False):
        """z = mu (+ eps * std), KL rows, Linear(L,h) + Tanh, with the constant-one column appended -> t_aug [B, ld(h+1)]."""
        sl, ws, lib, st = self.sl, self.ws, self.lib, self.stream
        z, kl = ws.get("z", (B, sl.ldL)), ws.get("kl_row", (B,))
        lib.reparam_kl(p(ml), sl.ld2L, p(eps), B, sl.L, p(z), sl.ldL, p(kl) if want_kl else None, st)
        dp = ws.get("dpre", (B, sl.ldh))
        self.gemm(B, sl.h, sl.L, p(z), sl.ldL, 1, self.P(sl.wd0), 1, sl.ldL, p(dp), sl.ldh, self.P(sl.bd0))
        t = ws.get("t_aug", (B, sl.lda))
        lib.tanh_drop_fwd(p(dp), None, 1.0, B, sl.h, sl.ldh, p(t), sl.lda, sl.h, st)
        return t

    def scores(self, t, B, S):
        sl = self.sl
        self.gemm(B, sl.N, sl.h + 1, p(t), sl.lda, 1, self.P(sl.waug), 1, sl.lda, p(S), sl.N)

    # -- one optimisation step (baseline.py:192-206: zero_grad, forward, loss, backward, Adam.step -- no clipping) -----------
    def train_step(self, b: Batch, noise, lr, beta, accumulate=True):
        self.ensure_optimizer()
        sl, ws, lib, st, eng = self.sl, self.ws, self.lib, self.stream, self._eng
        B, h, L, N = b.B, sl.h, sl.L, sl.N
        if B * N > 256 * 1024 * 1024:
            raise RuntimeError("MultVAE keeps one [B, N] score tile: batch too large for this catalogue")
        csr = b.csr
        lib.step_begin(p(self.state), lr, self.ADAM_B1, self.ADAM_B2, 0.0, beta, 0, B, 1, 0, st)
        sp = lambda f: self.state.data_ptr() + 4 * STATE_OFF[f]
        ml = self.encode(b, noise)
        t = self.decode_hidden(B, ml, noise["eps"], want_kl=True)
        S = ws.get("S", (B, N))
        self.scores(t, B, S)
        lse, dot, xsum = ws.get("lse", (B,)), ws.get("dot", (B,)), ws.get("xsum", (B,))
        lib.sparse_dot_xsum(p(csr.indptr), p(csr.indices), p(csr.values), p(b.rows), B, p(t), sl.lda, self.P(sl.waug), sl.lda, h + 1, 0,
                            p(dot), p(xsum), st)
        lib.row_lse(p(S), N, B, N, p(lse), st)
        lib.loss_finalize(p(lse), p(dot), p(xsum), p(ws.get("kl_row", (B,))), B, sp("inv_bg"), sp("beta_kl"), p(self.loss_out),
                          p(self.acc) if accumulate else None, st)
        # ---- backward.  dS = (softmax * |x| - x) / B: the softmax part is dense, the "- x" part touches the batch's items only
        lib.row_softmax_scale(p(S), N, B, N, p(lse), p(xsum), sp("inv_bg"), st)
        dT = ws.get("dT", (B, sl.lda))
        self.gemm(B, h + 1, N, p(S), N, 1, self.P(sl.waug), sl.lda, 1, p(dT), sl.lda)                 # dT = dS_dense W_aug
        self.gemm(N, h + 1, B, p(S), 1, N, p(t), sl.lda, 1, self.G(sl.waug), sl.lda)                  # dW_aug = dS_dense^T t_aug
        dT2 = ws.get("dT2", (B, sl.lda))
        lib.du_finalize(p(csr.indptr), p(csr.indices), p(csr.values), p(b.rows), B, p(dT), sl.lda, 1, None, None, self.P(sl.waug), sl.lda,
                        h + 1, 0, sp("inv_bg"), p(dT2), sl.lda, st)
        # item-major views of the batch: raw values for the output layer, transformed values for the first layer
        tb_raw = eng.transpose_batch(b)
        tb_raw = dict(tb_raw, ent_val=tb_raw["ent_val"].clone())
        gs2, _ = self._rows_grad(tb_raw, t, sl.lda, "gs_out")
        lib.rows_axpy(p(tb_raw["uniq"]), p(tb_raw["n_unique"]), tb_raw["cap"], p(gs2), sl.lda, -1.0 / B, self.G(sl.waug), sl.lda, sl.lda, st)
        lib.batch_release(p(tb_raw["uniq"]), p(tb_raw["n_unique"]), b.nnz_cap, p(self.slot_of_item), None, None, None, st)
        # decoder hidden layer
        cs = ws.get("colsum_ws", (64 * max(h, 2 * L) + 64,))
        ddp = ws.get("ddpre", (B, sl.ldh))
        lib.tanh_drop_bwd(p(dT2), sl.lda, p(ws.get("dpre", (B, sl.ldh))), None, 1.0, B, h, sl.ldh, p(ddp), st)
        z = ws.get("z", (B, sl.ldL))
        self.gemm(h, L, B, p(ddp), 1, sl.ldh, p(z), sl.ldL, 1, self.G(sl.wd0), sl.ldL)
        lib.colsum(p(ddp), sl.ldh, B, h, self.G(sl.bd0), p(cs), st)
        dz = ws.get("dz", (B, sl.ldL))
        self.gemm(B, L, h, p(ddp), sl.ldh, 1, self.P(sl.wd0), sl.ldL, 1, p(dz), sl.ldL)
        dml = ws.get("dml", (B, sl.ld2L))
        lib.latent_bwd(p(dz), sl.ldL, p(ml), sl.ld2L, p(noise["eps"]), B, L, sp("kl_coef"), p(dml), st)
        a1, a2 = ws.get("a1", (B, sl.ldh)), ws.get("a2", (B, sl.ldh))
        self.gemm(2 * L, h, B, p(dml), 1, sl.ld2L, p(a2), sl.ldh, 1, self.G(sl.wml), sl.ldh)
        lib.colsum(p(dml), sl.ld2L, B, 2 * L, self.G(sl.bml), p(cs), st)
        da2 = ws.get("da2", (B, sl.ldh))
        self.gemm(B, h, 2 * L, p(dml), sl.ld2L, 1, self.P(sl.wml), sl.ldh, 1, p(da2), sl.ldh)
        dh2 = ws.get("dh2pre", (B, sl.ldh))
        lib.tanh_drop_bwd(p(da2), sl.ldh, p(ws.get("h2pre", (B, sl.ldh))), p(noise["masks"][1]), self.keep_scale, B, h, sl.ldh, p(dh2), st)
        self.gemm(h, h, B, p(dh2), 1, sl.ldh, p(a1), sl.ldh, 1, self.G(sl.w2), sl.ldh)
        lib.colsum(p(dh2), sl.ldh, B, h, self.G(sl.b2), p(cs), st)
        da1 = ws.get("da1", (B, sl.ldh))
        self.gemm(B, h, h, p(dh2), sl.ldh, 1, self.P(sl.w2), sl.ldh, 1, p(da1), sl.ldh)
        dh1 = ws.get("dh1pre", (B, sl.ldh))
        lib.tanh_drop_bwd(p(da1), sl.ldh, p(ws.get("h1pre", (B, sl.ldh))), p(noise["masks"][0]), self.keep_scale, B, h, sl.ldh, p(dh1), st)
        lib.colsum(p(dh1), sl.ldh, B, h, self.G(sl.b1), p(cs), st)
        # first layer: item-major reduction over the transformed input values
        tb_in = eng.transpose_batch(Batch(self._in_csr, b.rows, B, b.nnz_cap))
        gs, rn2 = self._rows_grad(tb_in, dh1, sl.ldh, "gs")
        gn_ws = ws.get("gn_ws", (256,))
        lib.grad_norm_clip(p(self.gd), sl.n_dense, p(rn2), p(tb_in["n_unique"]), 1e30, p(self.state), p(gn_ws), st)      # (norm only: no clipping)
        lib.adam_step(p(self.arena), p(self.mom), p(self.var), sl.n_params, sl.n_w1, sl.ldh, p(self.slot_of_item), p(gs), p(self.gd),
                      p(self.state), 0.0, self.ADAM_B1, self.ADAM_B2, self.ADAM_EPS, st)
        lib.batch_release(p(tb_in["uniq"]), p(tb_in["n_unique"]), b.nnz_cap, p(self.slot_of_item), p(tb_in["overflow"]), p(self.loss_out),
                          p(self.acc), st)

    def _rows_grad(self, tb, dpre, ld, name):
        """Per touched item: sum over its entries of value * dpre[user, :] (hvae_w1_grad on a matrix of leading dimension ld)."""
        ws, lib = self.ws, self.lib
        n_rows = min(tb["cap"], self.sl.N)
        gs = ws.get(name, (n_rows, ld))
        rn2 = ws.get(name + "_norm2", (tb["cap"],))
        part = ws.get(name + "_partial", (int(lib.w1_max_partial_rows(tb["cap"])), ld))
        lib.w1_grad(p(tb["seg_start"]), p(tb["n_unique"]), p(tb["eid_sorted"]), p(tb["ent_user"]), p(tb["ent_val"]), tb["cap"],
                    p(tb["chunk_base"]), p(tb["part_base"]), p(tb["work_slot"]), p(tb["multi_slot"]), p(tb["n_work"]), p(dpre), ld, 0, 0,
                    p(gs), p(part), p(rn2), self.stream)
        return gs, rn2

    def last_losses(self):
        return tuple(self.loss_out.cpu().tolist())


class MultVAERecommender:
    """src/ml/baseline.py:163-231: fit(interaction_matrix) trains `epochs` epochs (batch 512, shuffled, Adam lr), predict(user)
    returns the user's scores over all items."""

    def __init__(self, hidden_dim: int = 600, latent_dim: int = 200, epochs: int = 50, lr: float = 1e-3, beta: float = 0.2,
                 device=None):
        self.hidden_dim, self.latent_dim, self.epochs, self.lr, self.beta = hidden_dim, latent_dim, epochs, lr, beta
        self.model = None
        self.device = torch.device(device) if device else torch.device("cuda")
        if self.device.type != "cuda":
            raise RuntimeError("hvae_b200.MultVAERecommender needs a CUDA device (there is no CPU fallback)")
        self.matrix = None

    def fit(self, interaction_matrix) -> None:
        logger.info("Fitting Mult-VAE (latent=%d, epochs=%d)...", self.latent_dim, self.epochs)
        self.matrix = interaction_matrix
        n_users, n_items = interaction_matrix.shape
        self.model = MultVAE(n_items, self.hidden_dim, self.latent_dim).to(self.device)
        self._csr = DeviceCSR.from_scipy(interaction_matrix, self.device)
        rt = self.model.rt
        self.model.train()
        lens = self._csr.host_lengths
        for epoch in range(self.epochs):
            order = torch.randperm(n_users).numpy()                 # DataLoader(shuffle=True) over all users (baseline.py:189)
            rt.acc.zero_()
            for s in range(0, n_users, 512):
                rows_h = order[s:s + 512]
                rows = torch.from_numpy(rows_h.astype(np.int32)).to(self.device)
                b = Batch(self._csr, rows, len(rows_h), max(1, int(lens[rows_h].sum())))
                rt.train_step(b, rt.draw_noise(b), self.lr, self.beta)
            if (epoch + 1) % 10 == 0:
                acc = rt.acc.cpu().numpy()
                if np.isnan(acc[0]):
                    self.model.rt._eng.check_overflow()
                logger.info("  Epoch %d/%d, Loss: %.4f", epoch + 1, self.epochs, acc[0] / max(acc[3], 1.0))
        self.model.eval()

    def predict(self, user_idx: int) -> np.ndarray:
        assert self.matrix is not None and self.model is not None, "Model not fitted"
        rows = torch.tensor([user_idx], dtype=torch.int32, device=self.device)
        scores, _, _ = self.model(Batch(self._csr, rows, 1, 1))
        return scores.cpu().numpy().flatten()
