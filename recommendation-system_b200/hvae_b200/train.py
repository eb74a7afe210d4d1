"""Drop-in trainer: UserInteractionDataset, VAETrainer, train_hybrid_vae (reference src/ml/train.py).

The step the reference issues as ~60 ATen calls (zero_grad, forward, vae_loss_function, backward,
clip_grad_norm_, Adam.step, three .item() syncs; src/ml/train.py:86-96) is one fused kernel sequence here
(engine.Engine.train_step), optionally replayed as a CUDA graph, with the loss kept on the device and read
back once per epoch.  Batches stay sparse end to end: `CSRLoader` hands the kernels row ids into a CSR that
is resident in HBM, instead of densifying one [N] row per user on the host (train.py:45-47).
"""
from __future__ import annotations

import json
import logging
import time
from pathlib import Path

import numpy as np
import torch
from scipy.sparse import csr_matrix
from torch.utils.data import DataLoader, Dataset, RandomSampler

from . import data as _data
from ._cabi import STATE_OFF, p
from .engine import Batch, DeviceCSR
from .model import HybridVAE, create_hybrid_vae, vae_loss_function  # noqa: F401  (re-exported like the reference)

logger = logging.getLogger(__name__)

load_training_data = _data.load_training_data
get_user_indices_from_df = _data.get_user_indices_from_df
_build_matrix = _data.build_matrix


class UserInteractionDataset(Dataset):
    """Same contract as the reference (src/ml/train.py:35-47): item i is a dense fp32 [N] row.  The trainer
    never calls __getitem__ on its own datasets -- it reads `.interaction_matrix` / `.user_indices` and goes
    through CSRLoader -- but the dense path stays available for callers that index the dataset directly."""

    def __init__(self, interaction_matrix: csr_matrix, user_indices: list[int] | None = None):
        self.interaction_matrix = interaction_matrix
        self.user_indices = user_indices or list(range(interaction_matrix.shape[0]))

    def __len__(self) -> int:
        return len(self.user_indices)

    def __getitem__(self, idx: int) -> torch.Tensor:
        return torch.FloatTensor(self.interaction_matrix[self.user_indices[idx]].toarray().flatten())


class CSRLoader:
    """Sparse batch iterator over a device-resident CSR.  With shuffle=True the permutation is drawn exactly
    as torch's RandomSampler does (one int64 seed from the global generator, then randperm on a private
    generator), so the batch composition equals the reference DataLoader's under the same torch.manual_seed."""

    def __init__(self, matrix, user_indices=None, batch_size=512, shuffle=False, device="cuda", drop_last=False):
        self.device = torch.device(device)
        self.csr = matrix if isinstance(matrix, DeviceCSR) else DeviceCSR.from_scipy(matrix, self.device)
        self.user_indices = np.arange(self.csr.n_users, dtype=np.int64) if user_indices is None else np.asarray(user_indices, dtype=np.int64)
        self.batch_size, self.shuffle, self.drop_last = int(batch_size), shuffle, drop_last

    def __len__(self):
        n = len(self.user_indices)
        return n // self.batch_size if self.drop_last else (n + self.batch_size - 1) // self.batch_size

    def max_batch_nnz(self, order):
        lens = self.csr.host_lengths[order]
        nb = len(self)
        csum = np.concatenate([[0], np.cumsum(lens)])
        ends = np.minimum(np.arange(1, nb + 1) * self.batch_size, len(order))
        return int((csum[ends] - csum[np.arange(nb) * self.batch_size]).max()) if nb else 1

    def __iter__(self):
        n = len(self.user_indices)
        # every DataLoader iterator -- shuffled or not -- draws its _base_seed from the global CPU generator
        # (torch/utils/data/dataloader.py, _BaseDataLoaderIter.__init__): the validation loader advances the stream too
        torch.empty((), dtype=torch.int64).random_()
        if self.shuffle:
            seed = int(torch.empty((), dtype=torch.int64).random_().item())   # RandomSampler's seed draw
            g = torch.Generator()
            g.manual_seed(seed)
            order = self.user_indices[torch.randperm(n, generator=g).numpy()]
        else:
            order = self.user_indices
        cap = max(1, self.max_batch_nnz(order))
        rows_dev = torch.from_numpy(order.astype(np.int32)).to(self.device, non_blocking=True)
        for i in range(len(self)):
            s, e = i * self.batch_size, min((i + 1) * self.batch_size, n)
            yield Batch(self.csr, rows_dev[s:e], e - s, cap)


class ShardedCSRLoader(CSRLoader):
    """Data-parallel view of CSRLoader: every rank walks the same global batches (same permutation: seed torch
    identically on all ranks) and yields its contiguous slice, tagged with the global batch size so that every
    mean is taken over the global batch (hvae_b200.dist.DataParallel)."""

    def __init__(self, matrix, user_indices=None, batch_size=512, shuffle=False, device="cuda", drop_last=False, dp=None):
        super().__init__(matrix, user_indices, batch_size, shuffle, device, drop_last)
        self.dp = dp

    def __iter__(self):
        from .dist import split_even
        for b in super().__iter__():
            lo, hi = split_even(b.B, self.dp.world, self.dp.rank)
            yield Batch(b.csr, b.rows[lo:hi], hi - lo, b.nnz_cap, b_global=b.B, nnz_cap_global=b.nnz_cap)


class FusedAdamState:
    """`trainer.optimizer`: torch.optim.Adam's hyper-parameters and state_dict() layout over the fused kernel
    (src/ml/train.py:63; checkpoint layout SURVEY.md §8 a12)."""

    def __init__(self, model: HybridVAE, lr, weight_decay):
        self.model = model
        self.defaults = dict(lr=lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=weight_decay, amsgrad=False,
                             maximize=False, foreach=None, capturable=False, differentiable=False, fused=None,
                             decoupled_weight_decay=False)
        self.param_groups = [dict(self.defaults, params=list(range(len(model.layout.reference_keys()))))]

    def zero_grad(self, set_to_none=True):
        return None

    def state_dict(self):
        eng, lay = self.model.engine, self.model.layout
        state = {}
        if eng.m is not None:
            step = float(eng.state[:1].view(torch.int32).item())
            if step > 0:
                for i, k in enumerate(lay.reference_keys()):
                    state[i] = {"step": torch.tensor(step), "exp_avg": lay.view(eng.m, k).clone().contiguous(),
                                "exp_avg_sq": lay.view(eng.v, k).clone().contiguous()}
        return {"state": state, "param_groups": [dict(g) for g in self.param_groups]}

    def load_state_dict(self, sd):
        """Resume support: restores moments and the step count."""
        eng, lay = self.model.engine, self.model.layout
        eng.ensure_optimizer()
        keys = lay.reference_keys()
        step = 0
        for i, k in enumerate(keys):
            if i in sd["state"]:
                lay.view(eng.m, k).copy_(sd["state"][i]["exp_avg"])
                lay.view(eng.v, k).copy_(sd["state"][i]["exp_avg_sq"])
                step = int(sd["state"][i]["step"])
        eng.state[:1].view(torch.int32).fill_(step)
        if sd.get("param_groups"):
            self.param_groups[0].update({k: v for k, v in sd["param_groups"][0].items() if k != "params"})


class VAETrainer:
    """src/ml/train.py:55-145 with the same constructor, methods, return values and attributes."""

    def __init__(self, model: HybridVAE, device: torch.device, lr: float = 0.001, weight_decay: float = 0.0,
                 noise: str = "philox", use_cuda_graph: bool = True):
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("hvae_b200.VAETrainer needs a CUDA device (there is no CPU fallback)")
        self.model = model.to(device)
        self.device = device
        self.lr, self.weight_decay = lr, weight_decay
        self.optimizer = FusedAdamState(self.model, lr, weight_decay)
        self.train_losses: list[float] = []
        self.val_losses: list[float] = []
        self.train_recon_losses: list[float] = []
        self.train_kl_losses: list[float] = []
        self.noise_mode = noise                 # "philox": in-kernel counter RNG; "torch": torch's CUDA generator
        self.use_cuda_graph = use_cuda_graph
        self._noise_seed = int(torch.initial_seed()) & 0x7FFFFFFFFFFFFFFF
        self._noise_ctr = 0
        self._loaders = {}
        self._graphs = {}
        self._loss_host = None      # pinned landing buffer of last_losses()
        self._anneal_dev = 0        # device copy of AnnealedVAE.current_step (hvae_step_state.anneal_step)
        self.graph_collectives = True   # data parallel: capture the NCCL exchange into the step graph as well
        logger.info("Trainer on %s, %s params", device, f"{self.model.num_parameters():,}")

    # -- multi-GPU ---------------------------------------------------------------------------------------------
    def enable_data_parallel(self, group=None):
        """Data-parallel training over torch.distributed (NCCL): see hvae_b200.dist.DataParallel."""
        from .dist import DataParallel
        dp = DataParallel(group)
        self.model.engine.dist = dp
        self._noise_seed = (self._noise_seed + 0x9E3779B97F4A7C15 * (dp.rank + 1)) & 0x7FFFFFFFFFFFFFFF
        return dp

    # -- loaders -------------------------------------------------------------------------------------------------
    def _sparse_loader(self, loader):
        """Map a torch DataLoader over a UserInteractionDataset onto the equivalent CSRLoader."""
        if isinstance(loader, CSRLoader):
            return loader
        ds = getattr(loader, "dataset", None)
        if isinstance(loader, DataLoader) and hasattr(ds, "interaction_matrix") and hasattr(ds, "user_indices") \
                and loader.batch_size is not None:
            key = id(loader)
            if key not in self._loaders:
                self._loaders[key] = CSRLoader(ds.interaction_matrix, ds.user_indices, loader.batch_size,
                                               isinstance(loader.sampler, RandomSampler), self.device, loader.drop_last)
            return self._loaders[key]
        return None

    def _batches(self, loader):
        sl = self._sparse_loader(loader)
        if sl is not None:
            yield from sl
            return
        for x in loader:                       # generic iterable of dense / sparse tensors
            yield self.model._as_batch(x.to(self.device))

    # -- noise ---------------------------------------------------------------------------------------------------
    def _noise_stride(self, B):
        m = self.model
        return (B * max(max(m.hidden_dims), m.embedding_dim, m.latent_dim) + 3) // 4

    def _noise(self, B):
        """Dropout keep-masks and eps for one step.  "philox": counter-based generator inside our kernels; the
        counter base lives in the device step state and is advanced by step_begin, so a captured CUDA graph draws
        fresh noise on every replay.  "torch": torch's CUDA generator in the reference's draw order."""
        m = self.model
        if self.noise_mode == "torch":
            return m._draw_noise(B)
        eng, lay = m.engine, m.layout
        keep = 1.0 - m.dropout
        seed, stp = self._noise_seed, p(eng.state)
        masks = []
        for i, h in enumerate(m.hidden_dims):
            if m.dropout > 0:
                mk = eng.ws.get(f"mask{i}", (B, h), torch.uint8)
                with eng.side(3):
                    eng.lib.fill_noise(p(mk), B * h, keep, None, 0, seed, 0, 1 + i, stp, eng.stream)
                masks.append(mk)
            else:
                masks.append(None)
        eps = eng.ws.get("eps", (B, m.latent_dim))
        with eng.side(4):
            eng.lib.fill_noise(None, 0, keep, p(eps), B * m.latent_dim, seed, 0, 100, stp, eng.stream)
        pmask = None
        if not lay.identity_proj and m.dropout > 0:
            pmask = eng.ws.get("pmask", (B, m.embedding_dim), torch.uint8)
            with eng.side(5):
                eng.lib.fill_noise(p(pmask), B * m.embedding_dim, keep, None, 0, seed, 0, 200, stp, eng.stream)
        return dict(masks=masks, eps=eps, pmask=pmask)

    # -- steps ---------------------------------------------------------------------------------------------------
    def _anneal(self):
        m = self.model
        if hasattr(m, "compute_loss"):       # AnnealedVAE duck-typing, as src/ml/train.py:74
            return dict(beta_min=m.beta_min, beta_max=m.beta_max, anneal_steps=int(m.anneal_steps))
        return dict(beta_min=0.0, beta_max=m.beta, anneal_steps=0)

    def _sync_anneal_step(self):
        """AnnealedVAE.current_step is a host attribute (src/ml/model.py:310); the kernels keep their own copy in the
        device step state and advance it themselves.  Write it only when the host value was changed from outside."""
        m, eng = self.model, self.model.engine
        if hasattr(m, "current_step") and int(m.current_step) != self._anneal_dev:
            eng.state[STATE_OFF["anneal_step"]:STATE_OFF["anneal_step"] + 1].view(torch.int32).fill_(int(m.current_step))
            self._anneal_dev = int(m.current_step)

    def _eager_step(self, batch: Batch, noise, b_global):
        eng = self.model.engine
        if noise is not None:       # injected noise (parity tests): nothing to generate
            eng.train_step(batch, noise, lr=self.lr, weight_decay=self.weight_decay, b_global=b_global, **self._anneal())
            return
        # step scalars first (advances the device-side noise counter), then the noise kernels fork onto side streams
        bg = batch.B if b_global is None else b_global
        eng.begin(bg, self.lr, advance=True, noise_stride=self._noise_stride(batch.B), **self._anneal())
        noise = self._noise(batch.B)
        eng.train_step(batch, noise, lr=self.lr, weight_decay=self.weight_decay, b_global=b_global, begun=True, **self._anneal())

    def _graph_entry(self, kind, batch: Batch, b_global):
        """Static input buffers + captured graph of one step for this (batch size, nnz bound)."""
        eng = self.model.engine
        cap = max(1024, -(-int(batch.nnz_cap * 1.05) // 1024) * 1024)        # bucketed so that re-captures are rare
        key = (kind, batch.B, b_global, self.lr, self.weight_decay)
        ent = self._graphs.get(key)
        capg_need = getattr(batch, "nnz_cap_global", None) or 0
        if (ent is not None and ent["cap"] >= batch.nnz_cap and ent["cap_global"] >= capg_need
                and ent["gen"] == eng.ws.generation):
            return ent
        dev, B = self.device, batch.B
        capg = max(1024, -(-int(capg_need * 1.05) // 1024) * 1024) if capg_need else None
        ent = {"cap": cap, "cap_global": capg or 0, "graph": None, "launches": 0, "gen": eng.ws.generation}
        if kind == "rows":      # batch = user ids into the resident CSR
            ent["rows"] = torch.zeros(B, dtype=torch.int32, device=dev)
            ent["batch"] = Batch(batch.csr, ent["rows"], B, cap, b_global=batch.b_global, nnz_cap_global=capg)
            ent["csr_id"] = id(batch.csr)
        else:                   # batch = its own CSR slice copied from the host every step
            ent["crow"] = torch.zeros(B + 1, dtype=torch.int64, device=dev)
            ent["col"] = torch.zeros(cap, dtype=torch.int32, device=dev)
            ent["val"] = torch.ones(cap, dtype=torch.float32, device=dev)
            ent["batch"] = Batch(DeviceCSR(ent["crow"], ent["col"], ent["val"], B, self.model.n_items, None), None, B, cap)
        self._graphs[key] = ent
        return ent

    def _load_static(self, ent, batch: Batch):
        if "rows" in ent:
            ent["rows"].copy_(batch.rows, non_blocking=True)
        else:
            nnz = batch.csr.indices.shape[0]
            ent["crow"].copy_(batch.csr.indptr, non_blocking=True)
            ent["col"][:nnz].copy_(batch.csr.indices, non_blocking=True)
            if batch.csr.values is None:
                ent["val"][:nnz].fill_(1.0)
            else:
                ent["val"][:nnz].copy_(batch.csr.values, non_blocking=True)

    def _graphed_step(self, batch: Batch, b_global):
        eng = self.model.engine
        kind = "rows" if batch.rows is not None else "csr"
        ent = self._graph_entry(kind, batch, b_global)
        if kind == "rows" and ent.get("csr_id") != id(batch.csr):
            self._graphs.pop((kind, batch.B, b_global, self.lr, self.weight_decay), None)
            ent = self._graph_entry(kind, batch, b_global)
        self._load_static(ent, batch)
        self._run_entry(ent, b_global)

    def _capture_stream(self):
        """High-priority capture stream: the step's critical path outranks the low-priority Adam branch (Engine.side(7))."""
        if getattr(self, "_cap_stream", None) is None:
            self._cap_stream = torch.cuda.Stream(self.device, priority=-1)
        return self._cap_stream

    def _run_entry(self, ent, b_global):
        eng = self.model.engine
        if ent["graph"] is None:
            # first use: one eager step on the static buffers (allocates every workspace), then capture for the next ones
            self._eager_step(ent["batch"], None, b_global)
            torch.cuda.synchronize(self.device)
            g = torch.cuda.CUDAGraph()
            l0 = eng.lib.launches
            eng.concurrent = True           # independent kernels -> parallel branches of the graph
            try:
                with torch.cuda.graph(g, stream=self._capture_stream()):
                    self._eager_step(ent["batch"], None, b_global)
            finally:
                eng.concurrent = False
            ent["launches"] = eng.lib.launches - l0
            eng.lib.launches = l0
            ent["graph"], ent["gen"] = g, eng.ws.generation
            return
        ent["graph"].replay()
        eng.lib.launches += ent["launches"]

    def train_step(self, batch: Batch, noise=None, b_global=None):
        """One optimisation step on a sparse batch; `noise` overrides the trainer's RNG (parity tests)."""
        m = self.model
        eng = m.engine
        self._sync_anneal_step()
        graphable = (noise is None and self.use_cuda_graph and self.noise_mode == "philox" and eng.prof is None
                     and (eng.dist is None or (self.graph_collectives and batch.rows is not None)))
        if graphable:
            self._graphed_step(batch, b_global)
        else:
            self._eager_step(batch, noise, b_global)
        if hasattr(m, "current_step"):
            m.step_annealing()
            self._anneal_dev += 1

    def set_interactions(self, matrix):
        """Keep the user x item CSR resident in HBM so that a batch can be named by its user ids alone."""
        self._resident = matrix if isinstance(matrix, DeviceCSR) else DeviceCSR.from_scipy(matrix, self.device)
        return self._resident

    def _host_batch(self, x, b_global=None, nnz_cap_global=None) -> Batch:
        """HOST (ideally pinned) or device batch -> device Batch.  A 1-D integer tensor is a list of user ids into the
        resident CSR (set_interactions); sparse CSR tensors are copied as their three arrays (a few KB per step);
        dense [B,N] rows are copied whole and compacted on the device."""
        if isinstance(x, Batch):
            return x
        if isinstance(x, torch.Tensor) and x.dim() == 1 and not x.dtype.is_floating_point:
            csr = getattr(self, "_resident", None)
            if csr is None:
                raise RuntimeError("user-id batches need trainer.set_interactions(matrix) first")
            rows_h = x.cpu().numpy() if x.is_cuda else x.numpy()
            cap = int(csr.host_lengths[rows_h].sum())
            rows_d = x.to(self.device, torch.int32, non_blocking=True)
            dp = self.model.engine.dist
            if dp is not None and nnz_cap_global is None:      # bound of the global batch's nnz: world * max over ranks
                t = torch.tensor([cap], dtype=torch.int64, device=self.device)
                torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX, group=dp.group)
                nnz_cap_global = int(t.item()) * dp.world
            return Batch(csr, rows_d, x.shape[0], max(1, cap), b_global=b_global, nnz_cap_global=nnz_cap_global)
        if isinstance(x, torch.Tensor) and x.layout == torch.sparse_csr:
            self._check_host_csr(x)
            crow, col, val = x.crow_indices(), x.col_indices(), x.values()
            nnz = int(col.shape[0])
            dev = self.device
            crow_d = crow.to(dev, torch.int64, non_blocking=True)
            col_d = (col if col.dtype == torch.int32 else col.to(torch.int32)).to(dev, non_blocking=True)
            val_d = (val if val.dtype == torch.float32 else val.float()).to(dev, non_blocking=True)
            csr = DeviceCSR(crow_d, col_d, val_d, x.shape[0], x.shape[1], None)
            return Batch(csr, None, x.shape[0], max(1, nnz))
        return self.model._as_batch(x.to(self.device, non_blocking=True))

    def _check_host_csr(self, x):
        """Shape / structure checks of a caller-supplied CSR batch before it reaches the kernels (an out-of-range column
        would index W1^T and slot_of_item out of bounds).  Host batches only: a few KB, microseconds."""
        crow, col = x.crow_indices(), x.col_indices()
        if x.dim() != 2 or x.shape[1] != self.model.n_items:
            raise ValueError(f"expected a [B, {self.model.n_items}] batch, got {tuple(x.shape)}")
        if crow.is_cuda:
            return
        nnz = int(col.shape[0])
        if crow.shape[0] != x.shape[0] + 1 or int(crow[0]) != 0 or int(crow[-1]) != nnz:
            raise ValueError("malformed CSR batch: crow_indices must run from 0 to nnz over B+1 entries")
        if nnz and (int(col.min()) < 0 or int(col.max()) >= self.model.n_items):
            raise ValueError(f"CSR batch has column indices outside [0, {self.model.n_items})")

    def _host_csr_step(self, x, b_global):
        """Host CSR batch -> static device buffers (one H2D copy per array) -> replay of the captured step."""
        self._check_host_csr(x)
        crow, col, val = x.crow_indices(), x.col_indices(), x.values()
        nnz, B = int(col.shape[0]), x.shape[0]
        ent = self._graph_entry("csr", Batch(None, None, B, max(1, nnz)), b_global)
        if (crow.dtype == torch.int64 and col.dtype == torch.int32 and val.dtype == torch.float32 and crow.is_contiguous()
                and col.is_contiguous() and val.is_contiguous()):
            eng = self.model.engine            # the three copies behind one C call (hvae_h2d_csr_batch)
            eng.lib.h2d_csr_batch(crow.data_ptr(), col.data_ptr(), val.data_ptr(), B, nnz, p(ent["crow"]), p(ent["col"]), p(ent["val"]),
                                  eng.stream)
        else:
            ent["crow"].copy_(crow, non_blocking=True)
            ent["col"][:nnz].copy_(col, non_blocking=True)
            ent["val"][:nnz].copy_(val, non_blocking=True)
        self._run_entry(ent, b_global)

    def train_on_batch(self, x, b_global=None, nnz_cap_global=None) -> dict[str, float]:
        """One iteration of the reference loop (src/ml/train.py:86-96) on one batch: host->device copy of the
        batch, the fused step, and the three loss scalars read back (the reference's three .item() calls)."""
        self.model.train()
        eng = self.model.engine
        if (isinstance(x, torch.Tensor) and x.layout == torch.sparse_csr and not x.is_cuda and self.use_cuda_graph
                and self.noise_mode == "philox" and eng.dist is None and eng.prof is None):
            self._sync_anneal_step()
            self._host_csr_step(x, b_global)
            if hasattr(self.model, "current_step"):
                self.model.step_annealing()
                self._anneal_dev += 1
        else:
            self.train_step(self._host_batch(x, b_global, nnz_cap_global), b_global=b_global)
        if eng.dist is not None:          # per-rank partial sums (already scaled by 1/B_global) -> global loss
            torch.distributed.all_reduce(eng.loss_out, group=eng.dist.group)
        total, recon, kl = self.last_losses()
        return {"total_loss": total, "recon_loss": recon, "kl_loss": kl}

    def train_epoch(self, loader) -> dict[str, float]:
        """src/ml/train.py:81-103: mean of the per-batch (total, recon, kl) over the epoch."""
        self.model.train()
        eng = self.model.engine
        eng.acc.zero_()
        for batch in self._batches(loader):
            self.train_step(batch, b_global=getattr(batch, "b_global", None))
        if eng.dist is not None:
            eng.dist.reduce_losses(eng.acc)
        acc = eng.acc.cpu().numpy().astype(np.float64)     # the epoch's only device->host read
        if np.isnan(acc[0]):
            eng.check_overflow()
        n = max(1.0, acc[3])
        return {"total_loss": float(acc[0] / n), "recon_loss": float(acc[1] / n), "kl_loss": float(acc[2] / n)}

    def validate(self, loader) -> dict[str, float]:
        """src/ml/train.py:105-124: eval mode, fixed model.beta."""
        self.model.eval()
        eng = self.model.engine
        eng.acc.zero_()
        with torch.no_grad():
            for batch in self._batches(loader):
                eng.eval_step(batch, float(self.model.beta))
        acc = eng.acc.cpu().numpy().astype(np.float64)
        n = max(1.0, acc[3])
        return {"total_loss": float(acc[0] / n), "recon_loss": float(acc[1] / n), "kl_loss": float(acc[2] / n)}

    def last_losses(self):
        """(total, recon, kl) of the most recent step (device->host read into pinned memory; waits for the step)."""
        eng = self.model.engine
        if self._loss_host is None:
            self._loss_host = torch.empty(3, dtype=torch.float32).pin_memory()
        eng.lib.d2h_floats(p(eng.loss_out), 3, self._loss_host.data_ptr(), eng.stream)
        out = tuple(self._loss_host.tolist())
        if out[0] != out[0]:          # NaN: either a genuine one or a step whose batch overflowed its nnz bound
            eng.check_overflow()
        return out

    def save_checkpoint(self, path, epoch: int, is_best: bool = False, extra: dict | None = None) -> None:
        """src/ml/train.py:126-145: same keys; tensors in the reference's shapes."""
        checkpoint = {
            "epoch": epoch,
            "model_state_dict": self.model.state_dict(),
            "optimizer_state_dict": self.optimizer.state_dict(),
            "train_losses": self.train_losses,
            "val_losses": self.val_losses,
            "train_recon_losses": self.train_recon_losses,
            "train_kl_losses": self.train_kl_losses,
            **(extra or {}),
        }
        if hasattr(self.model, "current_step"):
            checkpoint.setdefault("annealing_current_step", int(self.model.current_step))
        torch.save(checkpoint, path)
        if is_best:
            best_path = Path(path).parent / "best_model.pth"
            torch.save(checkpoint, best_path)
            logger.info("Saved best model to %s", best_path)


def resume_from_checkpoint(trainer: "VAETrainer", path) -> int:
    """Restore model weights, Adam moments / step count, loss histories and the annealing step from a checkpoint written by
    save_checkpoint (the reference saves optimizer_state_dict but never reloads it).  Returns the epoch."""
    ck = torch.load(path, map_location="cpu", weights_only=False)
    trainer.model.load_state_dict(ck["model_state_dict"])
    trainer.model.to(trainer.device)
    if ck.get("optimizer_state_dict", {}).get("state"):
        trainer.optimizer.load_state_dict(ck["optimizer_state_dict"])
    for k in ("train_losses", "val_losses", "train_recon_losses", "train_kl_losses"):
        setattr(trainer, k, list(ck.get(k, [])))
    if hasattr(trainer.model, "current_step") and "annealing_current_step" in ck:
        trainer.model.current_step = int(ck["annealing_current_step"])
    trainer._graphs.clear()
    return int(ck.get("epoch", 0))


def _get_device(device: str | None = None) -> torch.device:
    if device:
        return torch.device(device)
    if torch.cuda.is_available():
        return torch.device("cuda")
    raise RuntimeError("hvae_b200 needs a CUDA device (there is no CPU fallback)")


class _EarlyStop:
    """Best validation loss so far + epochs since it improved."""

    def __init__(self, patience: int):
        self.patience, self.best, self.stale = patience, float("inf"), 0

    def improved(self, value: float) -> bool:
        better = value < self.best
        self.best, self.stale = (value, 0) if better else (self.best, self.stale + 1)
        return better

    @property
    def exhausted(self) -> bool:
        return self.stale >= self.patience


def _write_history(trainer: "VAETrainer", path: Path, seconds: float) -> None:
    """training_history.json as the reference writes it (src/ml/train.py:325-336): same keys, same rounding."""
    doc = {k: getattr(trainer, k) for k in ("train_losses", "val_losses", "train_recon_losses", "train_kl_losses")}
    doc["training_time_seconds"] = round(seconds, 2)
    path.write_text(json.dumps(doc, indent=2))


def _split_loaders(data_dir: str, batch_size: int, dev):
    """(train loader, val loader, n_items) from the offline pipeline's files: each split keeps ONLY its own interactions."""
    full_matrix, train_df, val_df, mappings = load_training_data(data_dir)
    u2i, i2i = mappings["user_to_idx"], mappings["item_to_idx"]
    loaders = [CSRLoader(_build_matrix(df, u2i, i2i, full_matrix.shape), get_user_indices_from_df(df, u2i), batch_size, shuffle, dev)
               for df, shuffle in ((train_df, True), (val_df, False))]
    return loaders[0], loaders[1], full_matrix.shape[1]


def train_hybrid_vae(data_dir: str, embeddings_path: str, output_dir: str, latent_dim: int = 200,
                     hidden_dims: list[int] | None = None, batch_size: int = 512, epochs: int = 100,
                     learning_rate: float = 0.001, weight_decay: float = 0.0, beta: float = 0.2, dropout: float = 0.5,
                     use_annealing: bool = False, patience: int = 10, device: str | None = None,
                     ignore_embeddings: bool = False, precision: str | None = None) -> None:
    """The reference's training driver as a file contract (src/ml/train.py:201-342): same inputs, same outputs
    (checkpoint_epoch_{e}.pth, best_model.pth, training_history.json), early stopping on the validation loss."""
    dev = _get_device(device)
    out = Path(output_dir)
    out.mkdir(parents=True, exist_ok=True)
    train_loader, val_loader, n_items = _split_loaders(data_dir, batch_size, dev)
    emb_file = Path(embeddings_path)
    embeddings, emb_items, _ = _data.load_embeddings(embeddings_path, str(emb_file.with_name(f"{emb_file.stem}_mappings.pkl")))
    if not emb_items or len(emb_items) != n_items:
        raise AssertionError(f"Embedding mismatch: {len(emb_items) if emb_items else 0} vs {n_items}")
    if ignore_embeddings:          # ablation switch of the reference: random vectors of the same shape
        embeddings = np.random.normal(0, 0.01, embeddings.shape).astype(np.float32)
    model = create_hybrid_vae(n_items=n_items, item_embeddings=embeddings, latent_dim=latent_dim, hidden_dims=hidden_dims,
                              dropout=dropout, beta=beta, use_annealing=use_annealing,
                              anneal_steps=int(len(train_loader) * epochs * 0.5), precision=precision)
    trainer = VAETrainer(model, dev, learning_rate, weight_decay)
    config = {"n_items": n_items, "latent_dim": latent_dim, "hidden_dims": hidden_dims, "beta": beta, "dropout": dropout}
    stop, t0 = _EarlyStop(patience), time.time()
    for epoch in range(1, epochs + 1):
        tm, vm = trainer.train_epoch(train_loader), trainer.validate(val_loader)
        for series, value in ((trainer.train_losses, tm["total_loss"]), (trainer.val_losses, vm["total_loss"]),
                              (trainer.train_recon_losses, tm["recon_loss"]), (trainer.train_kl_losses, tm["kl_loss"])):
            series.append(value)
        logger.info("Epoch %d/%d train %.4f (recon %.4f, kl %.4f) val %.4f", epoch, epochs, tm["total_loss"], tm["recon_loss"],
                    tm["kl_loss"], vm["total_loss"])
        trainer.save_checkpoint(out / f"checkpoint_epoch_{epoch}.pth", epoch, stop.improved(vm["total_loss"]),
                                extra={"train_metrics": tm, "val_metrics": vm, "model_config": config})
        if stop.exhausted:
            logger.info("Early stopping at epoch %d", epoch)
            break
    _write_history(trainer, out / "training_history.json", time.time() - t0)
    logger.info("Training complete. Best val loss %.4f in %.1fs", stop.best, time.time() - t0)
