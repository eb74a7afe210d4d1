"""CPU: the C-ABI library loads and exports every symbol include/hvae_b200.h declares; the drop-in classes
keep the reference's constructor / state_dict contract; host-side logic (layout, loaders, metric functions)."""
import ctypes

import numpy as np
import pytest
import torch

from golden_util import TRAIN_CASES, Case
from hvae_b200 import _cabi
from hvae_b200.engine import Layout
from hvae_b200.model import AnnealedVAE, HybridVAE, create_hybrid_vae, vae_loss_function
from oracle import hvae_oracle as orc


def test_library_exports_every_declared_symbol():
    decls = _cabi.parse_header()
    assert len(decls) >= 30
    dll = ctypes.CDLL(str(_cabi.LIB_PATH))
    for name in decls:
        assert hasattr(dll, name), name
    assert _cabi.lib().abi_version() == 2
    assert ctypes.sizeof(_cabi.StepState) == 48


@pytest.mark.parametrize("name", TRAIN_CASES)
def test_init_and_state_dict_match_reference_layout(name):
    """Same torch.manual_seed -> same initial weights as the reference class, bit for bit; same keys, order, shapes."""
    c = Case(name)
    torch.manual_seed(c.seed)
    m = create_hybrid_vae(**c.model_kwargs(), use_annealing=c.annealing, anneal_steps=4)
    ref = c.state("init")
    sd = m.state_dict()
    assert list(sd.keys()) == list(ref.keys())
    for k in ref:
        assert sd[k].shape == ref[k].shape, k
        assert torch.equal(sd[k], ref[k]), k
    # round trip through load_state_dict
    m2 = create_hybrid_vae(**c.model_kwargs())
    m2.load_state_dict(c.state("final"))
    for k, v in c.state("final").items():
        assert torch.equal(m2.state_dict()[k], v), k
    # strict loading rejects foreign keys / wrong shapes like nn.Module does
    bad = dict(c.state("final"))
    bad["bogus"] = torch.zeros(1)
    with pytest.raises(RuntimeError):
        m2.load_state_dict(bad)
    bad = dict(c.state("final"))
    bad["fc_mu.weight"] = torch.zeros(3, 3)
    with pytest.raises(RuntimeError):
        m2.load_state_dict(bad)
    # the oracle (reference module tree) accepts our state_dict unchanged
    o = orc.OracleVAE(**c.model_kwargs())
    o.load_state_dict(m2.state_dict())


def test_constructor_contract():
    E = np.random.default_rng(0).standard_normal((50, 16)).astype(np.float32)
    m = HybridVAE(n_items=50, item_embeddings=E, latent_dim=8, hidden_dims=[12], dropout=0.1, beta=0.3)
    assert (m.n_items, m.latent_dim, m.embedding_dim, m.dropout, m.beta, m.hidden_dims) == (50, 8, 16, 0.1, 0.3, [12])
    assert HybridVAE(50, E).hidden_dims == [600, 200]
    a = create_hybrid_vae(50, E, latent_dim=8, hidden_dims=[12], use_annealing=True, anneal_steps=10, beta=0.4)
    assert isinstance(a, AnnealedVAE) and a.beta_max == 0.4 and a.beta_min == 0.0 and a.current_step == 0
    betas = []
    for _ in range(12):
        betas.append(a.get_current_beta())
        a.step_annealing()
    assert betas[0] == 0.0 and abs(betas[5] - 0.2) < 1e-12 and betas[10] == 0.4 and betas[11] == 0.4
    # annealing kwargs silently dropped when use_annealing=False (src/ml/model.py:376-385)
    assert type(create_hybrid_vae(50, E, latent_dim=8, hidden_dims=[12], anneal_steps=3)) is HybridVAE
    with pytest.raises(NotImplementedError):
        HybridVAE(50, E, freeze_embeddings=False)
    with pytest.raises(ValueError):
        HybridVAE(49, E)
    # identity projection when latent == embedding dim: no projection keys
    assert not any(k.startswith("projection") for k in HybridVAE(50, E, latent_dim=16, hidden_dims=[12]).state_dict())


def test_no_cpu_fallback():
    E = np.eye(8, dtype=np.float32)
    m = HybridVAE(8, E, latent_dim=4, hidden_dims=[6])
    with pytest.raises(RuntimeError, match="CUDA"):
        m(torch.zeros(2, 8))
    from hvae_b200.train import VAETrainer
    with pytest.raises(RuntimeError, match="CUDA"):
        VAETrainer(m, torch.device("cpu"))


def test_loss_function_matches_oracle():
    g = torch.Generator().manual_seed(0)
    s, x = torch.randn(5, 9, generator=g), (torch.rand(5, 9, generator=g) > 0.6).float()
    mu, lv = torch.randn(5, 3, generator=g), torch.randn(5, 3, generator=g)
    for a, b in zip(vae_loss_function(s, x, mu, lv, 0.3), orc.loss_terms(s, x, mu, lv, 0.3)):
        assert torch.equal(a, b)


def test_layout_padding_and_offsets():
    lay = Layout(101, 30, 10, [37, 22])
    assert lay.n_w1 == 101 * 40 and lay.n_params % 4 == 0
    for s in lay.slots.values():
        assert s.off % 4 == 0 and s.ld % 4 == 0
    arena = torch.arange(lay.n_params, dtype=torch.float32)
    assert lay.view(arena, "encoder.0.weight").shape == (37, 101)
    assert lay.view(arena, "encoder.4.weight").shape == (22, 37)
    assert lay.view(arena, "fc_logvar.bias").shape == (10,)
    assert lay.view(arena, "projection_layer.3.weight").shape == (30, 30)


def test_csr_loader_matches_dataloader_order():
    """Shuffled batches equal the reference DataLoader's under the same torch.manual_seed (RandomSampler draws)."""
    from scipy.sparse import random as sprand
    from hvae_b200.train import CSRLoader, UserInteractionDataset
    m = sprand(40, 30, density=0.2, format="csr", random_state=0)
    users = list(range(3, 37))
    ds = UserInteractionDataset(m, users)
    torch.manual_seed(5)
    ref_batches = [b for b in torch.utils.data.DataLoader(ds, batch_size=8, shuffle=True)]
    torch.manual_seed(5)
    ld = CSRLoader(m, users, 8, True, device="cpu")
    ours = list(ld)
    assert len(ours) == len(ref_batches) == len(ld)
    for ob, rb in zip(ours, ref_batches):
        dense = torch.from_numpy(np.asarray(m[ob.rows.numpy()].toarray(), dtype=np.float32))
        assert torch.equal(dense, rb)
        assert ob.nnz_cap >= int((dense != 0).sum())


def test_csr_loader_multi_epoch_train_validate_order():
    """Three epochs of train (shuffled) + validate (not shuffled): every DataLoader iterator -- the validation one too --
    draws a _base_seed from the global generator, so from epoch 2 on the permutations only match if CSRLoader draws it as well."""
    from scipy.sparse import random as sprand
    from hvae_b200.train import CSRLoader, UserInteractionDataset
    m = sprand(50, 30, density=0.2, format="csr", random_state=1)
    tr_users, va_users = list(range(0, 41)), list(range(5, 33))
    torch.manual_seed(11)
    ref_tr = torch.utils.data.DataLoader(UserInteractionDataset(m, tr_users), batch_size=16, shuffle=True)
    ref_va = torch.utils.data.DataLoader(UserInteractionDataset(m, va_users), batch_size=16, shuffle=False)
    ref = []
    for _ in range(3):
        ref.append([b.clone() for b in ref_tr])
        ref.append([b.clone() for b in ref_va])
    torch.manual_seed(11)
    tr, va = CSRLoader(m, tr_users, 16, True, device="cpu"), CSRLoader(m, va_users, 16, False, device="cpu")
    ours = []
    for _ in range(3):
        ours.append(list(tr))
        ours.append(list(va))
    for o_ep, r_ep in zip(ours, ref):
        assert len(o_ep) == len(r_ep)
        for ob, rb in zip(o_ep, r_ep):
            assert torch.equal(torch.from_numpy(np.asarray(m[ob.rows.numpy()].toarray(), dtype=np.float32)), rb)


def test_metric_functions_match_oracle():
    from hvae_b200 import evaluate as ev
    rng = np.random.default_rng(0)
    for _ in range(50):
        rec = rng.permutation(40)[:20]
        rel = rng.integers(0, 40, rng.integers(0, 4))
        for k in (1, 5, 20):
            assert ev.recall_at_k(rec, rel, k) == orc.recall_at_k(rec, rel, k)
            assert ev.ndcg_at_k(rec, rel, k) == orc.ndcg_at_k(rec, rel, k)
            assert ev.hit_ratio_at_k(rec, rel, k) == orc.hit_ratio_at_k(rec, rel, k)


def test_negative_sampling_candidates():
    """99 distinct unseen negatives per row, test item first; short catalogues take what is available
    (reference src/ml/evaluate.py:160-172)."""
    from hvae_b200.sampling import sample_negatives
    from hvae_b200.synth import make_interactions
    for (U, N) in [(300, 400), (20, 40)]:
        d = make_interactions(U, N, 1)
        cand, valid = sample_negatives(d.indptr, d.indices, N, np.arange(U), d.test_items, 99, np.random.default_rng(0))
        assert cand.shape == (U, 100) and np.array_equal(cand[:, 0], d.test_items)
        for u in range(U):
            seen = set(d.indices[d.indptr[u]:d.indptr[u + 1]].tolist())
            neg = cand[u, 1:1 + valid[u]].tolist()
            assert len(set(neg)) == len(neg) and not (set(neg) & seen) and d.test_items[u] not in neg
            assert valid[u] == min(99, N - len(seen) - 1)


def test_torch_custom_ops_registered_and_cuda_only():
    """torch.ops.hvae_b200.* exist, infer shapes on fake tensors, and refuse CPU tensors (no fallback)."""
    import hvae_b200.ops  # noqa: F401  (registers the ops)
    for name in ("gather_ln_fwd", "gemm", "score_lse", "score_grad", "score_topk", "adam_step"):
        assert hasattr(torch.ops.hvae_b200, name), name
    with pytest.raises(RuntimeError, match="CUDA"):
        torch.ops.hvae_b200.score_lse(torch.zeros(4, 8, dtype=torch.bfloat16), torch.zeros(10, 8, dtype=torch.bfloat16), 8)
    with pytest.raises(RuntimeError, match="CUDA"):
        torch.ops.hvae_b200.gemm(torch.zeros(4, 8), torch.zeros(6, 8), None, True)
    from torch._subclasses.fake_tensor import FakeTensorMode
    with FakeTensorMode():
        U, E = torch.empty(5, 16, dtype=torch.bfloat16), torch.empty(40, 16, dtype=torch.bfloat16)
        assert torch.ops.hvae_b200.score_lse(U, E, 12).shape == (5,)
        assert torch.ops.hvae_b200.score_grad(U, E, torch.empty(5), 12).shape == (5, 12)
        v, i = torch.ops.hvae_b200.score_topk(U, E, 12, torch.empty(6, dtype=torch.int64), torch.empty(9, dtype=torch.int32),
                                              torch.empty(5, dtype=torch.int32), 7, 0)
        assert v.shape == (5, 7) and i.dtype == torch.int32
        assert torch.ops.hvae_b200.gemm(torch.empty(5, 16), torch.empty(9, 16), None, True).shape == (5, 9)


def test_device_negative_sampler_on_cpu_tensors():
    """sampling.sample_negatives_device (torch index ops; runs on any device): test item first, 99 distinct unseen negatives,
    short catalogues fall back to whatever is available (reference src/ml/evaluate.py:160-172)."""
    from hvae_b200.engine import DeviceCSR
    from hvae_b200.sampling import sample_negatives_device
    from hvae_b200.synth import make_interactions
    for (U, N) in [(200, 12101), (20, 40)]:
        d = make_interactions(U, N, 1)
        csr = DeviceCSR.from_arrays(d.indptr, d.indices, None, N, torch.device("cpu"))
        cand, valid = sample_negatives_device(csr, np.arange(U), d.test_items, 99, seed=3)
        cand = cand.numpy()
        assert cand.shape == (U, 100) and np.array_equal(cand[:, 0], d.test_items)
        for u in range(U):
            seen = set(d.indices[d.indptr[u]:d.indptr[u + 1]].tolist())
            neg = cand[u, 1:1 + valid[u]].tolist()
            assert len(set(neg)) == len(neg) and not (set(neg) & seen) and d.test_items[u] not in neg
            assert valid[u] == min(99, N - len(seen) - 1)
