"""GPU: the tcgen05 scoring kernels (bf16 mode) against plain torch on the same bf16-rounded operands.
The kernels accumulate in fp32, so the only difference from `U.float() @ E.float().T` is summation order."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def _r8(n):
    return (n + 7) // 8 * 8


def _operands(B, N, d, dev, seed=0, scale=3.0):
    g = torch.Generator().manual_seed(seed)
    U = (torch.randn(B, d, generator=g) * scale / d ** 0.5 * 4).to(dev)
    E = torch.nn.functional.normalize(torch.randn(N, d, generator=g), dim=1).to(dev)
    return U, E


def _cast(lib, x, dev):
    rows, cols = x.shape
    ld = _r8(cols)
    out = torch.empty(rows, ld, dtype=torch.bfloat16, device=dev)
    lib.cast_bf16(x.data_ptr(), rows, cols, cols, out.data_ptr(), ld, torch.cuda.current_stream().cuda_stream)
    return out, ld


SHAPES = [(512, 12101, 384), (128, 256, 64), (77, 1000, 64), (300, 5000, 768), (130, 890, 384), (64, 300, 24), (1, 513, 200)]


@pytest.mark.parametrize("B,N,d", SHAPES)
def test_tc_lse_matches_torch(dev, B, N, d):
    from hvae_b200 import _cabi
    lib = _cabi.lib()
    st = torch.cuda.current_stream().cuda_stream
    U, E = _operands(B, N, d, dev)
    Ub, ldu = _cast(lib, U.contiguous(), dev)
    Eb, lde = _cast(lib, E.contiguous(), dev)
    assert torch.equal(Ub[:, :d], U.to(torch.bfloat16)) and torch.all(Ub[:, d:] == 0)
    ns = int(lib.tc_n_splits(B, N))
    ws = torch.empty(2 * B * ns, device=dev)
    lse = torch.empty(B, device=dev)
    lib.tc_score_lse(Ub.data_ptr(), ldu, B, Eb.data_ptr(), lde, N, d, lse.data_ptr(), ws.data_ptr(), st)
    ref = torch.logsumexp(Ub[:, :d].float().double() @ Eb[:, :d].float().double().t(), dim=1)
    np.testing.assert_allclose(lse.cpu().numpy(), ref.cpu().numpy(), rtol=2e-6, atol=2e-5)


@pytest.mark.parametrize("B,N,d,K", [(512, 12101, 384, 20), (77, 1000, 64, 10), (300, 5000, 768, 32), (130, 890, 384, 5), (3, 40, 24, 20),
                                     (300, 5000, 768, 100), (77, 1000, 64, 64), (512, 12101, 384, 128), (200, 9000, 200, 33), (64, 70000, 64, 24)])
def test_tc_topk_matches_torch(dev, B, N, d, K):
    from hvae_b200 import _cabi
    from hvae_b200.synth import make_interactions
    lib = _cabi.lib()
    st = torch.cuda.current_stream().cuda_stream
    U, E = _operands(B, N, d, dev, seed=1)
    Ub, ldu = _cast(lib, U.contiguous(), dev)
    Eb, lde = _cast(lib, E.contiguous(), dev)
    data = make_interactions(B, N, 3)
    indptr = torch.from_numpy(data.indptr).to(dev)
    indices = torch.from_numpy(data.indices).to(dev)
    S = (Ub[:, :d].float() @ Eb[:, :d].float().t())
    Sm = S.clone()
    rr = torch.from_numpy(np.repeat(np.arange(B), np.diff(data.indptr))).to(dev)
    Sm[rr, indices.long()] = -float("inf")
    for (lo, hi) in [(0, N), (N // 3, N - 7)]:
        n_it = hi - lo
        ns = int(lib.tc_n_splits(B, n_it))
        cv = torch.empty(B, ns * K, device=dev)
        ci = torch.empty(B, ns * K, dtype=torch.int32, device=dev)
        e_ptr = Eb.data_ptr() + 2 * lo * lde
        lib.tc_score_topk(Ub.data_ptr(), ldu, B, e_ptr, lde, n_it, d, lo, indptr.data_ptr(), indices.data_ptr(), None, K,
                          cv.data_ptr(), ci.data_ptr(), st)
        ov = torch.empty(B, K, device=dev)
        oi = torch.empty(B, K, dtype=torch.int32, device=dev)
        lib.topk_merge(cv.data_ptr(), ci.data_ptr(), B, ns * K, K, ov.data_ptr(), oi.data_ptr(), st)
        kk = min(K, n_it)
        rv, ri = torch.topk(Sm[:, lo:hi], kk, dim=1)
        got_v, got_i = ov.cpu().numpy(), oi.cpu().numpy()
        ref_v = rv.cpu().numpy()
        finite = np.isfinite(ref_v)
        np.testing.assert_allclose(got_v[:, :kk][finite], ref_v[finite], rtol=1e-5, atol=1e-5)
        # the returned ids carry the scores they claim, are unseen, and are inside the shard
        Sm_c = Sm.cpu().numpy()
        for b in range(B):
            ids = got_i[b, :kk][finite[b]]
            assert len(set(ids.tolist())) == len(ids)
            assert np.all((ids >= lo) & (ids < hi))
            np.testing.assert_allclose(Sm_c[b, ids], got_v[b, :kk][finite[b]], rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("B,N,d", [(512, 12101, 384), (128, 256, 64), (77, 1000, 64), (300, 5000, 768), (130, 890, 384), (64, 300, 24),
                                   (1, 513, 200), (256, 3000, 448)])
def test_tc_grad_matches_torch(dev, B, N, d):
    """O = softmax(S) E with S recomputed on the tensor cores and P rounded to bf16 before the second GEMM."""
    from hvae_b200 import _cabi
    lib = _cabi.lib()
    st = torch.cuda.current_stream().cuda_stream
    U, E = _operands(B, N, d, dev, seed=2)
    Ub, ldu = _cast(lib, U.contiguous(), dev)
    Eb, lde = _cast(lib, E.contiguous(), dev)
    ns = int(lib.tc_n_splits(B, N))
    ws = torch.empty(2 * B * ns, device=dev)
    lse = torch.empty(B, device=dev)
    lib.tc_score_lse(Ub.data_ptr(), ldu, B, Eb.data_ptr(), lde, N, d, lse.data_ptr(), ws.data_ptr(), st)
    gs = int(lib.tc_grad_splits(B, N, d))
    ldo = (d + 3) // 4 * 4
    Op = torch.full((gs, B, ldo), float("nan"), device=dev)
    lib.tc_score_grad(Ub.data_ptr(), ldu, B, Eb.data_ptr(), lde, N, d, lse.data_ptr(), Op.data_ptr(), ldo, st)
    O = Op.sum(0)[:, :d].double().cpu()
    S = Ub[:, :d].float().double() @ Eb[:, :d].float().double().t()
    Pm = torch.softmax(S, dim=1)
    ref = (Pm @ Eb[:, :d].float().double()).cpu()
    ref_b = (Pm.float().to(torch.bfloat16).double() @ Eb[:, :d].float().double()).cpu()
    scale = float(ref.abs().max())
    assert torch.isfinite(O).all()
    assert float((O - ref_b).abs().max()) < 2e-3 * scale + 1e-6      # same rounding point: only exp ulps / order differ
    assert float((O - ref).abs().max()) < 1e-2 * scale + 1e-6        # against exact softmax: bf16 rounding of P


def _onepass(lib, dev, U, E, st):
    """hvae_tc_score_onepass on the bf16-rounded operands -> (lse, O [B, d] combined as hvae_du_finalize does, c_part, S, E)."""
    B, d = U.shape
    N = E.shape[0]
    Eb, lde = _cast(lib, E.contiguous(), dev)
    Ub, ldu = _cast(lib, U.contiguous(), dev)
    S = Ub[:, :d].float().double() @ Eb[:, :d].float().double().t()
    gs = int(lib.tc_grad_splits(B, N, d))
    n_sub = int(lib.tc_onepass_subparts(d))
    ldo = (d + 3) // 4 * 4
    Op = torch.full((gs, B, ldo), float("nan"), device=dev)
    c_part = torch.full((gs, B), float("nan"), device=dev)
    l_part = torch.full((gs, n_sub, B), float("nan"), device=dev)
    w_part = torch.full((gs, B), float("nan"), device=dev)
    lse = torch.full((B,), float("nan"), device=dev)
    lib.tc_score_onepass(Ub.data_ptr(), ldu, B, Eb.data_ptr(), lde, N, d, c_part.data_ptr(), l_part.data_ptr(), Op.data_ptr(), ldo, st)
    lib.tc_onepass_combine(c_part.data_ptr(), l_part.data_ptr(), gs, n_sub, B, lse.data_ptr(), w_part.data_ptr(), st)
    # O = sum_p w_p O_p through the C ABI itself: hvae_du_finalize with an empty interaction matrix, oscale = 1 and 1/Bg = 1
    ip = torch.zeros(B + 1, dtype=torch.int64, device=dev)
    ix = torch.zeros(1, dtype=torch.int32, device=dev)
    one = torch.ones(B, device=dev)
    inv_bg = torch.ones(1, device=dev)
    dU = torch.empty(B, ldo, device=dev)
    lib.du_finalize(ip.data_ptr(), ix.data_ptr(), None, None, B, Op.data_ptr(), ldo, gs, one.data_ptr(), w_part.data_ptr(), Eb.data_ptr(), lde,
                    d, 1, inv_bg.data_ptr(), dU.data_ptr(), ldo, st)
    return lse, dU[:, :d].double(), c_part, S, Eb[:, :d].float().double()


@pytest.mark.parametrize("B,N,d", [(512, 12101, 384), (128, 256, 64), (77, 1000, 64), (300, 5000, 768), (130, 890, 384), (64, 300, 24),
                                   (1, 513, 200), (256, 3000, 448), (200, 1100, 1000), (5, 3, 16)])
def test_tc_onepass_matches_torch(dev, B, N, d):
    """Forward + backward through the scores in one sweep: lse and O = softmax(S) E from numerators against a fixed shift."""
    from hvae_b200 import _cabi
    lib = _cabi.lib()
    st = torch.cuda.current_stream().cuda_stream
    U, E = _operands(B, N, d, dev, seed=2)
    lse, O, c_part, S, Ed = _onepass(lib, dev, U, E, st)
    assert bool((c_part == 0).all())                                           # no sweep had to be repeated
    np.testing.assert_allclose(lse.cpu().numpy(), torch.logsumexp(S, dim=1).cpu().numpy(), rtol=2e-6, atol=2e-5)
    ref = torch.softmax(S, dim=1) @ Ed
    scale = float(ref.abs().max())
    assert torch.isfinite(O).all()
    assert float((O - ref).abs().max()) < 1e-2 * scale + 1e-6                 # bf16 rounding of the numerators


def test_tc_onepass_retry(dev):
    """Rows whose scores leave the window of the initial shift (largest score beyond ~ +70, or everything below ~ -35): the
    sweep is repeated with a moved shift -- only by the splits that saw the problem, so the per-split shifts differ and the
    combination has to weight them; peaked, flat and all-zero rows in the same batch."""
    from hvae_b200 import _cabi
    lib = _cabi.lib()
    st = torch.cuda.current_stream().cuda_stream
    for d, N in ((384, 3001), (768, 2500), (64, 40000)):
        B = 300                                                               # three user tiles
        U, E = _operands(B, N, d, dev, seed=5)
        v = torch.nn.functional.normalize(torch.ones(d, device=dev), dim=0)
        E = torch.nn.functional.normalize(E + 3.0 * v, dim=1)                 # every item shares a direction: cos(E_i, v) ~ 0.95
        U = U / U.norm(dim=1, keepdim=True) * 6.0
        U[3] = 48.0 * E[17]                                                   # large scores, but inside the first window
        U[130] = 150.0 * E[N - 5]                                             # S_max = 150 at the end of the catalogue: overflow
        U[7] = 500.0 * E[33]                                                  # many repeats
        U[131] = -100.0 * v                                                   # every score ~ -95: everything would be flushed
        U[260] = 0.0                                                          # uniform softmax
        lse, O, c_part, S, Ed = _onepass(lib, dev, U, E, st)
        moved = (c_part != 0).any(dim=0).cpu()
        assert bool(moved[130]) and bool(moved[7]) and bool(moved[131])
        assert not bool(moved[260]) and not bool(moved[200]) and not bool(moved[3])
        np.testing.assert_allclose(lse.cpu().numpy(), torch.logsumexp(S, dim=1).cpu().numpy(), rtol=3e-6, atol=5e-5)
        ref = torch.softmax(S, dim=1) @ Ed
        assert torch.isfinite(O).all()
        assert float((O - ref).abs().max()) < 1e-2 * float(ref.abs().max()) + 1e-6


def test_gemm_tf32_all_layouts(dev):
    """tcgen05 kind::tf32 GEMM: every transpose combination the MLP forward/backward uses, ragged sizes, bias, alpha."""
    from hvae_b200 import _cabi
    lib = _cabi.lib()
    g = torch.Generator().manual_seed(1)
    st = torch.cuda.current_stream().cuda_stream
    r4 = lambda n: (n + 3) // 4 * 4
    for (M, N, K) in [(512, 400, 600), (130, 70, 33), (512, 384, 200), (400, 600, 512), (5, 8, 7), (257, 768, 200), (37, 22, 64)]:
        A, B = torch.randn(M, K, generator=g), torch.randn(K, N, generator=g)
        bias = torch.randn(N, generator=g)
        ref = (2.0 * (A.double() @ B.double()) + bias.double()).numpy()
        scale = np.abs(ref).max()
        for a_t in (False, True):
            for b_t in (False, True):
                # pad the leading dimensions to multiples of 4 floats as the engine's buffers are
                if a_t:
                    Ad = torch.zeros(K, r4(M)); Ad[:, :M] = A.t(); a_rs, a_cs = 1, r4(M)
                else:
                    Ad = torch.zeros(M, r4(K)); Ad[:, :K] = A; a_rs, a_cs = r4(K), 1
                if b_t:
                    Bd = torch.zeros(N, r4(K)); Bd[:, :K] = B.t(); b_rs, b_cs = 1, r4(K)
                else:
                    Bd = torch.zeros(K, r4(N)); Bd[:, :N] = B; b_rs, b_cs = r4(N), 1
                Ad, Bd = Ad.to(dev), Bd.to(dev)
                assert lib.gemm_tf32_supported(Ad.data_ptr(), a_rs, a_cs, Bd.data_ptr(), b_rs, b_cs) == 1
                ldc = r4(N)
                C = torch.full((M, ldc), float("nan"), device=dev)
                lib.gemm_tf32(M, N, K, Ad.data_ptr(), a_rs, a_cs, Bd.data_ptr(), b_rs, b_cs, C.data_ptr(), ldc, bias.to(dev).data_ptr(), 2.0, st)
                got = C[:, :N].cpu().numpy()
                assert np.isfinite(got).all(), (M, N, K, a_t, b_t)
                assert np.abs(got - ref).max() < 2e-3 * scale, (M, N, K, a_t, b_t, np.abs(got - ref).max(), scale)   # TF32: 10-bit mantissa
                assert torch.isnan(C[:, N:]).all()        # pad columns untouched


def test_torch_custom_ops_match_plain_torch(dev):
    """torch.ops.hvae_b200.*: the registered operators give the library's results."""
    import hvae_b200.ops  # noqa: F401
    from hvae_b200.synth import make_interactions
    B, N, d, h = 200, 3000, 384, 600
    U, E = _operands(B, N, d, dev, seed=5)
    Ub, Eb = U.to(torch.bfloat16).contiguous(), E.to(torch.bfloat16).contiguous()
    S = Ub.float().double() @ Eb.float().double().t()
    lse = torch.ops.hvae_b200.score_lse(Ub, Eb, d)
    np.testing.assert_allclose(lse.cpu().numpy(), torch.logsumexp(S, 1).cpu().numpy(), rtol=2e-6, atol=2e-5)
    O = torch.ops.hvae_b200.score_grad(Ub, Eb, lse, d)
    ref = torch.softmax(S, 1) @ Eb.float().double()
    assert float((O.double() - ref).abs().max()) < 1e-2 * float(ref.abs().max())
    data = make_interactions(B, N, 1)
    indptr, indices = torch.from_numpy(data.indptr).to(dev), torch.from_numpy(data.indices).to(dev)
    rows = torch.arange(B, dtype=torch.int32, device=dev)
    val, idx = torch.ops.hvae_b200.score_topk(Ub, Eb, d, indptr, indices, rows, 10, 0)
    Sm = S.float().clone()
    Sm[torch.from_numpy(np.repeat(np.arange(B), np.diff(data.indptr))).to(dev), indices.long()] = -float("inf")
    np.testing.assert_allclose(val.cpu().numpy(), torch.topk(Sm, 10, dim=1)[0].cpu().numpy(), rtol=1e-5, atol=1e-5)
    # encoder layer 1 and an MLP GEMM
    g = torch.Generator().manual_seed(0)
    W1T = torch.randn(N, h, generator=g).to(dev)
    b1, gam, bet = torch.randn(h, generator=g).to(dev), torch.rand(h, generator=g).to(dev) + 0.5, torch.randn(h, generator=g).to(dev)
    act = torch.ops.hvae_b200.gather_ln_fwd(indptr, indices, rows, W1T, b1, gam, bet, h)
    from scipy.sparse import csr_matrix
    X = torch.from_numpy(csr_matrix((np.ones(len(data.indices), np.float32), data.indices, data.indptr), shape=(B, N)).toarray()).to(dev)
    ref_act = torch.nn.functional.gelu(torch.nn.functional.layer_norm(X @ W1T + b1, (h,), gam, bet, 1e-5))
    np.testing.assert_allclose(act.cpu().numpy(), ref_act.cpu().numpy(), rtol=2e-4, atol=2e-4)
    Wm = torch.randn(400, h, generator=g).to(dev)
    bm = torch.randn(400, generator=g).to(dev)
    out = torch.ops.hvae_b200.gemm(act, Wm, bm, True)
    ref_out = act.double() @ Wm.double().t() + bm.double()
    assert float((out.double() - ref_out).abs().max()) < 2e-3 * float(ref_out.abs().max())
    # Adam on a flat arena: one step equals torch.optim.Adam
    from hvae_b200._cabi import STATE_WORDS, p
    from hvae_b200 import _cabi
    prm = torch.randn(1024, generator=g).to(dev)
    grad = torch.randn(1024, generator=g).to(dev)
    ref_p = prm.clone().requires_grad_(True)
    opt = torch.optim.Adam([ref_p], lr=1e-3)
    ref_p.grad = grad.clone()
    opt.step()
    m, v = torch.zeros_like(prm), torch.zeros_like(prm)
    state = torch.zeros(STATE_WORDS, device=dev)
    lib = _cabi.lib()
    st = torch.cuda.current_stream().cuda_stream
    lib.step_begin(p(state), 1e-3, 0.9, 0.999, 0.0, 0.2, 0, 1, 1, 0, st)
    lib.grad_norm_clip(p(grad), 1024, None, None, 1e30, p(state), p(torch.empty(256, device=dev)), st)
    torch.ops.hvae_b200.adam_step(prm, m, v, grad, state, 0.0, 0.9, 0.999, 1e-8)
    np.testing.assert_allclose(prm.cpu().numpy(), ref_p.detach().cpu().numpy(), rtol=1e-6, atol=1e-7)
