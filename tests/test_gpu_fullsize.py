"""GPU: the kernels at BASELINE.json's FULL shapes (C3: 4,096 users x 200,000 items x d = 768; C4: 1,000,000 items), where no
materialised reference fits the time budget of a test -- checked through size-independent properties instead:
  * two different kernels must agree (one-pass forward+backward vs the two-launch LSE kernel; item-sharded top-K + merge vs one sweep),
  * checksums (a constant embedding column turns O = softmax(S) E into the softmax's own normalisation: that column of O must be 1),
  * sortedness / uniqueness / no seen item / every returned score recomputed from its id,
  * linearity of the layer-1 gather-sum in the interaction row,
  * sampled rows against float64 torch on the same bf16-rounded operands (src/ml/model.py:198,281, src/ml/evaluate.py:125-147)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env():
    assert torch.cuda.is_available()
    from hvae_b200 import _cabi
    return _cabi.lib(), torch.device("cuda:0")


def _r8(n):
    return (n + 7) // 8 * 8


def _cast(lib, x, dev):
    rows, cols = x.shape
    ld = _r8(cols)
    out = torch.empty(rows, ld, dtype=torch.bfloat16, device=dev)
    lib.cast_bf16(x.data_ptr(), rows, cols, cols, out.data_ptr(), ld, torch.cuda.current_stream().cuda_stream)
    return out, ld


def _operands(B, N, d, dev, seed):
    g = torch.Generator(device=dev).manual_seed(seed)
    U = torch.randn(B, d, generator=g, device=dev) * 4.0          # scores ~ N(0, 16): a peaked softmax, maxima around 18
    E = torch.nn.functional.normalize(torch.randn(N, d, generator=g, device=dev), dim=1)
    return U, E


def test_c3_shape_onepass_scoring_properties(env):
    lib, dev = env
    st = torch.cuda.current_stream().cuda_stream
    B, N, d = 4096, 200_000, 768
    U, E = _operands(B, N, d, dev, 11)
    E[:, 0] = 1.0                               # checksum column: O[:, 0] = sum_i softmax_i = 1
    U[:, 0] = 0.0                               # (and it must not move the scores)
    Ub, ldu = _cast(lib, U, dev)
    Eb, lde = _cast(lib, E, dev)
    # two-launch path: exact log-sum-exp of the forward kernel
    ns = int(lib.tc_n_splits(B, N))
    ws = torch.empty(2 * B * ns, device=dev)
    lse2 = torch.empty(B, device=dev)
    lib.tc_score_lse(Ub.data_ptr(), ldu, B, Eb.data_ptr(), lde, N, d, lse2.data_ptr(), ws.data_ptr(), st)
    # one-pass kernel (what a training step runs) + combination
    gs, n_sub, ldo = int(lib.tc_grad_splits(B, N, d)), int(lib.tc_onepass_subparts(d)), d
    Op = torch.empty(gs, B, ldo, device=dev)
    c_part, l_part, w_part = torch.empty(gs, B, device=dev), torch.empty(gs, n_sub, B, device=dev), torch.empty(gs, B, device=dev)
    lse1 = torch.empty(B, device=dev)
    lib.tc_score_onepass(Ub.data_ptr(), ldu, B, Eb.data_ptr(), lde, N, d, c_part.data_ptr(), l_part.data_ptr(), Op.data_ptr(), ldo, st)
    lib.tc_onepass_combine(c_part.data_ptr(), l_part.data_ptr(), gs, n_sub, B, lse1.data_ptr(), w_part.data_ptr(), st)
    torch.cuda.synchronize()
    assert bool((c_part == 0).all())                                    # scores of this size stay inside the first window
    np.testing.assert_allclose(lse1.cpu().numpy(), lse2.cpu().numpy(), rtol=2e-6, atol=2e-5)      # two kernels, one answer
    O = (Op * w_part[:, :, None]).sum(0)
    assert torch.isfinite(O).all()
    # checksum: the numerators are rounded to bf16 (up to 2^-9 relative each, unbiased) before the second GEMM while the denominator adds
    # them unrounded -- a row that one or two items dominate shows their rounding errors in full, the average row far less
    chk = O[:, 0].cpu().numpy()
    np.testing.assert_allclose(chk, 1.0, rtol=0, atol=4e-3)
    assert np.abs(chk - 1.0).mean() < 5e-4 and abs(chk.mean() - 1.0) < 1e-4
    # the folded variant (combination inside du_finalize) gives the same lse and the same weighted sum
    ip, ix = torch.zeros(B + 1, dtype=torch.int64, device=dev), torch.zeros(1, dtype=torch.int32, device=dev)
    one, inv_bg = torch.ones(B, device=dev), torch.ones(1, device=dev)
    dU, lse3 = torch.empty(B, ldo, device=dev), torch.empty(B, device=dev)
    lib.du_finalize_onepass(ip.data_ptr(), ix.data_ptr(), None, None, B, Op.data_ptr(), ldo, gs, one.data_ptr(), c_part.data_ptr(),
                            l_part.data_ptr(), n_sub, lse3.data_ptr(), Eb.data_ptr(), lde, d, 1, inv_bg.data_ptr(), dU.data_ptr(), ldo, st)
    torch.cuda.synchronize()
    np.testing.assert_allclose(lse3.cpu().numpy(), lse1.cpu().numpy(), rtol=1e-6, atol=1e-6)
    assert float((dU - O).abs().max()) <= 1e-5 * float(O.abs().max())
    # sampled rows against float64 on the same operands
    rows = torch.tensor([0, 1, 63, 64, 127, 128, 2047, 2048, 4000, 4095], device=dev)
    S = Ub[rows][:, :d].float().double() @ Eb[:, :d].float().double().t()
    np.testing.assert_allclose(lse1[rows].cpu().numpy(), torch.logsumexp(S, dim=1).cpu().numpy(), rtol=2e-6, atol=2e-5)
    ref = torch.softmax(S, dim=1) @ Eb[:, :d].float().double()
    assert float((O[rows].double() - ref).abs().max()) < 1e-2 * float(ref.abs().max())


def _topk(lib, dev, Ub, ldu, Eb, lde, B, d, lo, hi, K, indptr, indices):
    st = torch.cuda.current_stream().cuda_stream
    n_it = hi - lo
    ns = int(lib.tc_topk_splits(B, n_it))
    cv, ci = torch.empty(B, ns * K, device=dev), torch.empty(B, ns * K, dtype=torch.int32, device=dev)
    lib.tc_score_topk(Ub.data_ptr(), ldu, B, Eb.data_ptr() + 2 * lo * lde, lde, n_it, d, lo, indptr.data_ptr(), indices.data_ptr(), None, K,
                      cv.data_ptr(), ci.data_ptr(), st)
    ov, oi = torch.empty(B, K, device=dev), torch.empty(B, K, dtype=torch.int32, device=dev)
    lib.topk_merge(cv.data_ptr(), ci.data_ptr(), B, ns * K, K, ov.data_ptr(), oi.data_ptr(), st)
    return ov, oi


def test_c4_shape_fused_topk_properties(env):
    lib, dev = env
    from hvae_b200.synth import make_interactions
    B, N, d, K = 2048, 1_000_000, 768, 20
    U, E = _operands(B, N, d, dev, 12)
    Ub, ldu = _cast(lib, U, dev)
    Eb, lde = _cast(lib, E, dev)
    del E
    data = make_interactions(B, N, 5)
    indptr, indices = torch.from_numpy(data.indptr).to(dev), torch.from_numpy(data.indices).to(dev)
    ov, oi = _topk(lib, dev, Ub, ldu, Eb, lde, B, d, 0, N, K, indptr, indices)
    torch.cuda.synchronize()
    v, i = ov.cpu().numpy(), oi.cpu().numpy().astype(np.int64)
    # sorted by (score desc, id desc), ids unique and inside the catalogue
    assert np.all(v[:, :-1] >= v[:, 1:])
    ties = v[:, :-1] == v[:, 1:]
    assert np.all(i[:, :-1][ties] > i[:, 1:][ties])
    assert np.all((i >= 0) & (i < N)) and all(len(set(r.tolist())) == K for r in i)
    # no seen item is recommended
    seen = set(zip(np.repeat(np.arange(B), np.diff(data.indptr)).tolist(), data.indices.tolist()))
    assert not any((b, int(it)) in seen for b in range(B) for it in i[b])
    # every returned score is the score of its id (fp32 accumulation of bf16 products: order only)
    sc = (Ub[:, None, :d].float() * Eb[oi.long().reshape(-1)].reshape(B, K, -1)[:, :, :d].float()).sum(-1)
    np.testing.assert_allclose(v, sc.cpu().numpy(), rtol=2e-5, atol=2e-5)
    # item-sharded (four shards, ragged boundaries) + merge == one sweep: identical ids and values
    cuts = [0, 250_003, 499_999, 777_777, N]
    pv, pi = [], []
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        a, b = _topk(lib, dev, Ub, ldu, Eb, lde, B, d, lo, hi, K, indptr, indices)
        pv.append(a); pi.append(b)
    cv, ci = torch.cat(pv, dim=1).contiguous(), torch.cat(pi, dim=1).contiguous()
    mv, mi = torch.empty(B, K, device=dev), torch.empty(B, K, dtype=torch.int32, device=dev)
    lib.topk_merge(cv.data_ptr(), ci.data_ptr(), B, 4 * K, K, mv.data_ptr(), mi.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert torch.equal(mi, oi) and torch.equal(mv, ov)
    # idempotence: the top-K of the top-K is itself
    m2v, m2i = torch.empty(B, K, device=dev), torch.empty(B, K, dtype=torch.int32, device=dev)
    lib.topk_merge(ov.data_ptr(), oi.data_ptr(), B, K, K, m2v.data_ptr(), m2i.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert torch.equal(m2i, oi) and torch.equal(m2v, ov)
    # sampled rows against torch.topk over all 1M scores of the same operands
    rows = [0, 1, 127, 128, 1000, 2047]
    S = Ub[rows][:, :d].float() @ Eb[:, :d].float().t()
    for k, b in enumerate(rows):
        S[k, indices[indptr[b]:indptr[b + 1]].long()] = -float("inf")
    rv, _ = torch.topk(S, K, dim=1)
    np.testing.assert_allclose(v[rows], rv.cpu().numpy(), rtol=2e-5, atol=2e-5)


def test_c4_shape_gather_sum_is_linear(env):
    """Layer 1 at 1M items (W1^T = 2.4 GB): pre(x_a + x_b) = pre(x_a) + pre(x_b) - bias for rows with disjoint items, and a row
    with a single item returns that item's weight row + bias bit for bit (src/ml/model.py:111-114 as a gather-sum)."""
    lib, dev = env
    st = torch.cuda.current_stream().cuda_stream
    N, h, B = 1_000_000, 600, 1024
    g = torch.Generator(device=dev).manual_seed(13)
    W1T = torch.randn(N, h, generator=g, device=dev) * 0.02
    bias = torch.randn(h, generator=g, device=dev) * 0.1
    rng = np.random.default_rng(14)
    na, nb = rng.integers(1, 40, B), rng.integers(1, 40, B)
    rows_a = [np.sort(rng.choice(N // 2, n, replace=False)) for n in na]                   # items of x_a in the lower half,
    rows_b = [N // 2 + np.sort(rng.choice(N // 2, n, replace=False)) for n in nb]          # of x_b in the upper: disjoint
    rows_ab = [np.concatenate([a, b]) for a, b in zip(rows_a, rows_b)]
    single = rng.integers(0, N, B)

    def run(lists, vals=None):
        indptr = torch.from_numpy(np.concatenate([[0], np.cumsum([len(x) for x in lists])]).astype(np.int64)).to(dev)
        idx = torch.from_numpy(np.concatenate(lists).astype(np.int32)).to(dev)
        val = torch.ones(idx.shape[0], device=dev) if vals is None else torch.from_numpy(np.concatenate(vals).astype(np.float32)).to(dev)
        pre = torch.empty(len(lists), h, device=dev)
        lib.gather_ln_fwd(indptr.data_ptr(), idx.data_ptr(), val.data_ptr(), None, len(lists), W1T.data_ptr(), h, h, bias.data_ptr(), None, None,
                          None, 1.0, pre.data_ptr(), None, None, pre.data_ptr(), st)
        torch.cuda.synchronize()
        return pre

    pa, pb, pab = run(rows_a), run(rows_b), run(rows_ab)
    scale = float(pab.abs().max())
    assert float((pab - (pa + pb - bias)).abs().max()) <= 2e-6 * scale + 1e-7
    ps = run([np.array([s]) for s in single])
    assert torch.equal(ps, W1T[torch.from_numpy(single).to(dev)] + bias)
    # homogeneity: interaction values 2x -> (pre - bias) 2x, exactly (powers of two)
    p2 = run(rows_a, [np.full(len(a), 2.0) for a in rows_a])
    assert torch.equal(p2 - bias, 2.0 * (pa - bias)) or float(((p2 - bias) - 2.0 * (pa - bias)).abs().max()) <= 1e-6 * scale


def test_c4_shape_fp32_topk_scan_properties(env):
    """hvae_mask_topk over 256 x 1M materialised fp32 scores (the exact mode's evaluator): equals torch.topk, ids bit for bit."""
    lib, dev = env
    st = torch.cuda.current_stream().cuda_stream
    R, N, K = 256, 1_000_000, 20
    g = torch.Generator(device=dev).manual_seed(15)
    S = torch.randn(R, N, generator=g, device=dev)
    S[:, ::7] = torch.round(S[:, ::7] * 8) / 8                      # plenty of exact ties
    zero_ptr = torch.zeros(R + 1, dtype=torch.int64, device=dev)
    nc = int(lib.mask_topk_chunks(R, N))
    cv, ci = torch.empty(R, nc * K, device=dev), torch.empty(R, nc * K, dtype=torch.int32, device=dev)
    val, idx = torch.empty(R, K, device=dev), torch.empty(R, K, dtype=torch.int32, device=dev)
    ref = S.clone()
    lib.mask_topk(S.data_ptr(), N, R, N, 0, zero_ptr.data_ptr(), None, None, 0, K, cv.data_ptr(), ci.data_ptr(), val.data_ptr(), idx.data_ptr(), st)
    torch.cuda.synchronize()
    rv, _ = torch.topk(ref, K, dim=1)
    assert torch.equal(val, rv)
    v, i = val.cpu().numpy(), idx.cpu().numpy().astype(np.int64)
    ties = v[:, :-1] == v[:, 1:]
    assert np.all(i[:, :-1][ties] > i[:, 1:][ties])                 # the evaluator's order: ties by index, descending
    assert torch.equal(ref.gather(1, idx.long()), val)
