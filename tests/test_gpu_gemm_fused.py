"""GPU: the TF32 GEMMs with a fused epilogue (GELU + dropout, bf16 copy, GELU backward, latent backward, bias-gradient column
sums) against the same GEMM followed by the stand-alone element-wise kernels they replace (src/ml/model.py:90-95,157-179 and
their autograd).  The GEMM part is the same kernel with the same tiling; the epilogues use a branch-free erf (absolute error
< 1e-7) where the stand-alone kernels call erff, and add the column sums in another fixed order, hence float32 tolerances."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _r4(n):
    return (n + 3) // 4 * 4


def _r8(n):
    return (n + 7) // 8 * 8


@pytest.fixture(scope="module")
def env():
    assert torch.cuda.is_available()
    from hvae_b200 import _cabi
    return _cabi.lib(), torch.device("cuda:0")


def _rand(shape, dev, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(dev)


def _padded(x, ld):
    out = torch.zeros(x.shape[0], ld, device=x.device, dtype=x.dtype)
    out[:, :x.shape[1]] = x
    return out


# (rows, out columns, k): the C2 / C3 shapes of the projection MLP and ragged ones (pad columns, partial tiles, split-K on and off)
SHAPES = [(512, 384, 200), (4096, 768, 200), (300, 50, 37), (130, 200, 768), (1, 64, 64), (700, 131, 45)]


@pytest.mark.parametrize("M,N,K", SHAPES)
@pytest.mark.parametrize("with_mask", [True, False])
def test_gelu_drop_epilogue(env, M, N, K, with_mask):
    lib, dev = env
    st = torch.cuda.current_stream().cuda_stream
    ldk, ldn = _r4(K), _r4(N)
    A = _padded(_rand((M, K), dev, 1), ldk)
    W = _padded(_rand((N, K), dev, 2, K ** -0.5), ldk)           # [out, in], as the model stores it
    bias = _rand((N,), dev, 3, 0.1)
    mask = (torch.rand(M, N, generator=torch.Generator().manual_seed(4)) > 0.3).to(torch.uint8).to(dev) if with_mask else None
    ks = 1.0 / 0.7
    pre_ref, act_ref = torch.full((M, ldn), 7.0, device=dev), torch.full((M, ldn), 7.0, device=dev)
    lib.gemm_tf32(M, N, K, A.data_ptr(), ldk, 1, W.data_ptr(), 1, ldk, pre_ref.data_ptr(), ldn, bias.data_ptr(), 1.0, st)
    lib.gelu_drop_fwd(pre_ref.data_ptr(), None if mask is None else mask.data_ptr(), ks, M, N, ldn, act_ref.data_ptr(), st)
    pre, act = torch.full((M, ldn), 9.0, device=dev), torch.full((M, ldn), 9.0, device=dev)
    lib.gemm_tf32_gelu_drop(M, N, K, A.data_ptr(), ldk, 1, W.data_ptr(), 1, ldk, pre.data_ptr(), act.data_ptr(), ldn, bias.data_ptr(),
                            None if mask is None else mask.data_ptr(), ks, st)
    torch.cuda.synchronize()
    assert torch.equal(pre[:, :N], pre_ref[:, :N])
    # the epilogue's branch-free erf: |error| < 1e-7 absolute (gemm_tc.cu erf_fast) -> GELU within |x| * 1e-7 of the erff kernel
    np.testing.assert_allclose(act[:, :N].cpu().numpy(), act_ref[:, :N].cpu().numpy(), rtol=1e-6, atol=1e-6)
    assert torch.all(pre[:, N:] == 0) and torch.all(act[:, N:] == 0)        # pad columns are written as zeros
    # against fp64 on TF32-free arithmetic: the TF32 operand rounding bounds the error
    ref = torch.nn.functional.gelu(A[:, :K].double() @ W[:, :K].double().t() + bias.double())
    if mask is not None:
        ref = ref * mask.double() * ks
    assert (act[:, :N].double() - ref).abs().max().item() < 2e-2


@pytest.mark.parametrize("M,N,K", SHAPES)
def test_bf16_epilogue(env, M, N, K):
    lib, dev = env
    st = torch.cuda.current_stream().cuda_stream
    ldk, ldn, ld8 = _r4(K), _r4(N), _r8(N)
    A = _padded(_rand((M, K), dev, 5), ldk)
    W = _padded(_rand((N, K), dev, 6, K ** -0.5), ldk)
    bias = _rand((N,), dev, 7, 0.1)
    c_ref = torch.zeros(M, ldn, device=dev)
    lib.gemm_tf32(M, N, K, A.data_ptr(), ldk, 1, W.data_ptr(), 1, ldk, c_ref.data_ptr(), ldn, bias.data_ptr(), 1.0, st)
    b_ref = torch.full((M, ld8), 3.0, dtype=torch.bfloat16, device=dev)
    lib.cast_bf16(c_ref.data_ptr(), M, N, ldn, b_ref.data_ptr(), ld8, st)
    for keep_fp32 in (True, False):
        c = torch.full((M, ldn), 9.0, device=dev)
        cb = torch.full((M, ld8), 5.0, dtype=torch.bfloat16, device=dev)
        lib.gemm_tf32_bf16(M, N, K, A.data_ptr(), ldk, 1, W.data_ptr(), 1, ldk, c.data_ptr() if keep_fp32 else None, ldn, bias.data_ptr(),
                           cb.data_ptr(), ld8, st)
        torch.cuda.synchronize()
        assert torch.equal(cb, b_ref)
        if keep_fp32:
            assert torch.equal(c[:, :N], c_ref[:, :N])
        else:
            assert torch.all(c == 9.0)


@pytest.mark.parametrize("M,N,K", SHAPES)
@pytest.mark.parametrize("with_mask", [True, False])
def test_gelu_bwd_epilogue_and_bias_gradient(env, M, N, K, with_mask):
    lib, dev = env
    st = torch.cuda.current_stream().cuda_stream
    ldk, ldn = _r4(K), _r4(N)
    dY = _padded(_rand((M, K), dev, 8), ldk)                       # upstream gradient [M, K]
    W = _padded(_rand((K, N), dev, 9, K ** -0.5), ldn)             # [out = K, in = N]: dX = dY W, n contiguous
    q = _padded(_rand((M, N), dev, 10), ldn)
    mask = (torch.rand(M, N, generator=torch.Generator().manual_seed(11)) > 0.3).to(torch.uint8).to(dev) if with_mask else None
    ks = 1.0 / 0.7
    mp = None if mask is None else mask.data_ptr()
    dt = torch.zeros(M, ldn, device=dev)
    lib.gemm_tf32(M, N, K, dY.data_ptr(), ldk, 1, W.data_ptr(), ldn, 1, dt.data_ptr(), ldn, None, 1.0, st)
    dq_ref = torch.zeros(M, ldn, device=dev)
    lib.gelu_drop_bwd(dt.data_ptr(), q.data_ptr(), mp, ks, M, N, ldn, dq_ref.data_ptr(), st)
    cs_ref, cs_tmp = torch.zeros(N, device=dev), torch.zeros(64 * N + 64, device=dev)
    lib.colsum(dq_ref.data_ptr(), ldn, M, N, cs_ref.data_ptr(), cs_tmp.data_ptr(), st)
    wsz = int(lib.gemm_colsum_workspace_floats(M, N))
    work = torch.zeros(wsz, device=dev)
    for rep in range(3):                                            # the counters must come back to zero: repeat on the same workspace
        dq, cs = torch.full((M, ldn), 9.0, device=dev), torch.full((N,), 9.0, device=dev)
        lib.gemm_tf32_gelu_bwd(M, N, K, dY.data_ptr(), ldk, 1, W.data_ptr(), ldn, 1, dq.data_ptr(), ldn, q.data_ptr(), mp, ks,
                               cs.data_ptr(), work.data_ptr(), st)
        torch.cuda.synchronize()
        np.testing.assert_allclose(dq[:, :N].cpu().numpy(), dq_ref[:, :N].cpu().numpy(), rtol=1e-6, atol=3e-6)
        assert torch.all(dq[:, N:] == 0)
        ref64 = dq[:, :N].double().sum(0)
        scale = dq[:, :N].double().abs().sum(0).clamp_min(1e-30)
        assert ((cs.double() - ref64).abs() / scale).max().item() < 1e-6
        ref64_ref = dq_ref[:, :N].double().sum(0)
        assert ((cs_ref.double() - ref64_ref).abs() / dq_ref[:, :N].double().abs().sum(0).clamp_min(1e-30)).max().item() < 1e-6
    # without the column sums
    dq = torch.zeros(M, ldn, device=dev)
    lib.gemm_tf32_gelu_bwd(M, N, K, dY.data_ptr(), ldk, 1, W.data_ptr(), ldn, 1, dq.data_ptr(), ldn, q.data_ptr(), mp, ks, None, None, st)
    torch.cuda.synchronize()
    np.testing.assert_allclose(dq[:, :N].cpu().numpy(), dq_ref[:, :N].cpu().numpy(), rtol=1e-6, atol=3e-6)


@pytest.mark.parametrize("M,L,K", [(512, 200, 384), (4096, 200, 768), (300, 37, 50), (130, 64, 768), (1, 32, 64), (700, 45, 131)])
@pytest.mark.parametrize("with_eps", [True, False])
def test_latent_bwd_epilogue_and_bias_gradient(env, M, L, K, with_eps):
    lib, dev = env
    st = torch.cuda.current_stream().cuda_stream
    ldk, ldl, ldml = _r4(K), _r4(L), _r4(2 * L)
    dq = _padded(_rand((M, K), dev, 12), ldk)
    W0 = _padded(_rand((K, L), dev, 13, K ** -0.5), ldl)            # projection_layer.0.weight [d, L]: dz = dq W0
    ml = _padded(_rand((M, 2 * L), dev, 14, 0.5), ldml)
    eps = _rand((M, L), dev, 15) if with_eps else None
    coef = torch.tensor([0.2 / M], device=dev)
    ep = None if eps is None else eps.data_ptr()
    dz = torch.zeros(M, ldl, device=dev)
    lib.gemm_tf32(M, L, K, dq.data_ptr(), ldk, 1, W0.data_ptr(), ldl, 1, dz.data_ptr(), ldl, None, 1.0, st)
    dml_ref = torch.zeros(M, ldml, device=dev)
    lib.latent_bwd(dz.data_ptr(), ldl, ml.data_ptr(), ldml, ep, M, L, coef.data_ptr(), dml_ref.data_ptr(), st)
    work = torch.zeros(int(lib.gemm_colsum_workspace_floats(M, 2 * L)), device=dev)
    for rep in range(2):
        dml, cs = torch.full((M, ldml), 9.0, device=dev), torch.full((2 * L,), 9.0, device=dev)
        lib.gemm_tf32_latent_bwd(M, L, K, dq.data_ptr(), ldk, 1, W0.data_ptr(), ldl, 1, ml.data_ptr(), ldml, ep, coef.data_ptr(),
                                 dml.data_ptr(), cs.data_ptr(), work.data_ptr(), st)
        torch.cuda.synchronize()
        np.testing.assert_allclose(dml[:, :2 * L].cpu().numpy(), dml_ref[:, :2 * L].cpu().numpy(), rtol=2e-6, atol=1e-7)
        ref64 = dml[:, :2 * L].double().sum(0)
        scale = dml[:, :2 * L].double().abs().sum(0).clamp_min(1e-30)
        assert ((cs.double() - ref64).abs() / scale).max().item() < 1e-6


@pytest.mark.parametrize("name", ["c2_train", "d768_train"])
def test_train_steps_same_with_and_without_fused_epilogues(env, name):
    """The benchmarked shapes' training steps (bf16 mode, the reference's noise) with the fused epilogues against the stand-alone
    kernels: same losses, same gradients up to summation order."""
    lib, dev = env
    from golden_util import BigCase
    from hvae_b200.engine import DeviceCSR
    from hvae_b200.model import create_hybrid_vae
    from hvae_b200.train import VAETrainer
    c = BigCase(name)

    def run(fused):
        torch.manual_seed(c.seed)
        m = create_hybrid_vae(**c.model_kwargs(), precision="bf16").to(dev)
        m.engine.fuse = 63 if fused else 0
        tr = VAETrainer(m, dev, lr=1e-3)
        csr = DeviceCSR.from_scipy(c.csr, dev)
        m.train()
        rows = c.rows(0)
        n, u8 = c.noise(0), (lambda t: None if t is None else t.to(torch.uint8).to(dev).contiguous())
        noise = dict(masks=[u8(t) for t in n["masks"]], eps=n["eps"].to(dev).contiguous(), pmask=u8(n["pmask"]))
        tr.train_step(csr.batch(torch.tensor(rows, dtype=torch.int32, device=dev), rows), noise)
        losses = tr.last_losses()
        st = m.engine.read_state()
        return np.array(losses, dtype=np.float64), st["grad_norm"], m.engine.gd.clone()

    l1, g1, d1 = run(True)
    l0, g0, d0 = run(False)
    np.testing.assert_allclose(l1, l0, rtol=1e-6)
    np.testing.assert_allclose(g1, g0, rtol=1e-5)
    # a 1e-7 change in an activation can flip the TF32 rounding (2^-11) of an operand of the GEMMs downstream
    scale = d0.abs().max().item()
    assert (d1 - d0).abs().max().item() <= 1e-4 * scale


@pytest.mark.parametrize("M,N,K", [(8, 768, 200), (64, 600, 200), (512, 384, 600), (4096, 768, 768), (300, 50, 37)])
def test_gemm_tf32_same_bits_every_launch(env, M, N, K):
    """Fixed-order reductions everywhere (split-K partials in rank order, column sums in tile order): repeated launches -- with the
    cache contents churned in between -- return the same bits, for the plain GEMM in the three operand layouts of a step and for the
    fused epilogues."""
    lib, dev = env
    st = torch.cuda.current_stream().cuda_stream
    ldk, ldn = _r4(K), _r4(N)
    A, W, bias = _padded(_rand((M, K), dev, 21), ldk), _padded(_rand((N, K), dev, 22, K ** -0.5), ldk), _rand((N,), dev, 23)
    Wt = _padded(_rand((K, N), dev, 24, K ** -0.5), ldn)
    G = _padded(_rand((M, N), dev, 25), ldn)
    q = _padded(_rand((M, N), dev, 26), ldn)
    mask = (torch.rand(M, N, generator=torch.Generator().manual_seed(27)) > 0.3).to(torch.uint8).to(dev)
    work = torch.zeros(int(lib.gemm_colsum_workspace_floats(M, N)), device=dev)
    churn = torch.empty(64 << 20, dtype=torch.uint8, device=dev)

    def run():
        c1, c2, c3 = torch.zeros(M, ldn, device=dev), torch.zeros(M, ldn, device=dev), torch.zeros(N, ldk, device=dev)
        c4, cs = torch.zeros(M, ldn, device=dev), torch.zeros(N, device=dev)
        lib.gemm_tf32(M, N, K, A.data_ptr(), ldk, 1, W.data_ptr(), 1, ldk, c1.data_ptr(), ldn, bias.data_ptr(), 1.0, st)         # y = x W^T + b
        lib.gemm_tf32(M, N, K, A.data_ptr(), ldk, 1, Wt.data_ptr(), ldn, 1, c2.data_ptr(), ldn, None, 1.0, st)                   # dx = dy W
        lib.gemm_tf32(N, K, M, G.data_ptr(), 1, ldn, A.data_ptr(), ldk, 1, c3.data_ptr(), ldk, None, 1.0, st)                    # dW = dy^T x
        lib.gemm_tf32_gelu_bwd(M, N, K, A.data_ptr(), ldk, 1, Wt.data_ptr(), ldn, 1, c4.data_ptr(), ldn, q.data_ptr(), mask.data_ptr(),
                               1 / 0.7, cs.data_ptr(), work.data_ptr(), st)
        torch.cuda.synchronize()
        return c1, c2, c3, c4, cs

    first = run()
    for _ in range(15):
        churn.random_(0, 255)
        again = run()
        for a, b in zip(first, again):
            assert torch.equal(a, b)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_train_step_and_topk_same_bits_every_run(env, precision):
    """Two fresh runs of the same two training steps (the reference's noise) and of the evaluator's top-20: identical bits --
    no atomics or arrival-order reductions anywhere on the path."""
    lib, dev = env
    from golden_util import BigCase
    from hvae_b200.engine import DeviceCSR
    from hvae_b200.evaluate import RecommendationEvaluator
    from hvae_b200.model import create_hybrid_vae
    from hvae_b200.train import VAETrainer
    c = BigCase("d768_train")
    churn = torch.empty(64 << 20, dtype=torch.uint8, device=dev)

    def run():
        torch.manual_seed(c.seed)
        m = create_hybrid_vae(**c.model_kwargs(), precision=precision).to(dev)
        tr = VAETrainer(m, dev, lr=1e-3)
        csr = DeviceCSR.from_scipy(c.csr, dev)
        m.train()
        losses = []
        for s in range(c.steps):
            churn.random_(0, 255)
            rows = c.rows(s)
            n, u8 = c.noise(s), (lambda t: None if t is None else t.to(torch.uint8).to(dev).contiguous())
            noise = dict(masks=[u8(t) for t in n["masks"]], eps=n["eps"].to(dev).contiguous(), pmask=u8(n["pmask"]))
            tr.train_step(csr.batch(torch.tensor(rows, dtype=torch.int32, device=dev), rows), noise)
            losses.append(tr.last_losses())
        m.eval()
        ev = RecommendationEvaluator(m, c.csr, {}, {}, dev)
        val, idx = ev.topk_users(np.arange(64), 20)
        return losses, m.engine.arena.clone(), val.clone(), idx.clone()

    l0, a0, v0, i0 = run()
    for _ in range(3):
        l1, a1, v1, i1 = run()
        assert l0 == l1
        assert torch.equal(a0, a1) and torch.equal(v0, v1) and torch.equal(i0, i1)
