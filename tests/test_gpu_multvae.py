"""GPU: the Mult-VAE baseline on the hvae_b200 kernels (hvae_b200/multvae.py) against vectors frozen from the reference's own
MultVAE (src/ml/baseline.py:126-231): state_dict layout, three optimisation steps with the recorded noise, predictions."""
import numpy as np
import pytest
import torch

from test_multvae_cpu import load_case

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda:0")


def test_multvae_matches_reference(dev):
    from hvae_b200.engine import DeviceCSR
    from hvae_b200.multvae import MultVAE
    g, csr = load_case()
    n_items, h, L = int(g["n_items"]), int(g["hidden"]), int(g["latent"])
    torch.manual_seed(int(g["seed"]))
    m = MultVAE(n_items, h, L, dropout=float(g["dropout"]))
    sd = m.state_dict()
    assert list(sd) == [k[5:] for k in g.files if k.startswith("init/")]           # the reference's keys, in its order
    for k, v in sd.items():
        assert v.shape == g[f"init/{k}"].shape and np.array_equal(v.numpy(), g[f"init/{k}"]), k       # same init under the same seed
    m = m.to(dev)
    rt = m.rt
    dcsr = DeviceCSR.from_scipy(csr, dev)
    m.train()
    for s in range(int(g["steps"])):
        rows = g[f"rows/{s}"]
        b = dcsr.batch(torch.tensor(rows, dtype=torch.int32, device=dev), rows)
        noise = dict(keep=torch.from_numpy(g[f"noise/{s}/keep"]).to(dev), masks=[torch.from_numpy(g[f"noise/{s}/mask{i}"]).to(dev).contiguous()
                                                                                 for i in (0, 1)],
                     eps=torch.from_numpy(g[f"noise/{s}/eps"]).to(dev).contiguous())
        rt.train_step(b, noise, 1e-3, float(g["beta"]))
        np.testing.assert_allclose(rt.last_losses(), g["stats"][s], rtol=1e-5, err_msg=f"step {s}")
    sd = m.state_dict()
    for k, v in sd.items():
        np.testing.assert_allclose(v.cpu().numpy(), g[f"final/{k}"], rtol=2e-4, atol=1e-5, err_msg=k)   # (Adam: lr-sized moves of near-zero weights)
    m.eval()
    x6 = torch.from_numpy(np.asarray(csr[:6].toarray(), dtype=np.float32)).to(dev)
    s6, mu6, lv6 = m(x6)
    np.testing.assert_allclose(s6.cpu().numpy(), g["pred6/scores"], rtol=2e-4, atol=2e-5)
    np.testing.assert_allclose(mu6.cpu().numpy(), g["pred6/mu"], rtol=2e-4, atol=2e-5)
    # a reference-layout state_dict loads back (strict), round trip is exact
    m2 = MultVAE(n_items, h, L, dropout=float(g["dropout"]))
    m2.load_state_dict({k: v.cpu() for k, v in sd.items()})
    for k, v in m2.state_dict().items():
        assert torch.equal(v, sd[k].cpu()), k


def test_multvae_recommender_fit_predict(dev):
    """MultVAERecommender.fit / predict (baseline.py:163-231): the loss goes down, predict() returns one score per item."""
    from hvae_b200.multvae import MultVAERecommender
    g, csr = load_case()
    torch.manual_seed(0)
    rec = MultVAERecommender(hidden_dim=32, latent_dim=8, epochs=3, lr=1e-2, beta=0.2, device=dev)
    rec.fit(csr)
    first = rec.model.rt.acc.cpu().numpy()
    s = rec.predict(3)
    assert s.shape == (csr.shape[1],) and np.all(np.isfinite(s))
    rec2 = MultVAERecommender(hidden_dim=32, latent_dim=8, epochs=30, lr=1e-2, beta=0.2, device=dev)
    rec2.fit(csr)
    last = rec2.model.rt.acc.cpu().numpy()
    assert last[0] / last[3] < first[0] / first[3]
