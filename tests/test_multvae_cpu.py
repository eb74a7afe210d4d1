"""CPU: the Mult-VAE restatement (oracle/hvae_oracle.py: OracleMultVAE) against vectors frozen from the reference's own
MultVAE (oracle/make_golden_multvae.py; src/ml/baseline.py:126-206)."""
import numpy as np
import pytest
import torch
from scipy.sparse import csr_matrix

from golden_util import GOLDEN
from oracle import hvae_oracle as orc


def load_case():
    g = np.load(GOLDEN / "multvae.npz")
    csr = csr_matrix((g["values"].astype(np.float64), g["indices"], g["indptr"]), shape=(int(g["n_users"]), int(g["n_items"])))
    return g, csr


def dense_noise(g, csr, s):
    """The stored per-entry keep flags back on the dense [B, N] grid (entries that are zero in x do not matter)."""
    rows = g[f"rows/{s}"]
    sub = csr[rows]
    m = np.ones((len(rows), csr.shape[1]), dtype=np.float32)
    m[np.repeat(np.arange(len(rows)), np.diff(sub.indptr)), sub.indices] = g[f"noise/{s}/keep"]
    return dict(in_mask=torch.from_numpy(m), masks=[torch.from_numpy(g[f"noise/{s}/mask{i}"]).float() for i in (0, 1)],
                eps=torch.from_numpy(g[f"noise/{s}/eps"]))


@pytest.fixture(autouse=True)
def _one_thread():
    n = torch.get_num_threads()
    torch.set_num_threads(1)          # the golden was frozen with one thread (fixed reduction order)
    yield
    torch.set_num_threads(n)


def test_multvae_oracle_matches_reference():
    g, csr = load_case()
    n_items, h, L = int(g["n_items"]), int(g["hidden"]), int(g["latent"])
    torch.manual_seed(int(g["seed"]))
    m = orc.OracleMultVAE(n_items, h, L, float(g["dropout"]))
    for k, v in m.state_dict().items():
        assert np.array_equal(v.numpy(), g[f"init/{k}"]), k                       # same default initialisation, same draws
    opt = torch.optim.Adam(m.parameters(), lr=1e-3)
    for s in range(int(g["steps"])):
        x = torch.from_numpy(np.asarray(csr[g[f"rows/{s}"]].toarray(), dtype=np.float32))
        out = orc.multvae_train_step(m, opt, x, dense_noise(g, csr, s), float(g["beta"]))
        np.testing.assert_allclose(out, g["stats"][s], rtol=2e-6)
    for k, v in m.state_dict().items():
        np.testing.assert_allclose(v.numpy(), g[f"final/{k}"], rtol=1e-5, atol=1e-6, err_msg=k)
    m.eval()
    with torch.no_grad():
        s6, mu6, _ = m.forward_with(torch.from_numpy(np.asarray(csr[:6].toarray(), dtype=np.float32)), None)
    np.testing.assert_allclose(s6.numpy(), g["pred6/scores"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(mu6.numpy(), g["pred6/mu"], rtol=1e-5, atol=1e-6)
