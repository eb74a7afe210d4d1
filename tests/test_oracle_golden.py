"""CPU: pin the oracle (oracle/hvae_oracle.py) against vectors frozen from the reference itself
(oracle/make_golden.py).  Same library, same op order, one thread -> equality is expected to be
exact; a tiny tolerance is allowed only where the op sequence differs (noise passed in as tensors)."""
import numpy as np
import pytest
import torch

from golden_util import TRAIN_CASES, Case
from oracle import hvae_oracle as orc


@pytest.fixture(autouse=True)
def _one_thread():
    n = torch.get_num_threads()
    torch.set_num_threads(1)
    yield
    torch.set_num_threads(n)


@pytest.mark.parametrize("name", TRAIN_CASES)
def test_init_matches_reference(name):
    c = Case(name)
    torch.manual_seed(c.seed)
    m = orc.OracleVAE(**c.model_kwargs())
    ref = c.state("init")
    sd = m.state_dict()
    assert list(sd.keys()) == list(ref.keys())
    for k in ref:
        assert torch.equal(sd[k], ref[k]), k


@pytest.mark.parametrize("name", TRAIN_CASES)
def test_train_steps_match_reference(name):
    c = Case(name)
    m = orc.OracleVAE(**c.model_kwargs())
    m.load_state_dict(c.state("init"))
    opt = orc.make_adam(m, 1e-3, c.weight_decay)
    stats = c.z["stats"]
    for s in range(c.steps):
        x = c.dense(c.rows(s))
        out = orc.train_step(m, opt, x, c.noise(s), c.beta_at(s))
        if s == 0:
            # reference grads were stored after clipping: ours are clipped in place too
            for k, p in m.named_parameters():
                np.testing.assert_allclose(p.grad.numpy(), c.z[f"grad0/{k}"], rtol=1e-5, atol=1e-7, err_msg=k)
        np.testing.assert_allclose(out, stats[s], rtol=2e-6, err_msg=f"step {s}")
    for k, v in c.state("final").items():
        np.testing.assert_allclose(m.state_dict()[k].numpy(), v.numpy(), rtol=1e-5, atol=1e-6, err_msg=k)


@pytest.mark.parametrize("name", TRAIN_CASES)
def test_noise_replay_is_what_the_reference_draws(name):
    """draw_noise() under the generator state the reference started step 0 from must give the
    recorded tensors: the recorded ones were themselves produced that way, and the reference's loss
    with its own RNG equals the oracle's loss with them (previous test)."""
    c = Case(name)
    torch.manual_seed(c.seed)
    m = orc.OracleVAE(**c.model_kwargs())            # consumes the init draws
    torch.manual_seed(c.seed)
    orc.OracleVAE(**c.model_kwargs())
    orc.OracleVAE(**c.model_kwargs())                # the generator's shadow model (make_golden.py)
    n = orc.draw_noise(m, len(c.rows(0)))
    g = c.noise(0)
    assert torch.equal(n["eps"], g["eps"])
    for a, b in zip(n["masks"], g["masks"]):
        assert torch.equal(a, b)


@pytest.mark.parametrize("name", TRAIN_CASES)
def test_validate_and_forward(name):
    c = Case(name)
    m = orc.OracleVAE(**c.model_kwargs())
    m.load_state_dict(c.state("final"))
    m.eval()
    with torch.no_grad():
        s, mu, lv = m.forward_with(c.dense(np.arange(8)), None)
    np.testing.assert_allclose(s.numpy(), c.z["fwd8/scores"], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(mu.numpy(), c.z["fwd8/mu"], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(lv.numpy(), c.z["fwd8/logvar"], rtol=1e-6, atol=1e-6)
    n = min(c.n_users, 2 * c.batch)
    vals = [orc.validate_batch(m, c.dense(np.arange(i, min(i + c.batch, n))), c.beta) for i in range(0, n, c.batch)]
    np.testing.assert_allclose(np.mean(vals, axis=0), c.z["validate"], rtol=2e-6)


@pytest.mark.parametrize("name,train", [("tiny_two_hidden_eval", "tiny_two_hidden"),
                                        ("tiny_identity_eval", "tiny_identity")])
def test_full_ranking_matches_reference(name, train):
    c = Case(train)
    g = np.load(c.z.fid.name.replace(train + ".npz", name + ".npz")) if False else np.load(
        str(__import__("golden_util").GOLDEN / f"{name}.npz"))
    m = orc.OracleVAE(**c.model_kwargs())
    m.load_state_dict(c.state("final"))
    ks = tuple(int(k) for k in g["k_values"])
    res, tops = orc.full_ranking_eval(m, c.csr, c.test_items, ks)
    assert np.array_equal(tops, g["topk"])
    got = np.array([[res[k][mm] for mm in ("recall", "ndcg", "hit_ratio")] for k in ks])
    np.testing.assert_allclose(got, g["metrics"], rtol=0, atol=1e-12)
    for u in range(6):
        idx, sc = orc.recommend(m, c.csr, u, top_k=100)
        assert np.array_equal(idx, g[f"rec100/{u}/idx"])
        np.testing.assert_array_equal(sc, g[f"rec100/{u}/score"])
        idx, sc = orc.recommend(m, c.csr, u, top_k=10, exclude_seen=False)
        assert np.array_equal(idx, g[f"rec10_all/{u}/idx"])


def test_c1_shape_eval_matches_reference():
    from hvae_b200.synth import CONFIGS, make_interactions, make_item_embeddings
    g = np.load(str(__import__("golden_util").GOLDEN / "c1_eval.npz"))
    c = CONFIGS["c1"]
    data = make_interactions(c["n_users"], c["n_items"], 0)
    E = make_item_embeddings(c["n_items"], c["emb_dim"], 0)
    torch.manual_seed(0)
    m = orc.OracleVAE(c["n_items"], E, c["latent_dim"], c["hidden_dims"], c["dropout"], c["beta"])
    for k, v in m.state_dict().items():
        assert float(v.double().sum()) == float(g[f"sum/{k}"]), k
    users = g["users"]
    res, tops = orc.full_ranking_eval(m, data.scipy_csr(), data.test_items, (5, 10, 20), users)
    assert np.array_equal(tops, g["topk"])
    got = np.array([[res[k][mm] for mm in ("recall", "ndcg", "hit_ratio")] for k in (5, 10, 20)])
    np.testing.assert_allclose(got, g["metrics"], atol=1e-12)
