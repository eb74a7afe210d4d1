"""CPU: the oracle against the reference-frozen goldens at the benchmarked shapes (oracle/make_golden_big.py):
C2 training steps, the d = 768 case, the reference's own randn unit-test input, evaluate_config_on_val."""
import numpy as np
import pytest
import torch

from golden_util import GOLDEN, BigCase
from oracle import hvae_oracle as orc


def _ref_named(m):
    return {k: v for k, v in m.state_dict().items() if k != "item_embeddings"}


@pytest.mark.parametrize("name", ["c2_train", "d768_train"])
def test_big_train_steps_match_reference(name):
    c = BigCase(name)
    torch.manual_seed(c.seed)
    m = orc.OracleVAE(**c.model_kwargs())
    c.check_digest("init", _ref_named(m), rtol=0.0, atol_scale=0.0)          # same init draws -> same bits
    opt = orc.make_adam(m, 1e-3, 0.0)
    for s in range(c.steps):
        rows = c.rows(s)
        x = torch.from_numpy(np.asarray(c.csr[rows].toarray(), dtype=np.float32))
        out = orc.train_step(m, opt, x, c.noise(s), c.beta)
        np.testing.assert_allclose(out, c.stats[s], rtol=3e-6, err_msg=f"{name} step {s}")
    c.check_digest("final", _ref_named(m), rtol=2e-5)
    osd = opt.state_dict()["state"]
    names = [k for k, _ in m.named_parameters()]
    c.check_digest("adam_m", {k: osd[i]["exp_avg"] for i, k in enumerate(names)}, rtol=2e-5, atol_scale=1e-9)
    c.check_digest("adam_v", {k: osd[i]["exp_avg_sq"] for i, k in enumerate(names)}, rtol=4e-5, atol_scale=1e-12)
    vb = len(c.rows(0))
    n = min(c.n_users, 2 * vb)
    m.eval()
    vals = [orc.validate_batch(m, torch.from_numpy(np.asarray(c.csr[i:min(i + vb, n)].toarray(), dtype=np.float32)), c.beta)
            for i in range(0, n, vb)]
    np.testing.assert_allclose(np.mean(vals, axis=0), c.validate, rtol=3e-6)
    if "fwd8/scores" in c.z.files:
        with torch.no_grad():
            s8, mu8, lv8 = m.forward_with(torch.from_numpy(np.asarray(c.csr[:8].toarray(), dtype=np.float32)), None)
        # two fp32 executions of two Adam steps (other thread count -> other summation order) agree in the mean but not entry by
        # entry: the first Adam steps move a weight by +-lr whatever the size of its gradient, so a last-bit difference of a
        # near-zero gradient flips a whole lr.  Hence 2e-3 here; bit-level checks use the tiny cases (one thread, test_oracle_golden.py)
        np.testing.assert_allclose(s8.numpy(), c.z["fwd8/scores"], rtol=2e-3, atol=2e-3)
        np.testing.assert_allclose(mu8.numpy(), c.z["fwd8/mu"], rtol=2e-3, atol=2e-3)
        _, tops = orc.full_ranking_eval(m, c.csr, c.data.test_items, (20,), np.arange(64))
        assert sum(np.array_equal(a, b) for a, b in zip(tops, c.z["top20"])) >= 62


def test_randn_forward_matches_reference():
    """Dense randn rows (negative values included), the input of the reference's tests/test_unit.py:153-198."""
    g = np.load(GOLDEN / "randn_forward.npz")
    np.random.seed(42)
    emb = np.random.randn(50, 384).astype(np.float32)
    emb = emb / np.linalg.norm(emb, axis=1, keepdims=True)
    torch.manual_seed(0)
    m = orc.OracleVAE(50, emb, 64, [128], 0.5, 0.2)
    x = torch.randn(4, 50)
    assert np.array_equal(x.numpy(), g["x"])
    m.eval()
    with torch.no_grad():
        s, mu, lv = m.forward_with(x, None)
    np.testing.assert_allclose(s.numpy(), g["scores"], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(mu.numpy(), g["mu"], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(lv.numpy(), g["logvar"], rtol=1e-6, atol=1e-6)
    loss = [float(v) for v in orc.loss_terms(s, x, mu, lv, 0.2)]
    np.testing.assert_allclose(loss, g["loss"], rtol=1e-6)


def tune_val_inputs():
    import pandas as pd
    from hvae_b200.synth import make_interactions, make_item_embeddings
    g = np.load(GOLDEN / "tune_val.npz")
    n_users, n_items, d, seed = int(g["n_users"]), int(g["n_items"]), int(g["d"]), int(g["seed"])
    data = make_interactions(n_users, n_items, seed)
    E = make_item_embeddings(n_items, d, seed)
    u2i = {f"u{i:07d}": i for i in range(n_users)}
    i2i = {f"i{i:07d}": i for i in range(n_items)}
    val_df = pd.DataFrame({"user_id": [f"u{i:07d}" for i in range(n_users)] + ["unknown_user"],
                           "asin": [f"i{int(data.test_items[i]):07d}" for i in range(n_users)] + ["i0000001"]})
    init = {k[5:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("init/")}
    ref = dict(zip([str(k) for k in g["keys"]], g["values"]))
    return g, data, E, u2i, i2i, val_df, init, ref


def test_tune_val_matches_reference():
    """src/ml/tune.py:121-184 with the negatives the reference drew: the oracle's ranking gives the reference's numbers."""
    g, data, E, u2i, i2i, val_df, init, ref = tune_val_inputs()
    m = orc.OracleVAE(int(g["n_items"]), E, 16, [40], 0.3, 0.2)
    sd = m.state_dict()
    sd.update(init)
    m.load_state_dict(sd)
    m.eval()
    csr = data.scipy_csr()
    neg = g["negatives"]
    acc = {k: {"recall": [], "ndcg": [], "hit_ratio": []} for k in (5, 10)}
    for u in range(int(g["n_users"])):
        s = orc.user_scores(m, csr, u)
        ranked = orc.negative_sampling_rank(s, int(data.test_items[u]), neg[u])
        rel = np.array([int(data.test_items[u])])
        for k in (5, 10):
            acc[k]["recall"].append(orc.recall_at_k(ranked[:k], rel, k))
            acc[k]["ndcg"].append(orc.ndcg_at_k(ranked[:k], rel, k))
            acc[k]["hit_ratio"].append(orc.hit_ratio_at_k(ranked[:k], rel, k))
    for k in (5, 10):
        for mname in ("recall", "ndcg", "hit_ratio"):
            np.testing.assert_allclose(np.mean(acc[k][mname]), ref[f"{mname}@{k}"], rtol=0, atol=1e-12)
