"""Helpers shared by the CPU and GPU parity tests: load a golden case and rebuild its inputs."""
from pathlib import Path

import numpy as np
import torch
from scipy.sparse import csr_matrix

GOLDEN = Path(__file__).resolve().parent / "golden"
TRAIN_CASES = ["tiny_two_hidden", "tiny_identity", "tiny_annealed"]


class Case:
    def __init__(self, name):
        self.name = name
        self.z = np.load(GOLDEN / f"{name}.npz")
        z = self.z
        self.n_users, self.n_items, self.d = int(z["n_users"]), int(z["n_items"]), int(z["d"])
        self.latent, self.hidden = int(z["latent"]), [int(h) for h in z["hidden"]]
        self.dropout, self.beta = float(z["dropout"]), float(z["beta"])
        self.batch, self.steps, self.seed = int(z["batch"]), int(z["steps"]), int(z["seed"])
        self.annealing, self.weight_decay = bool(z["annealing"]), float(z["weight_decay"])
        self.csr = csr_matrix((z["values"].astype(np.float64), z["indices"], z["indptr"]),
                              shape=(self.n_users, self.n_items))
        self.test_items = z["test_items"]

    def embeddings(self):
        from hvae_b200.synth import make_item_embeddings
        return make_item_embeddings(self.n_items, self.d, self.seed)

    def state(self, prefix):
        return {k[len(prefix) + 1:]: torch.from_numpy(self.z[k]) for k in self.z.files if k.startswith(prefix + "/")}

    def rows(self, s):
        return self.z[f"rows/{s}"]

    def noise(self, s):
        z = self.z
        masks = [torch.from_numpy(z[f"noise/{s}/mask{i}"]).float() for i in range(len(self.hidden))]
        pm = f"noise/{s}/pmask"
        return dict(masks=masks, eps=torch.from_numpy(z[f"noise/{s}/eps"]),
                    pmask=torch.from_numpy(z[pm]).float() if pm in z.files else None)

    def dense(self, rows):
        return torch.from_numpy(np.asarray(self.csr[rows].toarray(), dtype=np.float32))

    def model_kwargs(self):
        return dict(n_items=self.n_items, item_embeddings=self.embeddings(), latent_dim=self.latent,
                    hidden_dims=self.hidden, dropout=self.dropout, beta=self.beta)

    def beta_at(self, step):
        if not self.annealing:
            return self.beta
        return min(self.beta, 0.0 + (step / 4) * self.beta) if step < 4 else self.beta


N_SAMPLE = 48


def sample_positions(shape, seed):
    """The fixed pseudo-random flat positions oracle/make_golden_big.py sampled inside a tensor of `shape`."""
    n = int(np.prod(shape))
    return np.random.default_rng(seed).integers(0, n, size=min(N_SAMPLE, n))


class BigCase:
    """A golden case whose inputs regenerate from seeds and whose tensors are stored as digests (oracle/make_golden_big.py)."""

    def __init__(self, name):
        self.name = name
        self.z = np.load(GOLDEN / f"{name}.npz")
        z = self.z
        self.n_users, self.n_items, self.d = int(z["n_users"]), int(z["n_items"]), int(z["d"])
        self.latent, self.hidden = int(z["latent"]), [int(h) for h in z["hidden"]]
        self.dropout, self.beta, self.seed, self.steps = float(z["dropout"]), float(z["beta"]), int(z["seed"]), int(z["steps"])
        from hvae_b200.synth import make_interactions, make_item_embeddings
        self.data = make_interactions(self.n_users, self.n_items, self.seed)
        self.csr = self.data.scipy_csr()
        self.E = make_item_embeddings(self.n_items, self.d, self.seed)
        self.stats, self.validate = z["stats"], z["validate"]
        self.stats64 = z["stats64"]          # the same steps evaluated in float64 (oracle in double, same noise)

    def model_kwargs(self):
        return dict(n_items=self.n_items, item_embeddings=self.E, latent_dim=self.latent, hidden_dims=self.hidden,
                    dropout=self.dropout, beta=self.beta)

    def rows(self, s):
        return self.z[f"rows/{s}"].astype(np.int64)

    def noise(self, s):
        """Keep-masks as float {0,1} tensors, eps: what the reference consumed in step s."""
        z, B = self.z, len(self.rows(s))
        unpack = lambda a, w: torch.from_numpy(np.unpackbits(a, axis=1)[:, :w].astype(np.float32))
        return dict(masks=[unpack(z[f"noise/{s}/mask{i}"], h) for i, h in enumerate(self.hidden)],
                    eps=torch.from_numpy(z[f"noise/{s}/eps"]), pmask=unpack(z[f"noise/{s}/pmask"], self.d))

    def check_digest(self, prefix, tensors, rtol, atol_scale=1e-6, seed=12345):
        """Compare tensors (name -> array-like, reference shapes) with the stored sum / |.|-sum / sampled entries.
        Order = the reference state_dict's (without item_embeddings) or named_parameters', as generated."""
        z = self.z
        keys = [k[len(prefix) + 1:-len("/sum")] for k in z.files if k.startswith(prefix + "/") and k.endswith("/sum")]
        assert keys, prefix
        for i, k in enumerate(keys):
            a = np.asarray(tensors[k].detach().cpu().double().numpy() if hasattr(tensors[k], "detach") else tensors[k], dtype=np.float64)
            ref_abs = float(z[f"{prefix}/{k}/abs"])
            scale = ref_abs / max(1, a.size)                      # mean |entry|
            np.testing.assert_allclose(np.abs(a).sum(), ref_abs, rtol=rtol, atol=atol_scale * a.size, err_msg=f"{prefix}/{k} abs")
            np.testing.assert_allclose(a.sum(), float(z[f"{prefix}/{k}/sum"]), rtol=0, atol=rtol * ref_abs + atol_scale * a.size,
                                       err_msg=f"{prefix}/{k} sum")
            got = a.reshape(-1)[sample_positions(a.shape, seed + i)]
            np.testing.assert_allclose(got, z[f"{prefix}/{k}/sample"], rtol=rtol * 10, atol=rtol * 10 * scale + 1e-12,
                                       err_msg=f"{prefix}/{k} samples")
