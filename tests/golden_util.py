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
