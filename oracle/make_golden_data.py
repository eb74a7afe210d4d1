"""Golden vectors for the on-disk formats of the hot path, frozen from the REFERENCE's own readers
(run in the build container only; TEST INFRASTRUCTURE, not product code).

    python oracle/make_golden_data.py        # writes tests/golden/data_files.npz

A tiny dataset directory is written exactly as the reference's offline pipeline writes it
(src/preprocessing/dataset.py:137-179: train/val/test.csv, interaction_matrix.pkl, mappings.pkl;
src/preprocessing/embeddings.py:93-131: item_embeddings.npy + _mappings.pkl), with the awkward cases the reference's
code has to deal with: binary_rating 0 rows (sampled negatives), duplicate (user, item) pairs (summed by
csr_matrix), users that appear only in val, an item without interactions.  The reference's load_training_data,
get_user_indices_from_df, _build_matrix (src/ml/train.py:153-182) and _build_input_matrix (src/ml/evaluate.py:73-87)
then read it; inputs (as column arrays) and outputs (dense) are stored.  tests/test_data_cpu.py rebuilds the files from
the stored columns and checks hvae_b200.data against these outputs.
"""
from __future__ import annotations

import pickle
import sys
import tempfile
from pathlib import Path

import numpy as np
import pandas as pd
from scipy.sparse import csr_matrix

sys.path.insert(0, str(Path(__file__).resolve().parent))
from make_golden import OUT, _import_reference  # noqa: E402


def make_frames(seed=0, U=12, N=9):
    rng = np.random.default_rng(seed)
    users = [f"A{u:03d}X" for u in range(U)]
    items = [f"B{i:05d}" for i in range(N)]
    def frame(n, user_pool, with_neg):
        u = rng.choice(user_pool, n)
        it = rng.choice(items[:-1], n)                      # the last item never interacts
        br = (rng.random(n) < (0.75 if with_neg else 1.1)).astype(np.int64)
        return pd.DataFrame({"user_id": u, "asin": it, "binary_rating": br, "rating": rng.integers(1, 6, n).astype(float)})
    train = frame(60, users[:-2], True)
    train = pd.concat([train, train.iloc[:5]], ignore_index=True)      # duplicate pairs
    val = frame(14, users, False)                                      # two users only here
    test = frame(12, users[:-2], False)
    return users, items, train, val, test


def write_dir(root: Path, users, items, train, val, test):
    u2i = {u: i for i, u in enumerate(users)}
    i2i = {a: i for i, a in enumerate(items)}
    train.to_csv(root / "train.csv", index=False)
    val.to_csv(root / "val.csv", index=False)
    test.to_csv(root / "test.csv", index=False)
    pos = pd.concat([train, val, test])
    pos = pos[pos["binary_rating"] == 1]
    full = csr_matrix((np.ones(len(pos)), (pos["user_id"].map(u2i), pos["asin"].map(i2i))), shape=(len(users), len(items)))
    with open(root / "interaction_matrix.pkl", "wb") as f:
        pickle.dump(full, f)
    mappings = {"user_to_idx": u2i, "item_to_idx": i2i, "idx_to_user": {v: k for k, v in u2i.items()},
                "idx_to_item": {v: k for k, v in i2i.items()}}
    with open(root / "mappings.pkl", "wb") as f:
        pickle.dump(mappings, f)
    emb = np.random.default_rng(1).standard_normal((len(items), 6)).astype(np.float32)
    np.save(root / "item_embeddings.npy", emb)
    with open(root / "item_embeddings_mappings.pkl", "wb") as f:
        pickle.dump({"item_to_idx": i2i, "idx_to_item": mappings["idx_to_item"]}, f)
    return u2i, i2i


def main():
    _, ref_train, ref_eval = _import_reference()
    users, items, train, val, test = make_frames()
    root = Path(tempfile.mkdtemp(prefix="hvae_data_"))
    u2i, i2i = write_dir(root, users, items, train, val, test)
    matrix, tr, va, mp = ref_train.load_training_data(str(root))
    shape = matrix.shape
    out = {"users": np.array(users), "items": np.array(items)}
    for name, df in (("train", train), ("val", val), ("test", test)):
        for col in df.columns:
            a = df[col].to_numpy()
            out[f"{name}/{col}"] = a.astype(str) if a.dtype == object else a      # no pickled objects in the fixture
    out["matrix"] = matrix.toarray()
    out["train_matrix"] = ref_train._build_matrix(tr, mp["user_to_idx"], mp["item_to_idx"], shape).toarray()
    out["val_matrix"] = ref_train._build_matrix(va, mp["user_to_idx"], mp["item_to_idx"], shape).toarray()
    out["input_matrix"] = ref_eval._build_input_matrix(tr, va, mp["user_to_idx"], mp["item_to_idx"], shape).toarray()
    out["train_users"] = np.array(ref_train.get_user_indices_from_df(tr, mp["user_to_idx"]), dtype=np.int64)
    out["val_users"] = np.array(ref_train.get_user_indices_from_df(va, mp["user_to_idx"]), dtype=np.int64)
    np.savez_compressed(OUT / "data_files.npz", **out)
    print("wrote", OUT / "data_files.npz", {k: v.shape for k, v in out.items() if k.endswith("matrix")})


if __name__ == "__main__":
    main()
