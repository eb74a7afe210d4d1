"""CPU: readers of the on-disk formats the hot path consumes (hvae_b200/data.py) against golden outputs
frozen from the reference's own load_training_data / _build_matrix / _build_input_matrix / get_user_indices_from_df
(oracle/make_golden_data.py): negatives rows (binary_rating 0), duplicate pairs (summed), users only in val, an item
without interactions, and the embeddings file with its mappings."""
import pickle

import numpy as np
import pandas as pd
from scipy.sparse import csr_matrix

from golden_util import GOLDEN
from hvae_b200 import data as hd


def _write(root, g):
    users, items = [str(u) for u in g["users"]], [str(i) for i in g["items"]]
    u2i, i2i = {u: i for i, u in enumerate(users)}, {a: i for i, a in enumerate(items)}
    frames = {}
    for name in ("train", "val", "test"):
        frames[name] = pd.DataFrame({c: g[f"{name}/{c}"] for c in ("user_id", "asin", "binary_rating", "rating")})
        frames[name].to_csv(root / f"{name}.csv", index=False)
    with open(root / "interaction_matrix.pkl", "wb") as f:
        pickle.dump(csr_matrix(g["matrix"]), f)
    with open(root / "mappings.pkl", "wb") as f:
        pickle.dump({"user_to_idx": u2i, "item_to_idx": i2i, "idx_to_user": {v: k for k, v in u2i.items()},
                     "idx_to_item": {v: k for k, v in i2i.items()}}, f)
    emb = np.arange(len(items) * 6, dtype=np.float32).reshape(len(items), 6)
    np.save(root / "item_embeddings.npy", emb)
    with open(root / "item_embeddings_mappings.pkl", "wb") as f:
        pickle.dump({"item_to_idx": i2i, "idx_to_item": {v: k for k, v in i2i.items()}}, f)
    return u2i, i2i, emb


def test_readers_match_reference(tmp_path):
    g = np.load(GOLDEN / "data_files.npz", allow_pickle=False)
    u2i, i2i, emb = _write(tmp_path, g)
    matrix, train_df, val_df, mappings = hd.load_training_data(str(tmp_path))
    assert mappings["user_to_idx"] == u2i and mappings["item_to_idx"] == i2i
    np.testing.assert_array_equal(matrix.toarray(), g["matrix"])
    shape = matrix.shape
    tm = hd.build_matrix(train_df, u2i, i2i, shape)
    assert tm.dtype == np.float64 and tm.max() > 1.0                    # duplicate pairs are summed, as in the reference
    np.testing.assert_array_equal(tm.toarray(), g["train_matrix"])
    np.testing.assert_array_equal(hd.build_matrix(val_df, u2i, i2i, shape).toarray(), g["val_matrix"])
    np.testing.assert_array_equal(hd.build_input_matrix(train_df, val_df, u2i, i2i, shape).toarray(), g["input_matrix"])
    assert hd.get_user_indices_from_df(train_df, u2i) == g["train_users"].tolist()
    assert hd.get_user_indices_from_df(val_df, u2i) == g["val_users"].tolist()
    assert hd.get_user_indices_from_df(pd.DataFrame({"user_id": ["nobody", str(g["users"][3])]}), u2i) == [3]
    assert g["train_matrix"][:, -1].sum() == 0 and g["train_matrix"][-2:].sum() == 0   # item / users without train interactions


def test_embeddings_reader(tmp_path):
    g = np.load(GOLDEN / "data_files.npz", allow_pickle=False)
    _, i2i, emb = _write(tmp_path, g)
    e, item_to_idx, idx_to_item = hd.load_embeddings(tmp_path / "item_embeddings.npy")
    np.testing.assert_array_equal(e, emb)
    assert item_to_idx == i2i and idx_to_item[0] == str(g["items"][0])
    (tmp_path / "item_embeddings_mappings.pkl").unlink()
    e2, a, b = hd.load_embeddings(tmp_path / "item_embeddings.npy")
    assert a is None and b is None and e2.shape == emb.shape
