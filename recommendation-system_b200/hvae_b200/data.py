"""Readers for the on-disk inputs of the hot path: the files the reference's offline
pipeline writes (src/preprocessing/dataset.py:137-179, src/preprocessing/embeddings.py:93-131) and the
helpers the reference's trainer/evaluator use to turn them into matrices
(src/ml/train.py:153-182, src/ml/evaluate.py:73-87)."""
from __future__ import annotations

import logging
import pickle
from pathlib import Path

import numpy as np
import pandas as pd
from scipy.sparse import csr_matrix

logger = logging.getLogger(__name__)


def load_embeddings(embeddings_path, mappings_path=None):
    """-> (embeddings [N,d], item_to_idx | None, idx_to_item | None); src/preprocessing/embeddings.py:134-158."""
    embeddings_path = Path(embeddings_path)
    emb = np.load(embeddings_path)
    if mappings_path is None:
        mappings_path = embeddings_path.with_name(f"{embeddings_path.stem}_mappings.pkl")
    item_to_idx = idx_to_item = None
    if Path(mappings_path).exists():
        with open(mappings_path, "rb") as f:
            m = pickle.load(f)
        item_to_idx, idx_to_item = m.get("item_to_idx"), m.get("idx_to_item")
    return emb, item_to_idx, idx_to_item


def load_training_data(data_dir: str):
    """-> (interaction_matrix, train_df, val_df, mappings); src/ml/train.py:153-167."""
    path = Path(data_dir)
    with open(path / "interaction_matrix.pkl", "rb") as f:
        matrix = pickle.load(f)
    train_df = pd.read_csv(path / "train.csv", low_memory=False)
    val_df = pd.read_csv(path / "val.csv", low_memory=False)
    with open(path / "mappings.pkl", "rb") as f:
        mappings = pickle.load(f)
    logger.info("Loaded: matrix %s, train %d, val %d", matrix.shape, len(train_df), len(val_df))
    return matrix, train_df, val_df, mappings


def get_user_indices_from_df(df: pd.DataFrame, user_to_idx: dict) -> list:
    """src/ml/train.py:170-172."""
    return [user_to_idx[uid] for uid in df["user_id"].unique() if uid in user_to_idx]


def build_matrix(df: pd.DataFrame, user_to_idx: dict, item_to_idx: dict, shape) -> csr_matrix:
    """Positives of `df` as CSR; duplicate (user,item) pairs sum (src/ml/train.py:175-182)."""
    pos = df[df["binary_rating"] == 1] if "binary_rating" in df.columns else df
    rows, cols = pos["user_id"].map(user_to_idx), pos["asin"].map(item_to_idx)
    return csr_matrix((np.ones(len(pos)), (rows, cols)), shape=shape)


def build_input_matrix(train_df, val_df, user_to_idx, item_to_idx, shape) -> csr_matrix:
    """train+val positives: encoder input and seen-mask of the evaluator (src/ml/evaluate.py:73-87)."""
    tp = train_df[train_df["binary_rating"] == 1] if "binary_rating" in train_df.columns else train_df
    vp = val_df[val_df["binary_rating"] == 1] if "binary_rating" in val_df.columns else val_df
    both = pd.concat([tp, vp])
    rows, cols = both["user_id"].map(user_to_idx), both["asin"].map(item_to_idx)
    return csr_matrix((np.ones(len(both)), (rows, cols)), shape=shape)
