"""CPU: host-side helpers of the training driver (hvae_b200/train.py) -- early stopping and the history file of the reference's
file contract (src/ml/train.py:286-336).  No kernel is launched."""
import json
import sys
from pathlib import Path
from types import SimpleNamespace

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "recommendation-system_b200"))


def test_early_stop_counts_epochs_without_improvement():
    from hvae_b200.train import _EarlyStop
    stop = _EarlyStop(patience=2)
    seen = []
    for val in (3.0, 2.5, 2.5, 2.6, 2.4, 2.41, 2.42):
        seen.append((stop.improved(val), stop.exhausted))
    # strictly lower is an improvement (an equal loss is not, as in the reference's `val_loss < best_val_loss`)
    assert [s[0] for s in seen] == [True, True, False, False, True, False, False]
    assert [s[1] for s in seen] == [False, False, False, True, False, False, True]
    assert stop.best == 2.4


def test_history_file_has_the_reference_keys(tmp_path):
    from hvae_b200.train import _write_history
    tr = SimpleNamespace(train_losses=[3.0, 2.0], val_losses=[3.5, 2.5], train_recon_losses=[2.9, 1.9], train_kl_losses=[0.5, 0.4])
    _write_history(tr, tmp_path / "training_history.json", 12.3456)
    doc = json.loads((tmp_path / "training_history.json").read_text())
    assert list(doc) == ["train_losses", "val_losses", "train_recon_losses", "train_kl_losses", "training_time_seconds"]
    assert doc["training_time_seconds"] == 12.35 and doc["val_losses"] == [3.5, 2.5]
