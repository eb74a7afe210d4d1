#!/usr/bin/env python
"""Encoder layer-1 gather-sum kernel (hvae_gather_ln_fwd) alone: achieved HBM GB/s at a named shape.
    python tools/bench_gather.py [--B 65536 --N 1000000 --h 600 --reps 5]
Algorithmic bytes per user = nnz*(ld*4 + 4 [index]) read + 2*ld*4 (pre, act) + h (dropout mask) + 16 (indptr) written/read.
Rows are drawn from the Zipf popularity of the synthetic generator, so popular rows are L2 hits: DRAM traffic <= algorithmic."""
import argparse
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "recommendation-system_b200")]
import numpy as np  # noqa: E402
import torch  # noqa: E402

from hvae_b200 import _cabi, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=65536)
    ap.add_argument("--N", type=int, default=1_000_000)
    ap.add_argument("--h", type=int, default=600)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--uniform", action="store_true", help="uniform item popularity (no L2-resident hot rows)")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    lib = _cabi.lib()
    st = torch.cuda.current_stream().cuda_stream
    pk = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"] if (ROOT / "MEASURED_PEAKS.json").exists() else 6550.7
    B, N, h = a.B, a.N, a.h
    ld = (h + 3) // 4 * 4
    if a.uniform:
        rng = np.random.default_rng(0)
        cnt = np.clip(np.rint(rng.lognormal(np.log(8.0), 0.6, B)), 3, 200).astype(np.int64)
        indptr = np.zeros(B + 1, np.int64); np.cumsum(cnt, out=indptr[1:])
        indices = rng.integers(0, N, int(cnt.sum())).astype(np.int32)
    else:
        data = synth.make_interactions(B, N, seed=0)
        indptr, indices = data.indptr, data.indices
    nnz = int(indptr[-1])
    ip = torch.from_numpy(indptr).to(dev)
    ix = torch.from_numpy(indices).to(dev)
    W = torch.randn(N, ld, device=dev) * 0.01
    bias = torch.zeros(ld, device=dev); gamma = torch.ones(ld, device=dev); beta = torch.zeros(ld, device=dev)
    mask = (torch.rand(B, h, device=dev) < 0.5).to(torch.uint8)
    pre = torch.empty(B, ld, device=dev); act = torch.empty(B, ld, device=dev)
    mean = torch.empty(B, device=dev); rstd = torch.empty(B, device=dev)
    run = lambda: lib.gather_ln_fwd(ip.data_ptr(), ix.data_ptr(), None, None, B, W.data_ptr(), ld, h, bias.data_ptr(), gamma.data_ptr(),
                                    beta.data_ptr(), mask.data_ptr(), 2.0, pre.data_ptr(), mean.data_ptr(), rstd.data_ptr(), act.data_ptr(), st)
    run(); torch.cuda.synchronize()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    ts = []
    for r in range(a.reps):
        flush.fill_(r)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ms = float(np.median(ts))
    algo = nnz * (ld * 4 + 4) + B * (2 * ld * 4 + h + 16)
    uniq = int(np.unique(indices).shape[0])
    print(json.dumps({"kernel": "gather_ln_fwd", "B": B, "N": N, "h": h, "nnz": nnz, "unique_rows": uniq, "ms": ms, "algorithmic_bytes": algo,
                      "GB/s": algo / ms / 1e6, "peak_GB/s": pk, "frac_of_hbm_peak": algo / ms / 1e6 / pk,
                      "popularity": "uniform" if a.uniform else "zipf"}), flush=True)


if __name__ == "__main__":
    main()
