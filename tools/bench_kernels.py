#!/usr/bin/env python
"""Kernel-level timings of the tcgen05 scoring kernels at a named shape (CUDA events, L2 flushed between reps).
    python tools/bench_kernels.py [--B 4096 --N 200000 --d 768 --K 20 --reps 5]
Prints one JSON line per kernel with achieved TFLOP/s against MEASURED_PEAKS.json."""
import argparse
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "recommendation-system_b200")]
import numpy as np  # noqa: E402
import torch  # noqa: E402

from hvae_b200 import _cabi  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=4096)
    ap.add_argument("--N", type=int, default=200_000)
    ap.add_argument("--d", type=int, default=768)
    ap.add_argument("--K", type=int, default=20)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    lib = _cabi.lib()
    st = torch.cuda.current_stream().cuda_stream
    pk = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {"bf16_tflops": 1590.0}
    B, N, d, K = a.B, a.N, a.d, a.K
    ld = (d + 7) // 8 * 8
    g = torch.Generator(device=dev).manual_seed(0)
    U = (torch.randn(B, ld, generator=g, device=dev) * 0.3).to(torch.bfloat16)
    E = torch.nn.functional.normalize(torch.randn(N, ld, generator=g, device=dev), dim=1).to(torch.bfloat16)
    ns = int(lib.tc_n_splits(B, N))
    ws = torch.empty(2 * B * ns, device=dev)
    lse = torch.empty(B, device=dev)
    gs = int(lib.tc_grad_splits(B, N, d))
    ldo = (d + 3) // 4 * 4
    Op = torch.empty(gs, B, ldo, device=dev)
    nst = int(lib.tc_topk_splits(B, N))
    cv = torch.empty(B, nst * K, device=dev)
    ci = torch.empty(B, nst * K, dtype=torch.int32, device=dev)
    indptr = torch.arange(0, 10 * (B + 1), 10, dtype=torch.int64, device=dev)
    indices = torch.sort(torch.randint(0, N, (B, 10), device=dev, generator=g), dim=1)[0].to(torch.int32).reshape(-1).contiguous()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    n_sub = int(lib.tc_onepass_subparts(d))
    # S = U E^T and O = P E, 2BNd each; beyond d = 768 every 384-column chunk of O recomputes S
    executed_bwd = 4.0 * B * N * d if d <= 768 else 2.0 * B * N * d * (1 + (d + 383) // 384)
    c_part = torch.empty(gs, B, device=dev)
    l_part = torch.empty(gs, n_sub, B, device=dev)
    kernels = {
        "tc_score_onepass": (lambda: lib.tc_score_onepass(U.data_ptr(), ld, B, E.data_ptr(), ld, N, d, c_part.data_ptr(), l_part.data_ptr(),
                                                          Op.data_ptr(), ldo, st), 4.0 * B * N * d, executed_bwd),
        "tc_score_lse": (lambda: lib.tc_score_lse(U.data_ptr(), ld, B, E.data_ptr(), ld, N, d, lse.data_ptr(), ws.data_ptr(), st), 2.0 * B * N * d, 2.0 * B * N * d),
        "tc_score_grad": (lambda: lib.tc_score_grad(U.data_ptr(), ld, B, E.data_ptr(), ld, N, d, lse.data_ptr(), Op.data_ptr(), ldo, st),
                          2.0 * B * N * d, executed_bwd),
        "tc_score_topk": (lambda: lib.tc_score_topk(U.data_ptr(), ld, B, E.data_ptr(), ld, N, d, 0, indptr.data_ptr(), indices.data_ptr(), None, K,
                                                    cv.data_ptr(), ci.data_ptr(), st), 2.0 * B * N * d, 2.0 * B * N * d),
    }
    for name, (fn, algo, executed) in kernels.items():
        if a.only and a.only != name:
            continue
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        ts = []
        for r in range(a.reps):
            flush.fill_(r)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = float(np.median(ts))
        print(json.dumps({"kernel": name, "B": B, "N": N, "d": d, "ms": ms, "algorithmic_tflops": algo / ms / 1e9,
                          "executed_tflops": executed / ms / 1e9, "peak_tflops_burst": pk["bf16_tflops"],
                          "frac_algorithmic": algo / ms / 1e9 / pk["bf16_tflops"], "frac_executed": executed / ms / 1e9 / pk["bf16_tflops"]}), flush=True)


if __name__ == "__main__":
    main()
