#!/usr/bin/env python
"""Standalone (materialised-score) top-K kernel of the fp32 mode: HBM roofline at the 1M-item shape."""
import json, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "recommendation-system_b200")]
import numpy as np, torch
from hvae_b200 import _cabi
lib = _cabi.lib(); dev = torch.device("cuda:0"); st = torch.cuda.current_stream().cuda_stream
pk = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"] if (ROOT / "MEASURED_PEAKS.json").exists() else 6650.0
for (R, N, K) in [(256, 1_000_000, 20), (1024, 200_000, 20), (4096, 12101, 20)]:
    S = torch.randn(R, N, device=dev)
    ip = torch.arange(0, 10 * (R + 1), 10, dtype=torch.int64, device=dev)
    ix = torch.sort(torch.randint(0, N, (R, 10), device=dev), dim=1)[0].to(torch.int32).reshape(-1).contiguous()
    val = torch.empty(R, K, device=dev); idx = torch.empty(R, K, dtype=torch.int32, device=dev)
    nc = int(lib.mask_topk_chunks(R, N))
    cv = torch.empty(R, nc * K, device=dev); ci = torch.empty(R, nc * K, dtype=torch.int32, device=dev)
    run = lambda: lib.mask_topk(S.data_ptr(), N, R, N, 0, ip.data_ptr(), ix.data_ptr(), None, 1, K, cv.data_ptr(), ci.data_ptr(), val.data_ptr(), idx.data_ptr(), st)
    run(); torch.cuda.synchronize()
    ts = []
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    for r in range(5):
        flush.fill_(r)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ms = float(np.median(ts)); gbs = R * N * 4 / ms / 1e6
    print(json.dumps({"kernel": "mask_topk(fp32)", "rows": R, "N": N, "K": K, "ms": ms, "GB/s": gbs, "frac_of_hbm_peak": gbs / pk}), flush=True)
