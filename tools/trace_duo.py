#!/usr/bin/env python
"""Cycle trace of the cta_group::2 scoring kernel (hvae_tc_duo_trace): where the producer / MMA issuer / softmax warps wait.
    python tools/trace_duo.py [--B 4096 --N 200000 --d 768]"""
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
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    lib = _cabi.lib()
    st = torch.cuda.current_stream().cuda_stream
    B, N, d = a.B, a.N, a.d
    ld = (d + 7) // 8 * 8
    g = torch.Generator(device=dev).manual_seed(0)
    U = (torch.randn(B, ld, generator=g, device=dev) * 0.3).to(torch.bfloat16)
    E = torch.nn.functional.normalize(torch.randn(N, ld, generator=g, device=dev), dim=1).to(torch.bfloat16)
    gs = int(lib.tc_grad_splits(B, N, d))
    ldo = (d + 3) // 4 * 4
    Op = torch.empty(gs, B, ldo, device=dev)
    c_part, l_part = torch.empty(gs, B, device=dev), torch.empty(gs, 2, B, device=dev)
    m_tiles = (B + 127) // 128
    n_cta = 2 * m_tiles * gs
    trace = torch.zeros(n_cta, 3, 8, dtype=torch.int64, device=dev)
    run = lambda: lib.tc_score_onepass(U.data_ptr(), ld, B, E.data_ptr(), ld, N, d, c_part.data_ptr(), l_part.data_ptr(), Op.data_ptr(), ldo, st)
    run(); run()
    torch.cuda.synchronize()
    lib.tc_duo_trace(trace.data_ptr())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(); e1.record()
    torch.cuda.synchronize()
    lib.tc_duo_trace(None)
    t = trace.cpu().numpy().astype(np.float64)
    tiles = -(-N // 128) / gs
    names = {0: ["ring slot free"], 1: ["S buffer free", "G1 operands landed", "P written", "G2 operands landed"],
             2: ["S ready", "P buffer free", "sweep MMAs done"]}
    out = {"ms_traced": e0.elapsed_time(e1), "ctas": n_cta, "tiles_per_cta": tiles}
    for role, rn in ((0, "producer"), (1, "mma"), (2, "softmax")):
        for rank in (0, 1):
            sel = t[rank::2, role, :]
            life = sel[:, 7].mean()
            out[f"{rn}_cta{rank}"] = {"lifetime_cycles": life, "cycles_per_tile": life / tiles,
                                      **{n: {"frac": float(sel[:, i].mean() / life), "cycles_per_tile": float(sel[:, i].mean() / tiles)}
                                         for i, n in enumerate(names[role])}}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
