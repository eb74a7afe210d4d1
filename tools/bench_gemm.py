#!/usr/bin/env python
"""Hot-loop timing of the MLP-stack GEMM kernels (TF32 tcgen05 vs fp32 FFMA) at the C2 shapes, CUDA events."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "recommendation-system_b200")]
import torch
from hvae_b200 import _cabi
lib = _cabi.lib(); dev = torch.device("cuda:0"); st = torch.cuda.current_stream().cuda_stream
r4 = lambda n: (n + 3) // 4 * 4
shapes = {"K640al  [512x400,K640] NT": (512, 400, 640, False, True), "K608    [512x400,K608] NT": (512, 400, 608, False, True),
          "K32     [512x400,K32 ] NT": (512, 400, 32, False, True), "K128    [512x400,K128] NT": (512, 400, 128, False, True),
          "K256    [512x400,K256] NT": (512, 400, 256, False, True), "K1200   [512x400,K1200] NT": (512, 400, 1200, False, True),
          "big     [4096x640,K600] NT": (4096, 640, 600, False, True),
          "fwd ml  [512x400,K600] NT": (512, 400, 600, False, True), "dX      [512x600,K400] NN": (512, 600, 400, False, False),
          "dW      [400x600,K512] TN": (400, 600, 512, True, False), "proj3   [512x384,K384] NT": (512, 384, 384, False, True)}
for name, (M, N, K, a_t, b_t) in shapes.items():
    A = torch.randn(K, r4(M), device=dev) if a_t else torch.randn(M, r4(K), device=dev)
    B = torch.randn(N, r4(K), device=dev) if b_t else torch.randn(K, r4(N), device=dev)
    a_rs, a_cs = (1, r4(M)) if a_t else (r4(K), 1)
    b_rs, b_cs = (1, r4(K)) if b_t else (r4(N), 1)
    C = torch.empty(M, r4(N), device=dev)
    for fn_name in ("gemm_tf32", "gemm_f32"):
        fn = getattr(lib, fn_name)
        run = lambda: fn(M, N, K, A.data_ptr(), a_rs, a_cs, B.data_ptr(), b_rs, b_cs, C.data_ptr(), r4(N), None, 1.0, torch.cuda.current_stream().cuda_stream)
        for _ in range(5): run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(50): run()
        g.replay(); torch.cuda.synchronize()
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        print(f"{name}  {fn_name:10s} {e0.elapsed_time(e1) / 50 * 1000:7.2f} us/launch (hot, back-to-back in a graph)", flush=True)
