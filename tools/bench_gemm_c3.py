#!/usr/bin/env python
"""The nine MLP-stack GEMMs of a training step at the C3 (B=4096) and C2 (B=512) shapes through hvae_gemm_tf32, per tile width
(HVAE_TF32_BN=64|128|256 forces one; unset = the library's cost model).  CUDA events, 20 launches back to back."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "recommendation-system_b200")]
import torch
from hvae_b200 import _cabi
lib = _cabi.lib(); dev = torch.device("cuda:0")
r4 = lambda n: (n + 3) // 4 * 4
h, L, d = 600, 200, 768
def shapes(B):
    return {"fwd ml   [B,2L]=act W^T   K=h": (B, 2 * L, h, False, True), "fwd q    [B,d]=z W0^T     K=L": (B, d, L, False, True),
            "fwd u    [B,d]=t W3^T     K=d": (B, d, d, False, True), "dt       [B,d]=dU W3      K=d": (B, d, d, False, False),
            "dz       [B,L]=dq W0      K=d": (B, L, d, False, False), "dact     [B,h]=dml Wml    K=2L": (B, h, 2 * L, False, False),
            "dW3      [d,d]=dU^T t     K=B": (d, d, B, True, False), "dW0      [d,L]=dq^T z     K=B": (d, L, B, True, False),
            "dWml     [2L,h]=dml^T act K=B": (2 * L, h, B, True, False)}
for B in (4096, 512):
    tot = 0.0
    for name, (M, N, K, a_t, b_t) in shapes(B).items():
        A = torch.randn(K, r4(M), device=dev) if a_t else torch.randn(M, r4(K), device=dev)
        Bm = torch.randn(N, r4(K), device=dev) if b_t else torch.randn(K, r4(N), device=dev)
        a_rs, a_cs = (1, r4(M)) if a_t else (r4(K), 1)
        b_rs, b_cs = (1, r4(K)) if b_t else (r4(N), 1)
        C = torch.empty(M, r4(N), device=dev)
        run = lambda: lib.gemm_tf32(M, N, K, A.data_ptr(), a_rs, a_cs, Bm.data_ptr(), b_rs, b_cs, C.data_ptr(), r4(N), None, 1.0, torch.cuda.current_stream().cuda_stream)
        for _ in range(3): run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): run()
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 20 * 1000
        tot += us
        ref = (A.t()[:M, :K] if a_t else A[:, :K]).double() @ (Bm[:, :K].t() if b_t else Bm[:, :N]).double()
        err = float((C[:, :N].double() - ref).abs().max() / ref.abs().max())
        print(f"B={B:5d} {name}  {us:7.2f} us  {2.0 * M * N * K / us / 1e6:7.1f} TFLOP/s  rel.err {err:.1e}", flush=True)
    print(f"B={B}: sum of the nine GEMMs {tot:.1f} us", flush=True)
