"""The fused-epilogue GEMMs of one training step at a given batch size, a few rounds (for ncu / CUDA-event timing)."""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "recommendation-system_b200"))
from hvae_b200 import _cabi

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
d = int(sys.argv[2]) if len(sys.argv) > 2 else 384
L = 200
lib, dev = _cabi.lib(), torch.device("cuda:0")
st = torch.cuda.current_stream().cuda_stream
g = torch.Generator().manual_seed(0)
rnd = lambda *s: torch.randn(*s, generator=g).to(dev)
z, W0, b0 = rnd(B, L), rnd(d, L) * L ** -0.5, rnd(d)
q, t, dq, dU = torch.empty(B, d, device=dev), torch.empty(B, d, device=dev), torch.empty(B, d, device=dev), rnd(B, d)
W3 = rnd(d, d) * d ** -0.5
mask = (torch.rand(B, d, generator=g) > 0.3).to(torch.uint8).to(dev)
ml, eps, dml = rnd(B, 2 * L) * 0.5, rnd(B, L), torch.empty(B, 2 * L, device=dev)
coef = torch.tensor([0.2 / B], device=dev)
cs0, cs1 = torch.empty(d, device=dev), torch.empty(2 * L, device=dev)
w0 = torch.zeros(int(lib.gemm_colsum_workspace_floats(B, d)), device=dev)
w1 = torch.zeros(int(lib.gemm_colsum_workspace_floats(B, 2 * L)), device=dev)
p = lambda x: x.data_ptr()


def fwd(): lib.gemm_tf32_gelu_drop(B, d, L, p(z), L, 1, p(W0), 1, L, p(q), p(t), d, p(b0), p(mask), 1 / 0.7, st)
def plain(): lib.gemm_tf32(B, d, L, p(z), L, 1, p(W0), 1, L, p(q), d, p(b0), 1.0, st)
def bwd(): lib.gemm_tf32_gelu_bwd(B, d, d, p(dU), d, 1, p(W3), d, 1, p(dq), d, p(q), p(mask), 1 / 0.7, p(cs0), p(w0), st)
def bwd_nocs(): lib.gemm_tf32_gelu_bwd(B, d, d, p(dU), d, 1, p(W3), d, 1, p(dq), d, p(q), p(mask), 1 / 0.7, None, None, st)
def plain_bwd(): lib.gemm_tf32(B, d, d, p(dU), d, 1, p(W3), d, 1, p(dq), d, None, 1.0, st)
def lat(): lib.gemm_tf32_latent_bwd(B, L, d, p(dq), d, 1, p(W0), L, 1, p(ml), 2 * L, p(eps), p(coef), p(dml), p(cs1), p(w1), st)
def plain_lat(): lib.gemm_tf32(B, L, d, p(dq), d, 1, p(W0), L, 1, p(dml), 2 * L, None, 1.0, st)


for name, fn in [("plain fwd", plain), ("gelu_drop", fwd), ("plain bwd", plain_bwd), ("gelu_bwd", bwd), ("gelu_bwd no colsum", bwd_nocs),
                 ("plain dz", plain_lat), ("latent_bwd", lat)]:
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print(f"B={B} d={d} {name:20s} {e0.elapsed_time(e1) / 50 * 1000:8.2f} us")
