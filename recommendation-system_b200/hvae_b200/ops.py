"""PyTorch custom ops (`torch.ops.hvae_b200.*`) over the C ABI of libhvae_b200.so.

The engine sequences the kernels through ctypes directly (one Python call per launch, no dispatcher overhead, graph
capturable); these registrations expose the same entry points as first-class torch operators -- schema-checked, usable
from other torch code, traceable with fake tensors -- for the hot ops of the path:

    gather_ln_fwd   encoder layer 1 as a CSR gather-sum (+ LayerNorm/GELU/dropout)      src/ml/model.py:114-117,149
    gemm            MLP-stack GEMM (tcgen05 TF32 or fp32 FFMA)                            src/ml/model.py:90-95,126-127
    score_lse       log-sum-exp of u E^T over all items, scores never materialised       src/ml/model.py:198,281
    score_grad      softmax(u E^T) E with the scores recomputed on the tensor cores      autograd of model.py:198,281
    score_topk      per-user top-K over all items with the seen items masked             src/ml/evaluate.py:125-147
    adam_step       clip + Adam over the parameter arena                                  src/ml/train.py:91-92

Every op raises on non-CUDA tensors: there is no CPU fallback.
"""
from __future__ import annotations

import torch

from . import _cabi
from ._cabi import p

_NS = "hvae_b200"


def _cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("hvae_b200 ops run on CUDA tensors only (there is no CPU fallback)")


def _stream(t):
    return torch.cuda.current_stream(t.device).cuda_stream


def _r4(n):
    return (n + 3) // 4 * 4


@torch.library.custom_op(f"{_NS}::gather_ln_fwd", mutates_args=())
def gather_ln_fwd(indptr: torch.Tensor, indices: torch.Tensor, rows: torch.Tensor, W1T: torch.Tensor, bias: torch.Tensor,
                  gamma: torch.Tensor, beta: torch.Tensor, h: int) -> torch.Tensor:
    """act [B, ld] = GELU(LayerNorm(sum_j W1T[indices_j] + bias)) for the users `rows` of a binary CSR (eval mode: no dropout)."""
    _cuda(indptr, indices, rows, W1T, bias, gamma, beta)
    B, ld = rows.shape[0], W1T.shape[1]
    act = torch.empty(B, ld, dtype=torch.float32, device=W1T.device)
    pre, mean, rstd = torch.empty_like(act), torch.empty(B, device=W1T.device), torch.empty(B, device=W1T.device)
    _cabi.lib().gather_ln_fwd(p(indptr), p(indices), None, p(rows), B, p(W1T), ld, h, p(bias), p(gamma), p(beta), None, 1.0,
                              p(pre), p(mean), p(rstd), p(act), _stream(W1T))
    return act


@gather_ln_fwd.register_fake
def _(indptr, indices, rows, W1T, bias, gamma, beta, h):
    return W1T.new_empty(rows.shape[0], W1T.shape[1])


@torch.library.custom_op(f"{_NS}::gemm", mutates_args=())
def gemm(A: torch.Tensor, B: torch.Tensor, bias: torch.Tensor | None, tensor_cores: bool) -> torch.Tensor:
    """C = A @ B^T (+ bias): A [M, K], B [N, K] row-major fp32 with leading dimensions % 4 == 0 (nn.Linear's layout)."""
    _cuda(A, B, bias)
    M, K = A.shape
    N = B.shape[0]
    C = torch.empty(M, _r4(N), dtype=torch.float32, device=A.device)
    lib = _cabi.lib()
    fn = lib.gemm_tf32 if tensor_cores and lib.gemm_tf32_supported(p(A), A.stride(0), 1, p(B), 1, B.stride(0)) else lib.gemm_f32
    fn(M, N, K, p(A), A.stride(0), 1, p(B), 1, B.stride(0), p(C), C.stride(0), p(bias), 1.0, _stream(A))
    return C[:, :N]


@gemm.register_fake
def _(A, B, bias, tensor_cores):
    return A.new_empty(A.shape[0], _r4(B.shape[0]))[:, :B.shape[0]]


@torch.library.custom_op(f"{_NS}::score_lse", mutates_args=())
def score_lse(U: torch.Tensor, E: torch.Tensor, d: int) -> torch.Tensor:
    """lse[b] = log sum_i exp(U_b . E_i); U [B, ld], E [N, ld] bf16 (ld % 8 == 0, columns d.. zero)."""
    _cuda(U, E)
    B, N = U.shape[0], E.shape[0]
    lib = _cabi.lib()
    ws = torch.empty(2 * B * int(lib.tc_n_splits(B, N)), dtype=torch.float32, device=U.device)
    lse = torch.empty(B, dtype=torch.float32, device=U.device)
    lib.tc_score_lse(p(U), U.stride(0), B, p(E), E.stride(0), N, d, p(lse), p(ws), _stream(U))
    return lse


@score_lse.register_fake
def _(U, E, d):
    return U.new_empty(U.shape[0], dtype=torch.float32)


@torch.library.custom_op(f"{_NS}::score_grad", mutates_args=())
def score_grad(U: torch.Tensor, E: torch.Tensor, lse: torch.Tensor, d: int) -> torch.Tensor:
    """O [B, d] = softmax(U E^T) E, scores recomputed tile by tile (never stored)."""
    _cuda(U, E, lse)
    B, N = U.shape[0], E.shape[0]
    lib = _cabi.lib()
    ldo = _r4(d)
    parts = torch.empty(int(lib.tc_grad_splits(B, N, d)), B, ldo, dtype=torch.float32, device=U.device)
    lib.tc_score_grad(p(U), U.stride(0), B, p(E), E.stride(0), N, d, p(lse), p(parts), ldo, _stream(U))
    return parts.sum(0)[:, :d]


@score_grad.register_fake
def _(U, E, lse, d):
    return U.new_empty(U.shape[0], d, dtype=torch.float32)


@torch.library.custom_op(f"{_NS}::score_topk", mutates_args=())
def score_topk(U: torch.Tensor, E: torch.Tensor, d: int, indptr: torch.Tensor, indices: torch.Tensor, rows: torch.Tensor,
               K: int, item_offset: int) -> tuple[torch.Tensor, torch.Tensor]:
    """(values [B, K] f32, global item ids [B, K] int32): top-K of U E^T over the items of E with the users' seen items excluded."""
    _cuda(U, E, indptr, indices, rows)
    B, N = U.shape[0], E.shape[0]
    lib = _cabi.lib()
    ns = int(lib.tc_topk_splits(B, N))
    cv = torch.empty(B, ns * K, dtype=torch.float32, device=U.device)
    ci = torch.empty(B, ns * K, dtype=torch.int32, device=U.device)
    st = _stream(U)
    lib.tc_score_topk(p(U), U.stride(0), B, p(E), E.stride(0), N, d, item_offset, p(indptr), p(indices), p(rows), K, p(cv), p(ci), st)
    val = torch.empty(B, K, dtype=torch.float32, device=U.device)
    idx = torch.empty(B, K, dtype=torch.int32, device=U.device)
    lib.topk_merge(p(cv), p(ci), B, ns * K, K, p(val), p(idx), st)
    return val, idx


@score_topk.register_fake
def _(U, E, d, indptr, indices, rows, K, item_offset):
    return U.new_empty(U.shape[0], K, dtype=torch.float32), U.new_empty(U.shape[0], K, dtype=torch.int32)


@torch.library.custom_op(f"{_NS}::adam_step", mutates_args=("params", "exp_avg", "exp_avg_sq"))
def adam_step(params: torch.Tensor, exp_avg: torch.Tensor, exp_avg_sq: torch.Tensor, grad: torch.Tensor, state: torch.Tensor,
              weight_decay: float, beta1: float, beta2: float, eps: float) -> None:
    """In-place Adam over a flat fp32 arena with a dense gradient; `state` is the device hvae_step_state (step size, bias
    correction, clip coefficient: see hvae_step_begin / hvae_grad_norm_clip)."""
    _cuda(params, exp_avg, exp_avg_sq, grad, state)
    _cabi.lib().adam_step(p(params), p(exp_avg), p(exp_avg_sq), params.numel(), 0, 4, None, None, p(grad), p(state), weight_decay,
                          beta1, beta2, eps, _stream(params))
