#!/usr/bin/env python
"""Benchmark of the HybridVAE hot path (BASELINE.json metric: train users/sec & eval top-K users/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2] [--precision bf16|fp32]
    python bench.py --impl reference ...      # the reference's CPU path (oracle port) on the host cores

A "step" is one optimisation step (forward, multinomial-NLL + KL, backward, clip, Adam) over one batch of
synthetic users of the named workload (hvae_b200.synth.CONFIGS, shapes from BASELINE.json `configs`).
Prints ONE JSON line (rank 0).  See DESIGN.md §Measurement for how each field is obtained.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
for _p in (str(ROOT), str(ROOT / "recommendation-system_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC, UNIT = "train_users_per_sec", "users/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=["c1", "c2", "c3"])
    ap.add_argument("--precision", default=os.environ.get("HVAE_B200_PRECISION", "bf16"), choices=["bf16", "fp32"])
    ap.add_argument("--batch", type=int, default=0, help="users per GPU per step (default: the workload's)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the bounded CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-eval", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    return ap.parse_args()


def peaks():
    f = ROOT / "MEASURED_PEAKS.json"
    if f.exists():
        d = json.loads(f.read_text())
        return dict(hbm=float(d["hbm_gbs"]), tc_burst=float(d["bf16_tflops"]), tc_sustained=float(d["bf16_tflops_sustained"]),
                    source="measured")
    return dict(hbm=6650.0, tc_burst=1590.0, tc_sustained=1400.0, source="fallback")


# ------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the timed region runs (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.samples, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.perf_counter(), line.strip()))

    def stop(self, t0=None, t1=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        time.sleep(0.06)
        self.proc.terminate()
        rows = [s for s in self.samples if t0 is None or t0 - 0.05 <= s[0] <= t1 + 0.05] or self.samples
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for _, line in rows:
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        med = float(np.median(sm)) if sm else None
        return {"sm_mhz": med, "sm_max_mhz": max(mx) if mx else None, "power_w_max": max(pw) if pw else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------------------
def make_workload(name, batch_override=0):
    from hvae_b200.synth import CONFIGS, make_interactions, make_item_embeddings
    c = dict(CONFIGS[name])
    if batch_override:
        c["batch"] = batch_override
    data = make_interactions(c["n_users"], c["n_items"], 0)
    E = make_item_embeddings(c["n_items"], c["emb_dim"], 0)
    return c, data, E


def cpu_reference_run(c, data, E, steps, warmup, seconds=None):
    """The reference's train_epoch arithmetic on the host cores (oracle port; src/ml/train.py:81-103): DataLoader row
    densification, dense fp32 forward/backward through ATen, clip, Adam.  Returns users/s over `steps` steps (or as many
    as fit in `seconds`), after `warmup` untimed ones."""
    from oracle import hvae_oracle as orc
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(0)
    m = orc.OracleVAE(c["n_items"], E, c["latent_dim"], c["hidden_dims"], c["dropout"], c["beta"])
    csr = data.scipy_csr()
    B = c["batch"]
    rng = np.random.default_rng(0)
    order = rng.permutation(c["n_users"])
    opt = orc.make_adam(m)
    m.train()
    done, t_used, s = 0, 0.0, 0
    total_steps = warmup + steps
    while s < total_steps:
        rows = order[(s * B) % c["n_users"]:][:B]
        if len(rows) < B:
            rows = order[:B]
        t0 = time.perf_counter()
        x = torch.stack([torch.FloatTensor(csr[int(r)].toarray().flatten()) for r in rows])   # train.py:45-47 + default collate
        opt.zero_grad()
        sc, mu, lv = m.forward_ref(x)
        loss, _, _ = orc.loss_terms(sc, x, mu, lv, c["beta"])
        loss.backward()
        torch.nn.utils.clip_grad_norm_(m.parameters(), max_norm=5.0)
        opt.step()
        loss.item()
        dt = time.perf_counter() - t0
        if s >= warmup:
            done += len(rows)
            t_used += dt
            if seconds is not None and t_used >= seconds and s - warmup + 1 >= 4:
                s += 1
                break
        s += 1
    n_steps = s - warmup
    return done / t_used, n_steps, t_used, torch.get_num_threads()


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    c, data, E = make_workload(args.workload, args.batch)
    steps = min(args.steps, 60)          # each CPU step is ~0.1-0.2 s at C2; keep the run within a few minutes
    ups, n_steps, secs, cores = cpu_reference_run(c, data, E, steps, min(args.warmup, 3))
    line = {"impl": "reference", "metric": METRIC, "value": ups, "unit": UNIT, "n_gpus": args.gpus, "steps": n_steps,
            "warmup": min(args.warmup, 3), "ms_per_step": 1e3 * secs / n_steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.workload, c, 1, "cpu"),
            "cpu_baseline": {"value": ups, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{n_steps} steps of batch {c['batch']} ({n_steps * c['batch']} users), "
                                       "oracle/hvae_oracle.py train loop incl. row densification"},
            "e2e": {"value": ups, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_config(name, c, n_gpus, precision):
    desc = {"c1": "Appliances-shaped synthetic", "c2": "All_Beauty-shaped synthetic", "c3": "1M users x 200k items synthetic"}[name]
    return {"workload": f"{name}: {desc} ({c['n_users']} users x {c['n_items']} items, d={c['emb_dim']}, latent {c['latent_dim']}, "
                        f"hidden {c['hidden_dims']}), HybridVAE training step",
            "users_per_gpu_per_step": c["batch"], "global_batch": c["batch"] * n_gpus, "precision": precision,
            "parallelism": f"dp{n_gpus}" if n_gpus > 1 else "single",
            "l2": "L2 flushed (256 MiB write) between timed steps, untimed"}


# ------------------------------------------------------------------------------------------------------------------
def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")

    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from hvae_b200 import _cabi
    from hvae_b200.engine import Batch, DeviceCSR
    from hvae_b200.evaluate import RecommendationEvaluator
    from hvae_b200.model import HybridVAE
    from hvae_b200.train import VAETrainer

    pk = peaks()
    c, data, E = make_workload(args.workload, args.batch)
    B, U, N, d, L, h = c["batch"], c["n_users"], c["n_items"], c["emb_dim"], c["latent_dim"], c["hidden_dims"]
    torch.manual_seed(0)
    model = HybridVAE(N, E, L, h, c["dropout"], c["beta"], precision=args.precision)
    trainer = VAETrainer(model, dev, lr=1e-3, use_cuda_graph=not args.no_graph)
    dp = trainer.enable_data_parallel() if world > 1 else None
    eng = model.engine
    lib = _cabi.lib()
    csr = DeviceCSR.from_arrays(data.indptr, data.indices, None, N, dev)

    # global batches: a fixed permutation of the users, walked cyclically; rank r takes slice r of each global batch
    Bg = B * world
    order = np.random.default_rng(0).permutation(U).astype(np.int32)
    order = np.concatenate([order, order[:Bg]])
    order_dev = torch.from_numpy(order).to(dev)
    lens = np.diff(data.indptr)
    n_batches = U // Bg if U >= Bg else 1

    cap_local = max(int(lens[order[i * Bg + rank * B:i * Bg + (rank + 1) * B]].sum()) for i in range(n_batches))
    cap_global = max(int(lens[order[i * Bg:(i + 1) * Bg]].sum()) for i in range(n_batches))

    def batch_at(s):
        g0 = (s % n_batches) * Bg
        return Batch(csr, order_dev[g0 + rank * B:g0 + (rank + 1) * B], B, max(1, cap_local), b_global=Bg, nnz_cap_global=cap_global)

    flush_buf = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    model.train()

    def step(s):
        trainer.train_step(batch_at(s), b_global=Bg)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    W, K = max(args.warmup, 3), args.steps
    clocks = ClockSampler(local) if rank == 0 else None
    for s in range(W):
        step(s)
    barrier()
    # ---- timed region 1 (the `value`): K steps, inputs resident in HBM, per-step CUDA events, L2 flushed between steps
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    l0 = lib.launches
    t_wall0 = time.perf_counter()
    for s in range(K):
        flush_buf.fill_(s & 0xFF)
        evs[s][0].record()
        step(W + s)
        evs[s][1].record()
    barrier()
    t_wall1 = time.perf_counter()
    launches = lib.launches - l0
    ms_total = sum(a.elapsed_time(b) for a, b in evs)
    # ---- back-to-back (hot L2) variant, one event pair around all K steps
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for s in range(K):
        step(W + K + s)
    e1.record()
    barrier()
    ms_hot = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms_total, ms_hot], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total, ms_hot = float(t[0]), float(t[1])
    value = K * Bg / (ms_total * 1e-3)
    clk = clocks.stop(t_wall0, t_wall1) if clocks else None

    # ---- per-kernel-group durations (CUDA events on the launching stream, same steps, un-graphed) -> roofline
    trainer_graph = trainer.use_cuda_graph
    trainer.use_cuda_graph = False
    eng.prof = {}
    PS = min(K, 50)
    for s in range(PS):
        flush_buf.fill_(1)
        step(W + 2 * K + s)
    spans = eng.span_ms()
    eng.prof = None
    trainer.use_cuda_graph = trainer_graph
    span_avg = {k: float(np.mean(v)) for k, v in spans.items()}
    lay = model.layout
    n_unique_avg = float(np.mean([len(np.unique(np.concatenate([data.indices[data.indptr[u]:data.indptr[u + 1]]
                                                               for u in order[i * Bg:(i + 1) * Bg]]))) for i in range(min(4, n_batches))]))
    ld1 = (h[0] + 3) // 4 * 4
    adam_bytes = 24.0 * lay.n_params + 4.0 * (lay.n_dense + n_unique_avg * ld1)
    flops_fwd = 2.0 * B * N * d
    kernels = {}
    if "adam" in span_avg:
        kernels["adam"] = {"ms": span_avg["adam"], "bound": "hbm", "achieved": adam_bytes / (span_avg["adam"] * 1e-3) / 1e9, "peak": pk["hbm"],
                           "unit": "GB/s", "algorithmic_bytes": adam_bytes}
    if "score_fwd" in span_avg:
        kernels["score_fwd"] = {"ms": span_avg["score_fwd"], "bound": "tensor", "achieved": flops_fwd / (span_avg["score_fwd"] * 1e-3) / 1e12,
                                "peak": pk["tc_sustained"], "unit": "TFLOP/s", "algorithmic_flops": flops_fwd}
    if "score_bwd" in span_avg:
        kernels["score_bwd"] = {"ms": span_avg["score_bwd"], "bound": "tensor", "achieved": flops_fwd / (span_avg["score_bwd"] * 1e-3) / 1e12,
                                "peak": pk["tc_sustained"], "unit": "TFLOP/s", "algorithmic_flops": flops_fwd,
                                "executed_flops": flops_fwd * (1 + -(-((d + 63) // 64 * 64) // 384))}
    if "score_onepass" in span_avg:     # forward + backward through the scores in one sweep: S = U E^T and O = P E, 2BNd each
        kernels["score_onepass"] = {"ms": span_avg["score_onepass"], "bound": "tensor",
                                    "achieved": 2 * flops_fwd / (span_avg["score_onepass"] * 1e-3) / 1e12, "peak": pk["tc_sustained"],
                                    "unit": "TFLOP/s", "algorithmic_flops": 2 * flops_fwd,
                                    "executed_flops": flops_fwd * (2 if d <= 768 else 1 + -(-((d + 63) // 64 * 64) // 384)),
                                    "launches": "one-pass scoring kernel + combine"}
    if "gather" in span_avg:
        nnz_avg = float(lens[order[:Bg * min(4, n_batches)]].sum()) / min(4, n_batches) / world
        gbytes = nnz_avg * (ld1 * 4 + 4) + B * (3 * ld1 * 4 + 16)
        kernels["gather"] = {"ms": span_avg["gather"], "bound": "hbm", "achieved": gbytes / (span_avg["gather"] * 1e-3) / 1e9, "peak": pk["hbm"],
                             "unit": "GB/s", "algorithmic_bytes": gbytes}
    for k in kernels.values():
        k["frac"] = k["achieved"] / k["peak"]
    dom = max(kernels, key=lambda k: kernels[k]["ms"]) if kernels else None
    traffic = None
    tf = ROOT / "profiles" / "traffic.json"
    if tf.exists() and dom:
        traffic = json.loads(tf.read_text()).get(f"{args.workload}:{args.precision}:{dom}")
    roofline = None
    if dom:
        kd = kernels[dom]
        roofline = {"kernel": dom, "bound": kd["bound"], "achieved": kd["achieved"], "peak": kd["peak"], "unit": kd["unit"],
                    "frac": kd["frac"], "traffic": traffic, "peak_source": pk["source"] + (" sustained" if kd["bound"] == "tensor" else ""),
                    "ms_per_launch": kd["ms"],
                    "share_of_step": kd["ms"] / max(1e-9, sum(v for k, v in span_avg.items() if not k.startswith("score_")))}

    # ---- timed region 2 (`e2e`): public API with HOST (pinned) batches; H2D of the batch and D2H of the loss every step
    KE = min(K, 200)
    host_batches = []
    if world == 1:      # the step's input is the batch's CSR slice (indptr, indices, values) in pinned host memory
        for s in range(KE):
            g0 = (s % n_batches) * Bg
            rows_h = order[g0 + rank * B:g0 + (rank + 1) * B].astype(np.int64)
            starts, ln = data.indptr[rows_h], lens[rows_h]
            crow = np.zeros(B + 1, dtype=np.int64)
            np.cumsum(ln, out=crow[1:])
            take = np.repeat(starts - crow[:-1], ln) + np.arange(int(ln.sum()), dtype=np.int64)
            col = data.indices[take].astype(np.int32)
            x = torch.sparse_csr_tensor(torch.from_numpy(crow).pin_memory(), torch.from_numpy(col).pin_memory(),
                                        torch.ones(col.shape[0], dtype=torch.float32).pin_memory(), size=(B, N), check_invariants=False)
            host_batches.append(x)
        h2d = float(np.mean([x.crow_indices().numel() * 8 + x.col_indices().numel() * 4 + x.values().numel() * 4 for x in host_batches]))
        e2e_api = "VAETrainer.train_on_batch(pinned host CSR batch) -> loss floats"
    else:               # data parallel: the interaction CSR is resident on every rank; the step's input is its user ids
        trainer.set_interactions(csr)
        for s in range(KE):
            g0 = (s % n_batches) * Bg
            host_batches.append(torch.from_numpy(order[g0 + rank * B:g0 + (rank + 1) * B].copy()).pin_memory())
        h2d = float(B * 4)
        e2e_api = "VAETrainer.train_on_batch(pinned host user-id batch into the resident CSR) -> loss floats"
    capg = cap_global if world > 1 else None
    for s in range(3):
        trainer.train_on_batch(host_batches[s], b_global=Bg, nnz_cap_global=capg)
    barrier()
    t_e2e = 0.0
    for s in range(KE):
        flush_buf.fill_(s & 0xFF)
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        trainer.train_on_batch(host_batches[s], b_global=Bg, nnz_cap_global=capg)     # returns python floats -> synchronises
        t_e2e += time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([t_e2e], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_e2e = float(t[0])
    e2e = {"value": KE * Bg / t_e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 12, "steps": KE,
           "api": e2e_api}

    # ---- evaluation: full-ranking top-K + Recall/NDCG/HR over every user (evaluate.py:243-265), user-sharded over ranks
    ev_out = None
    if not args.no_eval:
        model.eval()
        ev = RecommendationEvaluator(model, csr, {}, {}, dev, batch_users=4096)
        from hvae_b200.dist import split_even
        lo, hi = split_even(U, world, rank)
        users = np.arange(lo, hi, dtype=np.int64)
        rel_ptr = np.arange(len(users) + 1, dtype=np.int64)
        rel_idx = data.test_items[users].astype(np.int32)
        ev.evaluate_users(users[:min(len(users), 4096)], rel_ptr[:min(len(users), 4096) + 1], rel_idx[:min(len(users), 4096)], [5, 10, 20])
        barrier()
        reps = 3
        t0 = time.perf_counter()
        for _ in range(reps):
            res, _ = ev.evaluate_users(users, rel_ptr, rel_idx, [5, 10, 20])    # host ids in, metric sums out (sync)
        barrier()
        t_ev = (time.perf_counter() - t0) / reps
        if world > 1:
            t = torch.tensor([t_ev], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            t_ev = float(t[0])
        ev_out = {"metric": "eval_topk_users_per_sec", "value": U / t_ev, "unit": UNIT, "k_values": [5, 10, 20], "users": U,
                  "ndcg@10": res[10]["ndcg"], "timing": "wall clock incl. H2D of user ids and D2H of metric sums"}
        model.train()

    if rank != 0:
        _finish(world, dist, dev)
        return

    cpu = None
    if not args.no_cpu_baseline:
        ups, n_steps, secs, cores = cpu_reference_run(c, data, E, 10 ** 6, 2, seconds=args.cpu_seconds)
        cpu = {"value": ups, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{n_steps} steps of batch {B} ({n_steps * B} users, {secs:.1f} s) of the same workload through "
                         "oracle/hvae_oracle.py (reference train loop incl. per-row densification)"}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_total / K,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
            "config": workload_config(args.workload, c, world, args.precision),
            "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu, "clocks": clk,
            "value_hot_l2": K * Bg / (ms_hot * 1e-3), "ms_per_step_hot_l2": ms_hot / K,
            "kernels": kernels, "spans_ms": span_avg, "eval": ev_out, "cuda_graph": bool(trainer.use_cuda_graph)}
    print(json.dumps(line), flush=True)
    _finish(world, dist, dev)


def _finish(world, dist, dev):
    """Multi-rank exit: the captured step graphs hold NCCL work, so synchronise and leave without tearing the communicator down
    (process exit releases it; destroy_process_group() can wait forever on graph-captured collectives)."""
    if world > 1:
        torch.cuda.synchronize(dev)
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
