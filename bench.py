#!/usr/bin/env python
"""bench.py -- path-contexts/sec of one code2vec train step (batch 1024 x 200, java14m shape).

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA engine through the C ABI)
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm: the oracle port of the reference
                                                             # graph on the box's host cores (TensorFlow
                                                             # itself is not installable here)

A "step" is one pass of the hot path over one batch of synthetic path-context bags: three
embedding gathers, tanh(x.W), masked softmax attention, full-softmax logits + loss, the whole
backward pass including the sparse embedding-gradient scatter-add, and the TF1-faithful dense Adam
update (tensorflow_model.py:80 `sess.run([optimizer, train_loss])`).  One JSON line on stdout.

Timing: W >= 3 warm-up steps, then exactly K steps bracketed by barrier + synchronize, CUDA events
on the launching stream, max over ranks.  No L2 flush is needed: each step streams > 9 GB of
parameter / optimizer state through a 126 MB L2 and cycles through distinct input batches.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# BASELINE.json configs[1]: java14m shape (config.py:60-68; vocab + 1 special word)
WORKLOADS = {
    "java14m": dict(token_vocab=1301137, path_vocab=911418, target_vocab=261246, embed_dim=128, code_dim=384,
                    max_contexts=200, batch=1024),
    "tiny": dict(token_vocab=1001, path_vocab=501, target_vocab=1001, embed_dim=32, code_dim=96,
                 max_contexts=20, batch=64),
}
KEEP_PROB = 0.75      # config.py:69 DROPOUT_KEEP_RATE


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            j = json.load(open(p))
            return dict(hbm=float(j["hbm_gbs"]), tensor=float(j["bf16_tflops"]),
                        tensor_sustained=float(j.get("bf16_tflops_sustained", j["bf16_tflops"])), source="measured")
        except Exception:
            pass
    return dict(hbm=6650.0, tensor=1590.0, tensor_sustained=1400.0, source="fallback")


class ClockSampler:
    """Samples nvidia-smi SM clocks / throttle reasons during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0, period_ms=100):
        self.idx, self.period = gpu_index, period_ms
        self.proc, self.lines, self.thread = None, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", str(self.period)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        def pump():
            for ln in self.proc.stdout:
                self.lines.append((time.time(), ln.strip()))
        self.thread = threading.Thread(target=pump, daemon=True)
        self.thread.start()

    def stop(self, t0=None, t1=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        for ts, ln in self.lines:
            if t0 is not None and not (t0 <= ts <= t1 + 0.2):
                continue
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax = float(f[2])
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


def make_batches(w, n_batches, seed, bags="full", zipf=False):
    """bags: "full" = every bag has all MAX_CONTEXTS valid contexts (the HBM worst case the roofline is judged on,
    SURVEY 8d); "normal" = n_b ~ clip(N(120, 60), 1, 200); "ragged" = n_b ~ U{1..C}.  zipf: Zipfian indices."""
    from code2vec_b200.synthetic import synthetic_batch
    return [synthetic_batch(w["token_vocab"], w["path_vocab"], w["target_vocab"], w["max_contexts"], w["batch"],
                            seed=seed + 7919 * i, full_bags=(bags == "full"), normal_bags=(bags == "normal"), zipf=zipf)
            for i in range(n_batches)]


def algorithmic_work(w, B, touched_rows=None, world=1, fused_target_adam=False):
    """Per-step algorithmic FLOPs / bytes of each phase.

    adam: the dense TF1 update streams theta, m, v in and out = 24 B per parameter.  With the
    lazy-but-exact scheme (single GPU) only the rows the batch references are streamed, once: 32 B
    per element (theta, m, v and the deferred gradient, in and out) in the catch-up pass; the dense
    part is the target table (unless fused into dY), TRANSFORM and ATTENTION.  Under table sharding each rank updates 1/world of every table.
    fused_target_adam: the target table's update runs in the dY epilogue, so its 24 B/param move
    from "adam" to "dY", which then is HBM-bound (P^T read once + the update) rather than tensor-bound."""
    d, D, C, Y = w["embed_dim"], w["code_dim"], w["max_contexts"], w["target_vocab"]
    N = B * C
    emb = (w["token_vocab"] + w["path_vocab"]) * d
    rest = Y * D + 3 * d * D + D
    proj = 2.0 * N * 3 * d * D
    logit = 2.0 * B * D * Y
    if fused_target_adam:
        rest -= Y * D
    if touched_rows is not None:
        # lazy Adam: one pass over the batch's rows (theta, m, v, g in and out); the dense kernels keep the rest
        adam, catchup = 24.0 * rest, 32.0 * touched_rows * d
    else:
        adam, catchup = 24.0 * (emb + rest) / world, 0.0
    dy = ("hbm", 24.0 * Y * D + 4.0 * B * Y) if fused_target_adam else ("tensor", logit)
    return {
        "ctx_fwd": ("tensor", proj), "dW": ("tensor", proj), "dx_gemm": ("tensor", proj),
        "logits": ("tensor", logit), "dv": ("tensor", logit), "dY": dy,
        "adam": ("hbm", adam), "adam_catchup": ("hbm", catchup),
        "attn_fwd": ("hbm", 4.0 * N * D), "attn_bwd": ("hbm", 3 * 4.0 * N * D),
        "xent": ("hbm", 2 * 4.0 * B * Y),
        "gather": ("hbm", N * (3 * d * 4 + 16) + 4.0 * N * 3 * d),      # table rows + indices/mask in, X' out
        "dx_scatter": ("hbm", 4.0 * N * 3 * d + 2 * 4.0 * N * 3 * d),   # dX' in, read-modify-write of the table rows
    }


def ncu_traffic():
    """DRAM bytes per launch measured by `ncu --set full` (profiles/r01_traffic.json), keyed by phase."""
    p = os.path.join(ROOT, "profiles", "r01_traffic.json")
    try:
        return json.load(open(p))
    except Exception:
        return {}


# ================================ our arm ========================================================
def run_ours(args):
    import torch
    import torch.distributed as dist
    from code2vec_b200.engine import EngineDims, PathAttentionEngine
    from code2vec_b200.trainer import Trainer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    else:
        torch.cuda.set_device(0)
    dev = torch.device("cuda", local_rank)

    w = dict(WORKLOADS[args.workload])
    if args.batch:
        w["batch"] = args.batch
    B, C = w["batch"], w["max_contexts"]
    gdims = EngineDims(w["token_vocab"], w["path_vocab"], w["target_vocab"], w["embed_dim"], w["code_dim"], C, B, 10)
    if world not in (2, 4, 8) and args.dp_schedule in ("fully_sharded", "table_sharded"):
        args.dp_schedule = "sharded"              # peer-memory table sharding needs a power-of-two world <= 8
    if world > 1 and args.dp_schedule == "fully_sharded":
        from code2vec_b200.trainer import make_fully_sharded_engine
        eng = make_fully_sharded_engine(gdims, B, device=local_rank)
    else:
        eng = PathAttentionEngine(gdims, device=local_rank, training=True)
    eng.init_params(seed=4321)                       # replicated: same seed on every rank
    if args.math == "tf32":
        eng.set_option("math_mode", 1)
    eng.set_option("cta_pair", args.cta_pair)
    trainer = Trainer(eng, keep_prob=KEEP_PROB, seed=99, schedule=args.dp_schedule, fuse_target_adam=not args.no_fuse_adam)

    if args.dy_late >= 0:
        eng.set_option("dy_late", args.dy_late)
    if args.adam_rows_occ:
        eng.set_option("adam_rows_occupancy", args.adam_rows_occ)
    # a training loop knows its next batch (the reader prefetches); --hint passes it on so lazy Adam can run ahead
    nxt = (lambda seq, i: seq[(i + 1) % n_batches]) if args.hint else (lambda seq, i: None)
    n_batches = 4
    host = make_batches(w, n_batches, seed=1234 + 100003 * rank, bags=args.bags, zipf=args.zipf)
    pinned = [[torch.from_numpy(a).pin_memory() for a in b] for b in host]
    i32, f32 = torch.int32, torch.float32
    devb = [[b[0].to(dev), b[1].to(dev), b[2].to(dev), b[3].to(dev), b[4].to(dev)] for b in pinned]

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    # ---- device-resident timing (value) ---------------------------------------------------
    W, K = max(args.warmup, 3), args.steps
    for i in range(W):
        trainer.step_device(*devb[i % n_batches], next_batch=nxt(devb, i))
    sync_all()
    eng.set_option("profile", 1)
    eng.phase_stats(reset=True)
    clocks = ClockSampler(local_rank, period_ms=20)
    if rank == 0:
        clocks.start()
        time.sleep(0.3)
    sync_all()
    launches0 = eng.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.time()
    ev0.record()
    for i in range(K):
        loss_dev = trainer.step_device(*devb[i % n_batches], next_batch=nxt(devb, i))
    ev1.record()
    sync_all()
    t_wall1 = time.time()
    ms = ev0.elapsed_time(ev1)
    launches = eng.launch_count - launches0
    phases = eng.phase_stats(reset=True)
    eng.set_option("profile", 0)
    clk = clocks.stop(t_wall0, t_wall1) if rank == 0 else None
    last_loss = float(loss_dev.cpu()[0])

    # ---- end to end through the host-buffer API (e2e) ---------------------------------------
    for i in range(2):
        trainer.step_host(*pinned[i % n_batches], next_batch=nxt(pinned, i))
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        trainer.step_host(*pinned[i % n_batches], next_batch=nxt(pinned, i))
    e1.record()
    sync_all()
    ms_e2e = e0.elapsed_time(e1)

    if world > 1:
        t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(t[0]), float(t[1])

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    contexts = float(world) * B * C * K
    value = contexts / (ms * 1e-3)
    e2e_value = contexts / (ms_e2e * 1e-3)
    peaks = _peaks()
    touched = None
    if trainer.schedule == "single" and eng.get_option("lazy_adam"):
        touched = float(np.mean([len(np.unique(np.concatenate([b[0].ravel(), b[2].ravel()]))) + len(np.unique(b[1]))
                                 for b in host]))
    fused = bool(getattr(trainer, "fuse_tgt", False)) and args.math == "tf32"
    work = algorithmic_work(w, B, touched_rows=touched, world=world if trainer.schedule == "table_sharded" else 1,
                            fused_target_adam=fused and world == 1)
    traffic = ncu_traffic()
    phase_out = {}
    dominant, dom_ms = None, -1.0
    for name, (tot_ms, n) in phases.items():
        avg = tot_ms / K                      # per step (a phase may be several launches)
        entry = {"ms": round(avg, 4), "share": round(tot_ms / ms, 4)}
        if name in work:
            kind, amount = work[name]
            if kind == "tensor":
                entry["tflops"] = round(amount / (avg * 1e-3) / 1e12, 2)
            else:
                entry["gbs"] = round(amount / (avg * 1e-3) / 1e9, 1)
        phase_out[name] = entry
        if avg > dom_ms and name in work:
            dominant, dom_ms = name, avg
    roofline = None
    if dominant:
        kind, amount = work[dominant]
        if kind == "tensor":
            ach = amount / (dom_ms * 1e-3) / 1e12
            peak = peaks["tensor_sustained"]
            roofline = {"kernel": dominant, "bound": "tensor", "achieved": round(ach, 2), "peak": peak, "unit": "TFLOP/s",
                        "frac": round(ach / peak, 4), "traffic": traffic.get(dominant),
                        "peak_source": peaks["source"] + " bf16 dense (sustained); tf32 tcgen05 peak is half of it"}
        else:
            ach = amount / (dom_ms * 1e-3) / 1e9
            peak = peaks["hbm"]
            roofline = {"kernel": dominant, "bound": "hbm", "achieved": round(ach, 1), "peak": peak, "unit": "GB/s",
                        "frac": round(ach / peak, 4), "traffic": traffic.get(dominant),
                        "peak_source": peaks["source"] + " copy bandwidth"}

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(w, host[0], steps=1, threads=args.cpu_threads or None)

    h2d = sum(int(a.nbytes) for a in host[0])
    if world == 1 and args.hint and fused:
        h2d += sum(int(a.nbytes) for a in host[0][:3])      # the next batch's index arrays are copied once more as the hint
    out = {
        "metric": "path-contexts/sec (train step, batch 1024x200)", "value": round(value, 1), "unit": "path-contexts/s",
        "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": round(ms / K, 4), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None,
        "dtype": "tf32 operands / fp32 accumulate+storage" if args.math == "tf32" else "f32",
        "data": "synthetic",
        "config": {"workload": "java14m-shape train step: T=1301137 P=911418 Y=261246 d=128 D=384 C=200, "
                               "full softmax, dropout keep 0.75, TF1 dense Adam" if args.workload == "java14m" else args.workload,
                   "batch_per_gpu": B, "global_batch": B * world, "contexts_per_example": C,
                   "parallelism": "dp%d (%s)" % (world, trainer.schedule) if world > 1 else "single",
                   "l2": "no flush: >9 GB of parameter/optimizer traffic per step and 4 rotating input batches exceed the 126 MB L2",
                   "math_mode": args.math, "fused_target_adam": fused, "next_batch_hint": bool(world == 1 and args.hint and fused), "last_loss": round(last_loss, 5),
                   "inputs": "%s bags, %s indices; all %d slots per example are counted in the metric" % (
                       args.bags, "zipf(1.2)" if args.zipf else "uniform", C),
                   "valid_context_fraction": round(float(np.mean([b[3].mean() for b in host])), 4)},
        "e2e": {"value": round(e2e_value, 1), "unit": "path-contexts/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": 4, "ms_per_step": round(ms_e2e / K, 4)},
        "gpu_launches": int(launches),
        "clocks": clk,
        "roofline": roofline,
        "phases": phase_out,
        "cpu_baseline": cpu,
    }
    print(json.dumps(out))
    sys.stdout.flush()
    if world > 1:
        dist.destroy_process_group()


# ================================ CPU arm ========================================================
def default_cpu_threads():
    """Upper bound of the CPU arm's intra-op threads: every host core unless C2V_CPU_THREADS says otherwise."""
    return int(os.environ.get("C2V_CPU_THREADS", "0")) or os.cpu_count() or 1


def pick_cpu_threads(tr, batch, limit):
    """More threads are not always faster (two sockets x SMT: 128 threads ran the step at 9.0 s, 32 at 3.2 s on the
    GPU boxes' Xeons), so the CPU arm gets the best of {all, 1/2, 1/4, 1/8} of the host threads, probed with one
    step on a 64-example slice each (the dense Adam over all 383 M parameters is part of every probe)."""
    import torch
    src, pth, tgt, mask, target = (a[:64] for a in batch)
    cands = sorted({max(1, limit // d) for d in (1, 2, 4, 8)}, reverse=True)
    if limit <= 16 or len(cands) == 1:          # small hosts: every core helps, nothing to probe
        torch.set_num_threads(limit)
        return limit, {}
    timing = {}
    for c in cands:
        torch.set_num_threads(c)
        t0 = time.time()
        tr.train_step(src, pth, tgt, mask, target, keep=1.0)
        timing[c] = round(time.time() - t0, 3)
    best = min(timing, key=timing.get)
    torch.set_num_threads(best)
    return best, timing


def cpu_baseline(w, batch, steps=1, threads=None):
    """The oracle port of the reference graph (torch-CPU, all host threads): forward + backward +
    TF1 dense Adam on the same workload; `steps` full batches (bounded sample)."""
    import torch
    from oracle.path_attention_oracle import Dims, init_params
    from oracle.torch_crosscheck import TorchCpuTrainer
    cores = threads or default_cpu_threads()
    dims = Dims(w["token_vocab"], w["path_vocab"], w["target_vocab"], w["embed_dim"], w["code_dim"], w["max_contexts"])
    params = init_params(dims, seed=4321)
    tr = TorchCpuTrainer(params, threads=cores)
    probed = {}
    if not threads:
        _, probed = pick_cpu_threads(tr, batch, cores)
    src, pth, tgt, mask, target = batch
    rng = np.random.default_rng(0)
    times = []
    for s in range(steps):
        dm = (rng.random((src.shape[0] * src.shape[1], 3 * w["embed_dim"]), dtype=np.float32) < KEEP_PROB).astype(np.float32)
        t0 = time.time()
        tr.train_step(src, pth, tgt, mask, target, keep=KEEP_PROB, dropout_mask=dm)
        times.append(time.time() - t0)
    sec = float(np.mean(times))
    B, C = src.shape
    return {"value": round(B * C / sec, 1), "unit": "path-contexts/s", "cores": int(torch.get_num_threads()), "kind": "port",
            "sample": "%d full train step(s) of the same workload (B=%d x C=%d), torch-CPU restatement of "
                      "tensorflow_model.py:197-265 incl. dense Adam; %.2f s/step%s" % (
                          steps, B, C, sec, "; threads = fastest of a 64-example probe %s (s)" % probed if probed else "")}


def run_reference(args):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = dict(WORKLOADS[args.workload])
    if args.batch:
        w["batch"] = args.batch
    K, W = args.steps, args.warmup
    # bounded: each CPU step costs seconds; cap the total number of full-batch steps
    k_eff = max(1, min(K, 3))
    batch = make_batches(w, 1, seed=1234, bags=args.bags, zipf=args.zipf)[0]
    import torch
    from oracle.path_attention_oracle import Dims, init_params
    from oracle.torch_crosscheck import TorchCpuTrainer
    cores = args.cpu_threads or default_cpu_threads()
    dims = Dims(w["token_vocab"], w["path_vocab"], w["target_vocab"], w["embed_dim"], w["code_dim"], w["max_contexts"])
    tr = TorchCpuTrainer(init_params(dims, seed=4321), threads=cores)
    probed = {}
    if not args.cpu_threads:
        _, probed = pick_cpu_threads(tr, batch, cores)
    src, pth, tgt, mask, target = batch
    rng = np.random.default_rng(0)
    def one():
        dm = (rng.random((src.shape[0] * src.shape[1], 3 * w["embed_dim"]), dtype=np.float32) < KEEP_PROB).astype(np.float32)
        t0 = time.time()
        tr.train_step(src, pth, tgt, mask, target, keep=KEEP_PROB, dropout_mask=dm)
        return time.time() - t0
    for _ in range(min(W, 1)):
        one()
    times = [one() for _ in range(k_eff)]
    sec = float(np.mean(times))
    B, C = src.shape
    value = B * C / sec
    sample = ("%d timed full-batch step(s) (of --steps %d; bounded) after %d warm-up, torch-CPU restatement of the "
              "reference graph (TensorFlow not installable here), %d threads%s" % (
                  k_eff, K, min(W, 1), torch.get_num_threads(), " (fastest of a 64-example probe %s s)" % probed if probed else ""))
    out = {"impl": "reference", "metric": "path-contexts/sec (train step, batch 1024x200)", "value": round(value, 1),
           "unit": "path-contexts/s", "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": round(sec * 1e3, 2),
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": "java14m-shape train step: T=1301137 P=911418 Y=261246 d=128 D=384 C=200, full softmax, "
                                  "dropout keep 0.75, TF1 dense Adam" if args.workload == "java14m" else args.workload,
                      "batch_per_gpu": B, "global_batch": B, "contexts_per_example": C, "parallelism": "cpu"},
           "cpu_baseline": {"value": round(value, 1), "unit": "path-contexts/s", "cores": int(torch.get_num_threads()),
                            "kind": "port", "sample": sample},
           "e2e": {"value": round(value, 1), "unit": "path-contexts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="java14m", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--math", default=os.environ.get("C2V_MATH", "tf32"), choices=["fp32", "tf32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-threads", type=int, default=0, help="threads of the CPU arm (0 = all host cores)")
    ap.add_argument("--dy-late", type=int, default=-1, choices=[-1, 0, 1, 2],
                    help="engine option dy_late (-1 = the schedule's default)")
    ap.add_argument("--adam-rows-occ", type=int, default=0, choices=[0, 4, 5], help="engine option adam_rows_occupancy (0 = default)")
    ap.add_argument("--hint", action="store_true",
                    help="hint the next batch to the engine (c2v_hint_next_batch); measured no gain on one GPU, off by default")
    ap.add_argument("--no-fuse-adam", action="store_true",
                    help="keep the target table's Adam update as a separate pass instead of the dY epilogue")
    ap.add_argument("--bags", default="full", choices=["full", "normal", "ragged"],
                    help="valid contexts per bag: full (default; worst case for HBM), normal ~N(120,60), ragged ~U{1..C}")
    ap.add_argument("--zipf", action="store_true", help="Zipfian instead of uniform indices (hot rows, L2 reuse)")
    ap.add_argument("--cta-pair", type=int, default=int(os.environ.get("C2V_CTA_PAIR", "2")),
                    help="tcgen05 GEMMs as CTA pairs (cta_group::2): 0 never, 1 always, 2 auto (default)")
    ap.add_argument("--dp-schedule", default=os.environ.get("C2V_DP_SCHEDULE", "fully_sharded"),
                    choices=["fully_sharded", "table_sharded", "sharded", "allreduce"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
