#!/usr/bin/env python
"""Step-time distribution of the train step on a LONG stream of DISTINCT power-law (Zipf-like) batches.

bench.py cycles through 16 batches of uniform indices: every row it uses is at most 16 steps stale.  Real data is
Zipfian: a few rows are used every step, most sit idle for thousands of steps -- the case in which a lazily deferred
Adam update could pile up replay work.  This tool draws `--steps` different batches on the device
(indices ~ floor(U^(-1/(a-1))), a = 1.2, folded into the vocabulary; bag lengths ~ N(120, 60)), runs them once each and
reports p50 / p99 / max of the per-step device time (one CUDA-event pair per step) for
    lazy  : lazy Adam with the sweep (adam_sweep_period = 32) -- the default of Trainer("single")
    nosweep: lazy Adam, sweep off (rows are only replayed when a batch references them, or at the final flush)
    dense : dense Adam over both embedding tables every step
plus the time of c2v_sync_tables after the run (what a checkpoint pays).  One JSON line per configuration."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

W = dict(token_vocab=1301137, path_vocab=911418, target_vocab=261246, embed_dim=128, code_dim=384, max_contexts=200, batch=1024)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=512)
    ap.add_argument("--warmup", type=int, default=96)
    ap.add_argument("--configs", default="lazy,nosweep,dense")
    ap.add_argument("--alpha", type=float, default=1.2)
    args = ap.parse_args()
    import torch
    from code2vec_b200.engine import EngineDims, PathAttentionEngine
    from code2vec_b200.trainer import Trainer
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    B, C = W["batch"], W["max_contexts"]
    n_total = args.steps + args.warmup
    g = torch.Generator(device=dev)
    g.manual_seed(1234)

    def power_law(vocab, shape):
        u = torch.rand(shape, device=dev, generator=g).clamp_(min=1e-12)
        x = torch.floor(u.pow(-1.0 / (args.alpha - 1.0))).clamp_(max=float(2 ** 40))
        return (1 + (x.to(torch.int64) - 1) % (vocab - 1)).to(torch.int32)

    batches = []
    for _ in range(n_total):
        n_valid = torch.clamp(torch.round(torch.randn(B, device=dev, generator=g) * 0.3 * C + 0.6 * C), 1, C)
        valid = torch.arange(C, device=dev)[None, :] < n_valid[:, None]
        src = torch.where(valid, power_law(W["token_vocab"], (B, C)), 0).to(torch.int32)
        pth = torch.where(valid, power_law(W["path_vocab"], (B, C)), 0).to(torch.int32)
        tgt = torch.where(valid, power_law(W["token_vocab"], (B, C)), 0).to(torch.int32)
        target = torch.randint(1, W["target_vocab"], (B,), device=dev, generator=g, dtype=torch.int32)
        batches.append((src.contiguous(), pth.contiguous(), tgt.contiguous(), valid.to(torch.float32).contiguous(), target))
    uniq = float(np.mean([len(torch.unique(torch.cat([b[0].ravel(), b[2].ravel()]))) + len(torch.unique(b[1])) for b in batches[:8]]))
    for name in args.configs.split(","):
        eng = PathAttentionEngine(EngineDims(W["token_vocab"], W["path_vocab"], W["target_vocab"], W["embed_dim"], W["code_dim"],
                                             C, B, 10), device=0, training=True)
        eng.init_params(seed=4321)
        eng.set_option("math_mode", 1)
        tr = Trainer(eng, keep_prob=0.75, seed=99, lazy_adam=(name != "dense"))
        if name == "nosweep":
            eng.set_option("adam_sweep_period", 0)
        for i in range(args.warmup):
            tr.step_device(*batches[i])
        torch.cuda.synchronize()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        for i in range(args.steps):
            evs[i][0].record()
            tr.step_device(*batches[args.warmup + i])
            evs[i][1].record()
        torch.cuda.synchronize()
        ms = np.array([a.elapsed_time(b) for a, b in evs])
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        eng.sync_tables()
        s1.record()
        torch.cuda.synchronize()
        print(json.dumps({"config": name, "steps": args.steps, "warmup": args.warmup, "distinct_batches": n_total,
                          "index_distribution": "power law, exponent %.2f" % args.alpha, "bags": "n ~ N(120, 60)",
                          "unique_rows_per_batch": round(uniq, 1),
                          "ms_p50": round(float(np.percentile(ms, 50)), 4), "ms_p99": round(float(np.percentile(ms, 99)), 4),
                          "ms_max": round(float(ms.max()), 4), "ms_mean": round(float(ms.mean()), 4),
                          "p99_over_p50": round(float(np.percentile(ms, 99) / np.percentile(ms, 50)), 3),
                          "sync_tables_ms": round(s0.elapsed_time(s1), 3),
                          "adam_sweep_period": int(eng.get_option("adam_sweep_period")) if name != "dense" else None,
                          "path_contexts_per_s": round(B * C / (ms.mean() * 1e-3), 1)}))
        sys.stdout.flush()
        eng.close()
        del eng, tr
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
