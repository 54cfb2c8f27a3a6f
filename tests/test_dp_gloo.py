"""Host-side logic of the data-parallel path (BASELINE config 4) on CPU with gloo, world size 2:
the gradient all-reduce(mean) helper the trainer uses between c2v_train_step and c2v_adam_step,
batch sharding, and the identity it relies on -- averaging the per-shard mean-loss gradients of
equal shards equals the gradient of the mean loss over the global batch (checked with the oracle)."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from code2vec_b200.trainer import allreduce_mean_, shard_bounds
from oracle import path_attention_oracle as O

DIMS = O.Dims(token_vocab=61, path_vocab=37, target_vocab=53, embed_dim=8, code_dim=24, max_contexts=6)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        params = O.init_params(DIMS, seed=3)
        B = 8
        src, pth, tgt, mask, target = O.synthetic_batch(DIMS, B, seed=9)
        lo, hi = shard_bounds(B, rank, world)
        _, g_local, _ = O.train_loss_and_grads(params, src[lo:hi], pth[lo:hi], tgt[lo:hi], mask[lo:hi], target[lo:hi])
        tensors = [torch.from_numpy(g_local[k].copy()) for k in O.PARAM_NAMES]
        allreduce_mean_(tensors)
        q.put((rank, [t.numpy() for t in tensors]))
    finally:
        dist.destroy_process_group()


def test_shard_bounds_cover_batch():
    for n in (1, 7, 8, 1024, 8192):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_allreduce_mean_is_noop_without_process_group():
    t = torch.ones(3)
    allreduce_mean_([t])
    assert torch.equal(t, torch.ones(3))


@pytest.mark.timeout(120)
def test_gloo_two_ranks_average_equals_global_batch_gradient():
    world, port = 2, 29000 + (os.getpid() % 2000)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=90) for _ in range(world))
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    params = O.init_params(DIMS, seed=3)
    src, pth, tgt, mask, target = O.synthetic_batch(DIMS, 8, seed=9)
    _, g_full, _ = O.train_loss_and_grads(params, src, pth, tgt, mask, target)
    for r in range(world):
        for k, got in zip(O.PARAM_NAMES, results[r]):
            np.testing.assert_allclose(got, g_full[k], atol=2e-7, err_msg="%s rank %d" % (k, r))
    # both replicas hold identical averaged gradients -> identical Adam updates
    for a, b in zip(results[0], results[1]):
        assert np.array_equal(a, b)
