"""Data-parallel schedules on real GPUs (needs >= 2; skipped otherwise): two ranks, each with half of
a global batch, must end every step with bit-identical parameters on both ranks, equal (to fp32
rounding) to one engine training on the whole batch -- for both the all-reduce and the sharded
(reduce-scatter / sliced Adam / all-gather, overlapped) schedules."""
import os

import numpy as np
import pytest

from oracle import path_attention_oracle as O

pytestmark = pytest.mark.gpu

DIMS = O.Dims(token_vocab=2003, path_vocab=1009, target_vocab=3001, embed_dim=32, code_dim=96, max_contexts=20)
B_LOCAL = 32


def _worker(rank, world, port, schedule, math_mode, out_dir):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from code2vec_b200.engine import EngineDims, PathAttentionEngine
    from code2vec_b200.trainer import Trainer, make_fully_sharded_engine, target_row_block
    gdims = EngineDims(DIMS.token_vocab, DIMS.path_vocab, DIMS.target_vocab, DIMS.embed_dim, DIMS.code_dim,
                       DIMS.max_contexts, B_LOCAL, 10)
    full = O.init_params(DIMS, seed=4321)
    if schedule == "fully_sharded":
        eng = make_fully_sharded_engine(gdims, B_LOCAL, device=rank)
        r0, r1 = target_row_block(DIMS.target_vocab, rank, world)
        eng.load_params(dict(full, tgt=full["tgt"][r0:r1]))
    else:
        eng = PathAttentionEngine(gdims, device=rank, training=True)
        eng.load_params(full)
    eng.set_option("math_mode", math_mode)
    tr = Trainer(eng, keep_prob=1.0, seed=0, schedule=schedule, allow_single_rank=True)
    assert tr.schedule == schedule
    src, pth, tgt, mask, target = O.synthetic_batch(DIMS, B_LOCAL * world, seed=77)
    lo, hi = rank * B_LOCAL, (rank + 1) * B_LOCAL
    losses = []
    for _ in range(3):
        losses.append(tr.step_host(src[lo:hi], pth[lo:hi], tgt[lo:hi], mask[lo:hi], target[lo:hi]))
    out = eng.export_params()
    if schedule in ("table_sharded", "fully_sharded"):
        sh = eng.export_table_shards()
        out["tok_shard"], out["path_shard"] = sh["tok"], sh["path"]
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), losses=np.array(losses), **out)
    dist.destroy_process_group()


@pytest.mark.parametrize("schedule,math_mode", [("allreduce", 0), ("sharded", 0), ("table_sharded", 0), ("fully_sharded", 0),
                                                 ("fully_sharded", 1), ("fully_sharded", 2)])
def test_two_rank_data_parallel_matches_single_engine(tmp_path, schedule, math_mode):
    """math_mode 1 / 2: the tensor-core paths of the fully sharded schedule (row-sharded target table: the gradient GEMMs
    build dL/dlogits from the local logits slab with the target's row offset; 3xTF32 splits), at their tolerances."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    _run_and_compare(tmp_path, schedule, math_mode, world=2)


@pytest.mark.parametrize("schedule,math_mode", [("fully_sharded", 0), ("fully_sharded", 1), ("fully_sharded", 2)])
def test_sharded_schedules_on_a_single_rank_match_the_fused_step(tmp_path, schedule, math_mode):
    """The same schedules with a process group of ONE rank, so a one-GPU box exercises the multi-GPU code path end to end:
    tables re-homed into CUDA-IPC memory, the phase-split entry points (c2v_context_forward / c2v_target_forward /
    c2v_lse_combine / c2v_target_backward / c2v_context_backward), the collectives (trivial here) and the per-shard
    Adam -- against c2v_train_batch_host's fused step on the same batch."""
    _run_and_compare(tmp_path, schedule, math_mode, world=1)


def _run_and_compare(tmp_path, schedule, math_mode, world):
    import torch.multiprocessing as mp
    port = 29500 + (os.getpid() % 1000)
    # tf32: an element whose tiny gradient changes sign moves the other way by a full Adam step (1e-3) in each of the 3 steps
    tol = {0: 5e-5, 1: 8e-3, 2: 2e-4}[math_mode]
    mp.spawn(_worker, args=(world, port, schedule, math_mode, str(tmp_path)), nprocs=world, join=True)
    r0 = np.load(str(tmp_path / "rank0.npz"))
    r1 = np.load(str(tmp_path / "rank1.npz")) if world > 1 else r0
    replicated = {"table_sharded": ("tgt", "W", "a"), "fully_sharded": ("W", "a")}.get(schedule, O.PARAM_NAMES)
    for k in replicated:
        assert np.array_equal(r0[k], r1[k]), "replicas diverged on %s" % k
    # single engine on the global batch (mean loss over 2*B_LOCAL == average of the two local means)
    from tests.util import make_engine
    eng, _ = make_engine(DIMS, max_batch=B_LOCAL * world)
    src, pth, tgt, mask, target = O.synthetic_batch(DIMS, B_LOCAL * world, seed=77)
    for _ in range(3):
        eng.train_batch_host(src, pth, tgt, mask, target, keep=1.0)
    ref = eng.export_params()
    for k in replicated:
        assert np.abs(r0[k] - ref[k]).max() < tol, k
    if schedule == "fully_sharded":
        from code2vec_b200.trainer import target_row_block
        for r, res in ((0, r0), (1, r1)):
            if r >= world:
                continue
            lo, hi = target_row_block(DIMS.target_vocab, r, world)
            assert np.abs(res["tgt"][:hi - lo] - ref["tgt"][lo:hi]).max() < tol, r
        assert abs(float(r0["losses"][0]) - float(r1["losses"][0])) < 1e-6      # the loss is global here
    if schedule in ("table_sharded", "fully_sharded"):
        # row r of the global table lives on rank r % world at local row r // world
        for r, res in ((0, r0), (1, r1)):
            if r >= world:
                continue
            for name, shard in (("tok", res["tok_shard"]), ("path", res["path_shard"])):
                want = ref[name][r::world]
                assert np.abs(shard[:want.shape[0]] - want).max() < tol, (name, r)
