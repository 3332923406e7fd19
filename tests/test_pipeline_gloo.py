"""CPU, world_size 2 on gloo: the layer-group pipeline's host logic (partition rule, hop order, wrap-around to
rank 0 for norm + lm_head) with a toy stage function standing in for the CUDA stage."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from kvquant_b200 import decode as kd


def test_partition_rule_matches_reference_set_devices():
    # modeling_llama.py:2442-2453: n_layers // n_gpus consecutive layers per device, remainder on the last one
    assert [kd.partition_layers(32, 4, r) for r in range(4)] == [(0, 8), (8, 16), (16, 24), (24, 32)]
    assert [kd.partition_layers(40, 8, r) for r in range(8)][-1] == (35, 40)
    assert kd.partition_layers(32, 1, 0) == (0, 32)
    cov = [kd.partition_layers(33, 4, r) for r in range(4)]
    assert cov[0][0] == 0 and cov[-1][1] == 33 and all(cov[i][1] == cov[i + 1][0] for i in range(3))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    hidden = 16
    lo, hi = kd.partition_layers(8, world, rank)
    # toy "layers": x -> x * (l + 2) + l, applied for l in [lo, hi)

    def stage_fn(x):
        for l in range(lo, hi):
            x = x * (l + 2) + l
        return x

    pipe = kd.PipelineDecoder(rank, world, hidden, torch.float64, "cpu", stage_fn,
                              head_fn=lambda y: y.sum().reshape(1), embed_fn=lambda tok: torch.full((hidden,), float(tok), dtype=torch.float64))
    outs = []
    for tok in (1, 2, 3):
        r = pipe.step(tok)
        if rank == 0:
            outs.append(float(r))
    if rank == 0:
        q.put(outs)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_pipeline_equals_single_process():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    outs = q.get(timeout=90)
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    want = []
    for tok in (1, 2, 3):
        x = torch.full((16,), float(tok), dtype=torch.float64)
        for l in range(8):
            x = x * (l + 2) + l
        want.append(float(x.sum()))
    assert outs == want


# ---------------------------------------------------------------------------------------------------------------
# sequence-sharded decode (sp): every rank owns a contiguous slice of the tokens, computes its partial (out, lse),
# one all_gather of [H*128 + H] floats per rank, exact merge -- host logic on gloo with the oracle standing in for
# the CUDA attend (the device merge kernel is covered by test_sequence_shard_merge_equals_single_attend)
# ---------------------------------------------------------------------------------------------------------------
def _sp_worker(rank, world, port, q):
    import numpy as np
    from _util import O, oracle_cache, spec
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    bits, L, H = 4, 96, 32
    c, k, v = oracle_cache(bits, L)
    qv = O.rope_rotate_q(spec().q_vec(4), L, 10000.0)
    per = L // world
    lo, hi = rank * per, (rank + 1) * per if rank < world - 1 else L
    # this rank's shard: tokens [lo, hi) at absolute positions lo.. (LayerCache.pos_base = lo on the device)
    s = c.k_scores(qv, 10000.0, 0)[:, lo:hi]

    def v_fn(p):
        full = np.zeros((H, L))
        full[:, lo:hi] = p
        return c.v_output(full)

    out, lse = O.attend_partial(s, v_fn)
    part = torch.from_numpy(np.concatenate([out.ravel(), lse]))          # the device layout: [H*128 | H]
    gath = torch.empty(world * part.numel(), dtype=part.dtype)
    dist.all_gather_into_tensor(gath, part)
    g = gath.view(world, -1).numpy()
    merged = O.merge_partials(g[:, :H * 128].reshape(world, H, 128), g[:, H * 128:])
    if rank == 0:
        _, want = O.attend_ideal(c.k_scores(qv, 10000.0, 0), lambda p: c.v_output(p))
        q.put(float(np.abs(merged - want).max() / np.abs(want).max()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_two_rank_sequence_shards_merge_to_full_attention():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_sp_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    err = q.get(timeout=150)
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    assert err < 1e-6, err   # the oracle casts through fp32 in places; the merge itself is exact
