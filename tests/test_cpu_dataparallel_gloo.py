"""world_size-2 gloo test of the data-parallel host logic: sharding by sample with
GLOBAL loss weights, the reference's sample-0 smoothness evaluated on rank 0 only, per-rank noise
slices and ONE sum all-reduce of the flat gradient reproduce the single-process gradients.
The per-rank compute here is the CPU oracle (the CUDA path needs a GPU); what is under test is
clslam_b200.sharding."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from clslam_b200.sharding import allreduce_sum_, global_sample_weights, shard_batch, shard_noise, shard_range
from oracle import clslam_oracle as orc

B, H, W = 4, 64, 96


def _free_port():
    s = socket.socket(); s.bind(('127.0.0.1', 0)); p = s.getsockname()[1]; s.close()
    return p


def _flat_grads(adapter):
    return torch.cat([g.reshape(-1) for g in adapter.grads().values()])


def _make_full(replay):
    """plain coherent batch, or a replay-like one (per-sample intrinsics, colour-augmented network inputs)"""
    return orc.make_replay_batch(B, H, W, seed=7) if replay else orc.make_batch(B, H, W, seed=7, coherent=True)


def _worker(rank, world, port, out, replay):
    torch.set_num_threads(2)
    dist.init_process_group('gloo', init_method=f'tcp://127.0.0.1:{port}', rank=rank, world_size=world)
    params = orc.make_params(0)
    full = _make_full(replay)
    noise = orc.make_noise(B, H, W, seed=7)
    w_global = global_sample_weights(B, online_loss_weight=0.4, online_index=0)
    lo, hi = shard_range(B, rank, world)
    shard = shard_batch(full, rank, world)
    adapter = orc.OracleAdapter(params, orc.OracleConfig(H, W, hi - lo))
    _, losses = adapter.step(shard, shard_noise(noise, rank, world), sample_weights=w_global[lo:hi], update=False,
                             smooth_weights=w_global if rank == 0 else None, with_smooth=(rank == 0))
    flat = _flat_grads(adapter)
    loss = losses['loss'].detach().clone()
    allreduce_sum_([flat, loss])
    if rank == 0:
        torch.save({'flat': flat, 'loss': loss}, out)
    dist.destroy_process_group()


@pytest.mark.parametrize('replay', [False, True])
def test_two_rank_sharding_matches_single_process(tmp_path, replay):
    out = tmp_path / 'dp.pt'
    mp.spawn(_worker, args=(2, _free_port(), str(out), replay), nprocs=2, join=True)
    got = torch.load(out)
    params = orc.make_params(0)
    full = _make_full(replay)
    noise = orc.make_noise(B, H, W, seed=7)
    w_global = global_sample_weights(B, online_loss_weight=0.4, online_index=0)
    adapter = orc.OracleAdapter(params, orc.OracleConfig(H, W, B))
    _, losses = adapter.step(full, noise, sample_weights=w_global, update=False)
    ref = _flat_grads(adapter)
    assert abs(float(got['loss']) - float(losses['loss'])) < 1e-6 * abs(float(losses['loss']))
    assert float((got['flat'] - ref).norm() / ref.norm()) < 1e-5
