"""Run under torchrun on >= 2 GPUs: data-parallel adaptation (one process per GPU, NCCL all-reduce of
the flat gradient) against the single-process CPU oracle on the full batch.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/dp_check.py"""
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from clslam_b200 import Config, DatasetConfig, DepthPosePrediction  # noqa: E402
from clslam_b200.sharding import shard_batch  # noqa: E402
from oracle import clslam_oracle as orc  # noqa: E402


def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    per, h, w = 2, 64, 96
    b = per * world
    params = orc.make_params(0)
    full = orc.make_batch(b, h, w, seed=11, coherent=True)
    noise = orc.make_noise(b, h, w, seed=11)
    pred = DepthPosePrediction(DatasetConfig(height=h, width=w), Config(batch_size=b), data_parallel=True)
    for n in orc.NET_NAMES:
        pred.models[n].load_state_dict(params[n])
    pred.noise_provider = lambda s, shape: noise[s]
    shard = shard_batch(full, rank, world)
    pred._set_adapt(True)
    _, losses = pred._process_batch(dict(shard))
    pred.backward()
    torch.cuda.synchronize()
    ok = True
    if rank == 0:
        adapter = orc.OracleAdapter(params, orc.OracleConfig(h, w, b))
        _, ref_losses = adapter.step(full, noise, update=False)
        ref = adapter.grads()
        e_loss = abs(float(losses['loss']) - float(ref_losses['loss'].detach())) / abs(float(ref_losses['loss'].detach()))
        worst = 0.0
        for key, g in pred.gradients().items():
            r = ref[key]
            worst = max(worst, float((g.cpu().double() - r.double()).norm() / (r.double().norm() + 1e-30)))
        print(f'dp_check world={world}: loss rel err {e_loss:.2e}, worst grad rel-L2 {worst:.2e}')
        ok = e_loss < 1e-4 and worst < 2e-3
    # graph-replay path with the eager all-reduce in between; weights must stay identical on all ranks
    _, l3 = pred.adapt(dict(shard), dict(shard), steps=3)
    chk = pred._flat.params.double().sum().reshape(1)
    gathered = [torch.zeros_like(chk) for _ in range(world)]
    dist.all_gather(gathered, chk)
    same = all(float(g) == float(gathered[0]) for g in gathered)
    if rank == 0:
        print(f'dp_check: 3 adapt steps ok, loss {float(l3["loss"]):.6f}, replicas identical: {same}')
        print('DP_CHECK_PASS' if (ok and same) else 'DP_CHECK_FAIL', flush=True)
    pred.close()          # the step graph holds captured NCCL kernels: release it before the communicator

    # ---- full training plan, sharded (SURVEY 8(f) row 3 x 8(e)): _set_adapt(freeze_encoder=False) -- BatchNorm in evaluation mode,
    # so the sharded step equals the full-batch step -- decoder AND encoder gradients after the all-reduce of the three flat
    # buffers against the oracle on the global batch (exact-fp32 convolution mode: tests/test_gpu_train_e2e.py explains why),
    # then two captured steps: every replica must hold bit-identical encoders
    from clslam_b200 import _lib as L
    L.load().clslam_set_conv_mode(0)
    pred = DepthPosePrediction(DatasetConfig(height=h, width=w), Config(batch_size=b), data_parallel=True)
    for n in orc.NET_NAMES:
        pred.models[n].load_state_dict(params[n])
    pred.noise_provider = lambda s, shape: noise[s]
    pred._set_adapt(freeze_encoder=False)
    assert pred._mode == 'adapt_enc'
    pred.optimizer.zero_grad()
    _, losses = pred._process_batch(dict(shard))
    pred.backward()
    torch.cuda.synchronize()
    ok2 = True
    if rank == 0:
        adapter = orc.OracleAdapter(params, orc.OracleConfig(h, w, b), mode='adapt_enc')
        _, ref_losses = adapter.step(full, noise, update=False)
        ref = adapter.grads()
        grads = dict(pred.gradients()); grads.update(pred.encoder_gradients())
        worst, wkey = 0.0, None
        for key, r in ref.items():
            e = float((grads[key].cpu().double() - r.double()).norm() / (r.double().norm() + 1e-30))
            if e > worst:
                worst, wkey = e, key
        e_loss = abs(float(losses['loss']) - float(ref_losses['loss'].detach())) / abs(float(ref_losses['loss'].detach()))
        print(f'dp_check (encoders trainable) world={world}: loss rel err {e_loss:.2e}, worst grad rel-L2 {worst:.2e} at {wkey} '
              f'over {len(ref)} tensors')
        ok2 = e_loss < 1e-4 and worst < 2e-3 and len(ref) == 88
    pred.optimizer.step()
    for _ in range(2):
        eng = pred._stage_inputs(dict(shard)); pred._set_weights(eng, None); pred._draw_noise(eng); pred._step(eng)
    torch.cuda.synchronize()
    chk = torch.stack([ef.params.double().sum() for ef in pred._main.enc_flat.values()] + [pred._flat.params.double().sum()])
    gathered = [torch.zeros_like(chk) for _ in range(world)]
    dist.all_gather(gathered, chk)
    same2 = all(torch.equal(g, gathered[0]) for g in gathered)
    moved = not torch.equal(pred.models['depth_encoder'].state_dict()['resnet.conv1.weight'].cpu(), params['depth_encoder']['resnet.conv1.weight'])
    if rank == 0:
        print(f'dp_check (encoders trainable): 3 steps ok, replicas identical: {same2}, encoder updated: {moved}')
        print('DP_TRAIN_CHECK_PASS' if (ok2 and same2 and moved) else 'DP_TRAIN_CHECK_FAIL', flush=True)
    L.load().clslam_set_conv_mode(1)
    pred.close()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
