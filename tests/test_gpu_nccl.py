"""Multi-process NCCL tests (-m gpu, skipped with fewer than two devices): the training step's overlapped gradient
all-reduce (mmpose/apis/train.py:129-133: what MMDistributedDataParallel does for the reference) on real NVLink, and a
stress run of the fused GEMM + LayerNorm kernels — whose column-tile CTAs wait for each other — while NCCL's own CTAs
are on the device ('needs a stress test ... >= 200 steps')."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, steps, out_dir):
    import torch.distributed as dist
    import vitpose_b200 as V
    from vitpose_b200 import configs, parallel, synthetic
    from vitpose_b200.optim import LayerDecayOptimizerConstructor
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    try:
        cfg = configs.baseline_model_cfg('B-classic-17')
        cfg['backbone'].update(depth=2, drop_path_rate=0.0)
        sd = synthetic.scaled_init_state_dict(cfg, 0)
        n, K = 8, 17
        img = synthetic.synthetic_crops(n, 100 + rank).cuda()
        tgt = torch.rand(n, K, 64, 48, generator=torch.Generator().manual_seed(rank)).cuda()
        tw = torch.ones(n, K, 1).cuda()
        batch = dict(img=img, target=tgt, target_weight=tw, img_metas=None)
        # (1) overlapped exchange inside backward() == local backward + explicit bucketed all-reduce
        grads = {}
        for mode in ('in_backward', 'explicit'):
            model = V.build_posenet(cfg)
            model.load_state_dict(sd)
            model = model.cuda().train()
            model.allreduce_in_backward = mode == 'in_backward'
            out = model.train_step(batch, None)
            assert isinstance(out['log_vars']['loss'], float)          # rank-averaged, as base.py:66-74
            out['loss'].backward()
            if mode == 'explicit':
                parallel.allreduce_gradients(list(model.parameters()))
            grads[mode] = {k: p.grad.clone() for k, p in model.named_parameters()}
        worst = max(float((grads['in_backward'][k] - grads['explicit'][k]).norm() / (grads['explicit'][k].norm() + 1e-30))
                    for k in grads['explicit'])
        # (2) stress: `steps` training steps; a side stream keeps an NCCL all-reduce in flight all the time, so NCCL's
        # CTAs share the device with the fused GEMM + LayerNorm kernels of every forward pass
        model = V.build_posenet(cfg)
        model.load_state_dict(sd)
        model = model.cuda().train()
        opt = LayerDecayOptimizerConstructor(dict(type='AdamW', lr=1e-4, betas=(0.9, 0.999), weight_decay=0.1),
                                             dict(num_layers=2, layer_decay_rate=0.75))(model)
        noise = torch.zeros(32 << 20, device=dev)
        side = torch.cuda.Stream(dev)
        losses = []
        for it in range(steps):
            with torch.cuda.stream(side):
                h = dist.all_reduce(noise, async_op=True)
            out = model.train_step(batch, opt)
            opt.zero_grad(set_to_none=True)
            out['loss'].backward()
            opt.step(max_norm=1.0)
            h.wait()
            losses.append(out['log_vars']['loss'])
        torch.cuda.synchronize()
        flat = torch.cat([p.detach().flatten() for p in model.parameters()])
        other = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(other, flat)
        same = all(torch.equal(o, flat) for o in other)
        torch.save(dict(worst=worst, same=same, losses=losses, finite=bool(torch.isfinite(flat).all())),
                   os.path.join(out_dir, f'n{rank}.pt'))
    finally:
        dist.destroy_process_group()


def test_overlapped_allreduce_and_ln_gemm_under_nccl(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip('needs two GPUs (run with gpurun --gpus 2)')
    import torch.multiprocessing as mp
    world, steps = 2, 200
    mp.spawn(_worker, args=(world, _free_port(), steps, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        d = torch.load(os.path.join(str(tmp_path), f'n{r}.pt'))
        assert d['worst'] < 1e-5, f"overlapped vs explicit all-reduce differ by {d['worst']:.2e}"
        assert d['same'] and d['finite'], 'replicas diverged'
        assert len(d['losses']) == steps and d['losses'][-1] < d['losses'][0], 'loss did not decrease'


def _moe_worker(rank, world, port, out_dir):
    """ViTPose+ on two ranks whose batches hold DIFFERENT datasets (rank 0: datasets 0 and 2, rank 1: dataset 1 only):
    the gradient arena is exchanged piecewise inside backward(), so its layout must not depend on which experts /
    heads a rank's batch touches — every expert gets a (possibly zero) gradient on every rank, as the reference's
    dense masked experts do "to support ddp training" (vit_moe.py:107-111)."""
    import torch.distributed as dist
    import vitpose_b200 as V
    from vitpose_b200 import configs, parallel, synthetic
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    try:
        cfg = configs.tiny_model_cfg(5, depth=2)
        cfg['type'] = 'TopDownMoE'
        cfg['backbone'].update(type='ViTMoE', num_expert=3, part_features=32, drop_path_rate=0.0)
        cfg['associate_keypoint_head'] = [dict(cfg['keypoint_head']), dict(cfg['keypoint_head'])]
        torch.manual_seed(0)
        ref = V.build_posenet(cfg)
        sd = {k: v.clone() for k, v in ref.state_dict().items()}
        src = [2, 0, 0, 2] if rank == 0 else [1, 1, 1, 1]
        n, K = len(src), 5
        img = synthetic.synthetic_crops(n, 200 + rank).cuda()
        tgt = torch.rand(n, K, 64, 48, generator=torch.Generator().manual_seed(10 + rank)).cuda()
        tw = torch.ones(n, K, 1).cuda()
        batch = dict(img=img, target=tgt, target_weight=tw, img_metas=[dict(dataset_idx=d) for d in src])
        grads = {}
        for mode in ('in_backward', 'explicit'):
            model = V.build_posenet(cfg)
            model.load_state_dict(sd)
            model = model.cuda().train()
            model.allreduce_in_backward = mode == 'in_backward'
            out = model.train_step(batch, None)
            out['loss'].backward()
            if mode == 'explicit':
                parallel.allreduce_gradients([p for p in model.parameters() if p.grad is not None])
            grads[mode] = {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None}
        assert set(grads['in_backward']) == set(grads['explicit'])
        worst = max(float((grads['in_backward'][k] - grads['explicit'][k]).norm() / (grads['explicit'][k].norm() + 1e-30))
                    for k in grads['explicit'])
        flat = torch.cat([grads['in_backward'][k].flatten() for k in sorted(grads['in_backward'])])
        other = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(other, flat)
        # every expert of block 0 has a non-zero averaged gradient on BOTH ranks (each dataset is on one of them)
        nz = [float(grads['in_backward'][f'backbone.blocks.0.mlp.experts.{e}.weight'].abs().max()) > 0 for e in range(3)]
        torch.save(dict(worst=worst, same=all(torch.equal(o, flat) for o in other), nz=nz,
                        finite=bool(torch.isfinite(flat).all())), os.path.join(out_dir, f'm{rank}.pt'))
    finally:
        dist.destroy_process_group()


def test_vitpose_plus_training_ranks_with_different_datasets(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip('needs two GPUs (run with gpurun --gpus 2)')
    import torch.multiprocessing as mp
    world = 2
    mp.spawn(_moe_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        d = torch.load(os.path.join(str(tmp_path), f'm{r}.pt'))
        assert d['worst'] < 1e-5, f"overlapped vs explicit all-reduce differ by {d['worst']:.2e}"
        assert d['same'] and d['finite'], 'ranks hold different averaged gradients'
        assert d['nz'] == [True, True, True], d['nz']
