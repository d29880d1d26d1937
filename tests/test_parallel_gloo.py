"""N>1 host logic on CPU: world_size-2 gloo processes shard a crop list and gather results in dataset order,
reproducing DistributedSampler + collect_results_gpu semantics (mmpose/apis/test.py:168-173,195-223)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vitpose_b200 import parallel


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_total, K, out_dir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        # "model": result for crop i is a deterministic function of i
        def fake_result(i):
            return torch.full((K, 3), float(i)) + torch.arange(K * 3).reshape(K, 3) * 1e-3

        idx = parallel.strided_shard(n_total, rank, world)
        local = torch.stack([fake_result(i) for i in idx])
        full = parallel.gather_strided(local, n_total)
        lo, hi = parallel.contiguous_shard(n_total, rank, world)
        local_c = torch.stack([fake_result(i) for i in range(lo, hi)]) if hi > lo else torch.zeros(0, K, 3)
        counts = [parallel.contiguous_shard(n_total, r, world)[1] - parallel.contiguous_shard(n_total, r, world)[0]
                  for r in range(world)]
        full_c = parallel.gather_contiguous(local_c, counts)
        torch.save(dict(full=full, full_c=full_c, idx=idx), os.path.join(out_dir, f'r{rank}.pt'))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('n_total', [7, 8, 1])
def test_shard_and_gather_world2(tmp_path, n_total):
    world, K = 2, 5
    mp.spawn(_worker, args=(world, _free_port(), n_total, K, str(tmp_path)), nprocs=world, join=True)
    expect = torch.stack([torch.full((K, 3), float(i)) + torch.arange(K * 3).reshape(K, 3) * 1e-3
                          for i in range(n_total)])
    for r in range(world):
        d = torch.load(os.path.join(str(tmp_path), f'r{r}.pt'))
        assert torch.equal(d['full'], expect)          # every rank holds the ordered, truncated result
        assert torch.equal(d['full_c'], expect)
    # the reference's split: pad by wrapping, rank::world
    assert parallel.strided_shard(7, 0, 2) == [0, 2, 4, 6] and parallel.strided_shard(7, 1, 2) == [1, 3, 5, 0]


def test_single_process_passthrough():
    x = torch.arange(12.).reshape(4, 3)
    assert torch.equal(parallel.gather_strided(x, 3), x[:3])
    assert torch.equal(parallel.gather_contiguous(x), x)
    assert parallel.contiguous_shard(10, 1, 4) == (2, 5)


def _grad_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        torch.manual_seed(0)                               # identical replicas
        params = [torch.nn.Parameter(torch.randn(s)) for s in [(5, 7), (3,), (2, 2, 2), (11,)]]
        g = torch.Generator().manual_seed(100 + rank)      # rank-specific gradients
        for p in params:
            p.grad = torch.randn(p.shape, generator=g)
        params[1].grad = None if False else params[1].grad
        ncoll = parallel.allreduce_gradients(params, bucket_bytes=100)   # tiny buckets: several collectives
        torch.save(dict(grads=[p.grad.clone() for p in params], ncoll=ncoll), os.path.join(out_dir, f'g{rank}.pt'))
    finally:
        dist.destroy_process_group()


def test_allreduce_gradients_world2(tmp_path):
    world = 2
    mp.spawn(_grad_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    shapes = [(5, 7), (3,), (2, 2, 2), (11,)]
    expect = []
    for shp in shapes:
        expect.append(torch.zeros(shp))
    for r in range(world):
        g = torch.Generator().manual_seed(100 + r)
        for i, shp in enumerate(shapes):
            expect[i] += torch.randn(shp, generator=g) / world
    for r in range(world):
        d = torch.load(os.path.join(str(tmp_path), f'g{r}.pt'))
        assert d['ncoll'] >= 2
        for got, exp in zip(d['grads'], expect):
            assert torch.allclose(got, exp, atol=1e-6)


def _log_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from vitpose_b200.detectors.top_down import TopDown
        losses = dict(heatmap_loss=torch.tensor(1.0 + rank), acc_pose=torch.tensor(0.25 * (rank + 1)), extra=0.5,
                      aux_loss=[torch.tensor([2.0 * rank, 2.0 * rank + 2.0])])
        loss, log_vars = TopDown._parse_losses(TopDown, losses)
        torch.save(dict(loss=loss, log_vars=log_vars), os.path.join(out_dir, f'l{rank}.pt'))
    finally:
        dist.destroy_process_group()


def test_parse_losses_averages_logged_values_over_ranks(tmp_path):
    """mmpose/models/detectors/base.py:37-76: the returned loss is the LOCAL sum of the '*loss*' entries (backward runs
    on it); the logged values are averaged over the ranks and handed out as Python floats."""
    world = 2
    mp.spawn(_log_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        d = torch.load(os.path.join(str(tmp_path), f'l{r}.pt'))
        local_aux = 2.0 * r + 1.0
        assert abs(float(d['loss']) - (1.0 + r + local_aux)) < 1e-6
        lv = d['log_vars']
        assert all(isinstance(v, float) for v in lv.values())
        assert abs(lv['heatmap_loss'] - 1.5) < 1e-6 and abs(lv['acc_pose'] - 0.375) < 1e-6 and lv['extra'] == 0.5
        assert abs(lv['aux_loss'] - 2.0) < 1e-6 and abs(lv['loss'] - 3.5) < 1e-6
    # single process: same values, no collective
    from vitpose_b200.detectors.top_down import TopDown
    loss, lv = TopDown._parse_losses(TopDown, dict(heatmap_loss=torch.tensor(2.0), acc_pose=torch.tensor(0.5)))
    assert float(loss) == 2.0 and lv == dict(heatmap_loss=2.0, acc_pose=0.5, loss=2.0)
    with pytest.raises(TypeError):
        TopDown._parse_losses(TopDown, dict(bad_loss='x'))
