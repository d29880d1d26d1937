#!/usr/bin/env python
"""Headline benchmark: crops/sec of ViTPose forward (+flip test) + decode on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload B-classic-17] [--crops 256]
  python bench.py --impl reference ...      # the reference algorithm (oracle port) on the host CPU cores

A step = one pass of the hot path over one batch of synthetic crops: backbone + head on the crops and their
horizontal flips, then the fused decode.  `value` times it with the crops already resident in HBM (CUDA events
on the launching stream); `e2e` times TopDown.forward_test — the call a user of the reference makes — from
pinned HOST crops to HOST keypoints (H2D + D2H inside the timed region).  One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from vitpose_b200 import configs, synthetic  # noqa: E402

METRIC = 'crops/sec ViTPose fwd(+flip)+decode 256x192'
GFLOP_PER_CROP = {  # BASELINE.md §3, one forward, no flip
    'S-classic-17': 11.19, 'B-classic-17': 37.05, 'L-simple-17': 120.85, 'H-classic-133': 251.84}
DEFAULT_CROPS = {'S-classic-17': 256, 'B-classic-17': 256, 'L-simple-17': 128, 'H-classic-133': 64}


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d['hbm_gbs'], tf_burst=d['bf16_tflops'], tf_sustained=d['bf16_tflops_sustained'],
                    source='measured')
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source='fallback')


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms. The process is started before the warm-up (it needs
    about a second to print its first line); ``mark()`` at both ends of the timed region selects the samples taken
    DURING it (plus the one straddling each end, so that a region shorter than the sampling period still has data)."""
    Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.rows, self.proc, self.marks = index, [], None, []

    def mark(self):
        self.marks.append(len(self.rows))

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), f'--query-gpu={self.Q}',
                                          '--format=csv,noheader,nounits', '-lms', '50'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['nvidia-smi unavailable'])
        time.sleep(0.12)
        self.proc.terminate()
        rows = self.rows
        if len(self.marks) >= 2:
            rows = self.rows[max(0, self.marks[0] - 1):self.marks[-1] + 1]
        sm = [float(r[0]) for r in rows if len(r) >= 6 and r[0].replace('.', '').isdigit()]
        mx = [float(r[1]) for r in rows if len(r) >= 6 and r[1].replace('.', '').isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = sorted({names[i] for r in rows if len(r) >= 6 for i in range(4) if r[2 + i] == 'Active'})
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=reasons, samples=len(sm))


def oracle_crops_per_sec(cfg, sd, n, K, threads, steps=1, warmup=1):
    """The reference algorithm (oracle/vitpose_torch.py + oracle/decode_np.py) on the host CPU."""
    from oracle import vitpose_torch as VT
    torch.set_num_threads(threads)
    img = synthetic.synthetic_crops(n, 0)
    metas = synthetic.synthetic_metas(n, K, 0)
    for _ in range(warmup):
        VT.forward_test(sd, img[:min(n, 4)], metas[:min(n, 4)], cfg)
    t0 = time.perf_counter()
    for _ in range(steps):
        VT.forward_test(sd, img, metas, cfg)
    dt = (time.perf_counter() - t0) / steps
    return n / dt, dt


def run_reference(args, cfg, K):
    rank = int(os.environ.get('RANK', 0))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    sd = synthetic.scaled_init_state_dict(cfg, 0)
    # bounded sample: calibrate on 4 crops, then size each step for ~10 s of CPU work
    rate, _ = oracle_crops_per_sec(cfg, sd, 4, K, threads, steps=1, warmup=1)
    n = int(max(4, min(args.crops, rate * 10)))
    cps, dt = oracle_crops_per_sec(cfg, sd, n, K, threads, steps=max(1, args.steps), warmup=0)
    line = dict(metric=METRIC, value=cps, unit='crops/s', impl='reference', n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=dt * 1e3, higher_is_better=True, scaling='weak', vs_baseline=None,
                dtype='f32', data='synthetic',
                config=dict(workload=args.workload, crops_per_step=n, flip_test=True,
                            decode=cfg['test_cfg'].get('use_udp') and 'udp_dark' or 'default'),
                cpu_baseline=dict(value=cps, unit='crops/s', cores=threads, kind='port',
                                  sample=f'{n} crops/step x {max(1, args.steps)} steps, torch fp32 eager CPU + numpy '
                                         f'decode (oracle port of the reference; the reference tree is not on the box)'),
                e2e=dict(value=cps, unit='crops/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line))


TRAIN_METRIC = 'crops/sec ViTPose-B training step (forward_train + backward + grad all-reduce + layer-decay AdamW) 256x192'


def run_train(args, cfg, K):
    """BASELINE.json configs[4]: one step = TopDown.forward_train -> loss.backward() -> gradient all-reduce (NCCL,
    N > 1) -> grad-norm clip + layer-decay AdamW, per-GPU batch fixed (weak scaling)."""
    import torch.distributed as dist
    import vitpose_b200 as V
    from vitpose_b200 import _lib, parallel
    from vitpose_b200.optim import LayerDecayOptimizerConstructor

    world = int(os.environ.get('WORLD_SIZE', 1))
    rank = int(os.environ.get('RANK', 0))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    n = args.crops
    drop = float(cfg['backbone'].get('drop_path_rate', 0.0))      # the config's stochastic depth (ViTPose-B: 0.3)
    sd = synthetic.scaled_init_state_dict(cfg, 0)
    model = V.build_posenet(cfg)
    model.load_state_dict(sd, strict=True)
    model = model.cuda().train()
    params = [p for p in model.parameters()]
    opt = LayerDecayOptimizerConstructor(dict(type='AdamW', lr=5e-4, betas=(0.9, 0.999), weight_decay=0.1),
                                         dict(num_layers=cfg['backbone']['depth'], layer_decay_rate=0.75))(model)
    g = torch.Generator().manual_seed(rank)

    def make_batch(seed):
        img = synthetic.synthetic_crops(n, seed=seed)
        ys, xs = torch.meshgrid(torch.arange(64.), torch.arange(48.), indexing='ij')
        cx, cy = torch.rand(n, K, generator=g) * 47, torch.rand(n, K, generator=g) * 63
        tgt = torch.exp(-((xs - cx[..., None, None]) ** 2 + (ys - cy[..., None, None]) ** 2) / 8.0)
        return img.pin_memory(), tgt.contiguous().pin_memory(), torch.ones(n, K, 1).pin_memory()

    host = [make_batch(rank * 7 + i) for i in range(2)]
    devb = [tuple(t.to(dev) for t in h) for h in host]

    def step(batch):
        out = model.train_step(dict(img=batch[0], target=batch[1], target_weight=batch[2], img_metas=None), opt)
        opt.zero_grad(set_to_none=True)
        out['loss'].backward()          # N > 1: the gradient all-reduce (NCCL) is issued inside, overlapped
        opt.step(max_norm=1.0)
        return out['loss']

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step(devb[i & 1])
    barrier()
    sys.stdout.flush()
    os.dup2(saved_stdout, 1)
    os.close(saved_stdout)
    calls0 = _lib.ABI_CALLS[0]
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.mark()
    ev0.record()
    for i in range(args.steps):
        loss = step(devb[i & 1])
    ev1.record()
    barrier()
    sampler.mark()
    ms = ev0.elapsed_time(ev1) / args.steps
    calls = _lib.ABI_CALLS[0] - calls0
    clocks = sampler.stop() if rank == 0 else None
    # end to end: pinned host batch -> device every step, loss value back on the host
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        hb = host[i & 1]
        loss = step(tuple(t.to(dev, non_blocking=True) for t in hb))
        loss_host = loss.item()
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) / args.steps * 1e3
    t = torch.tensor([ms, e2e_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_ms = t.tolist()
    if rank == 0:
        pk = peaks()
        total = n * world
        value = total / (ms / 1e3)
        gf = GFLOP_PER_CROP[args.workload] * 3          # forward + dgrad + wgrad
        tf = value * gf / 1e3 / world
        line = dict(metric=TRAIN_METRIC, value=value, unit='crops/s', n_gpus=world, steps=args.steps,
                    warmup=args.warmup, ms_per_step=ms, higher_is_better=True, scaling='weak', vs_baseline=None,
                    dtype='bf16', data='synthetic',
                    config=dict(workload=args.workload + '-train', crops_per_gpu=n, global_crops=total,
                                optimizer='AdamW lr 5e-4 wd 0.1, layer decay 0.75, grad clip 1.0', drop_path=drop,
                                parallelism=f'dp{world}', collective='gradient all-reduce (NCCL)' if world > 1 else None,
                                l2='activations per step >> 126 MB L2; inputs ping-pong between two buffers'),
                    roofline=dict(bound='tensor', kernel='whole training step (3 x forward GEMM FLOPs)', achieved=tf,
                                  peak=pk['tf_sustained'], unit='TFLOP/s', frac=tf / pk['tf_sustained'], traffic=None,
                                  peak_source=f"{pk['source']} sustained bf16"),
                    clocks=clocks, gpu_launches=int(calls), final_loss=loss_host,
                    e2e=dict(value=total / (e2e_ms / 1e3), unit='crops/s',
                             h2d_bytes_per_step=int(sum(t.numel() * 4 for t in host[0])), d2h_bytes_per_step=4,
                             ms_per_step=e2e_ms, api='TopDown.train_step + loss.backward() + LayerDecayAdamW.step'))
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--workload', default='B-classic-17', choices=sorted(configs.BASELINE_CONFIGS))
    ap.add_argument('--crops', type=int, default=0, help='crops per GPU per step (default: BASELINE batch)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--train', action='store_true', help='measure the training step (BASELINE configs[4]) instead')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'b200' else args.warmup
    if args.crops <= 0:
        args.crops = DEFAULT_CROPS[args.workload]
    cfg = configs.baseline_model_cfg(args.workload)
    K = cfg['keypoint_head']['out_channels']
    if args.impl == 'reference':
        run_reference(args, cfg, K)
        return
    if args.train:
        if args.crops == DEFAULT_CROPS[args.workload]:
            args.crops = 64                      # per-GPU batch of the upstream training logs
        run_train(args, cfg, K)
        return

    import torch.distributed as dist
    import vitpose_b200 as V
    from vitpose_b200 import _lib

    world = int(os.environ.get('WORLD_SIZE', 1))
    rank = int(os.environ.get('RANK', 0))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    sampler = ClockSampler(local_rank)      # started now: nvidia-smi needs about a second before its first sample
    if rank == 0:
        sampler.start()
    # NCCL prints its version banner on stdout at the first collective; keep stdout for the one JSON line
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)

    n = args.crops
    sd = synthetic.scaled_init_state_dict(cfg, 0)
    model = V.build_posenet(cfg)
    model.load_state_dict(sd, strict=True)
    model = model.cuda().eval()
    test_cfg = cfg['test_cfg']
    # two distinct pinned host batches (ping-pong) so no step re-reads the previous step's input
    host = [synthetic.synthetic_crops(n, seed=rank * 7 + i).pin_memory() for i in range(2)]
    metas = synthetic.synthetic_metas(n, K, seed=rank)
    dev_img = [h.to(dev) for h in host]
    eng = model._engine()
    from vitpose_b200.core.post_processing import flip_index_from_pairs
    from vitpose_b200.engine import decode_mode_from_cfg
    flip_index = torch.from_numpy(flip_index_from_pairs(K, metas[0]['flip_pairs'])).to(dev)
    center = torch.from_numpy(np.stack([m['center'] for m in metas])).to(dev)
    scale = torch.from_numpy(np.stack([m['scale'] for m in metas])).to(dev)
    mode = decode_mode_from_cfg(test_cfg)
    from vitpose_b200 import parallel

    def device_step(i):
        hm, _ = eng.forward_heatmaps(dev_img[i & 1], flip=True)
        r = eng.decode(hm, n, True, flip_index, bool(test_cfg.get('shift_heatmap', False)), mode,
                       test_cfg.get('modulate_kernel', 11), bool(test_cfg.get('use_udp', False)), center, scale)
        out = torch.cat([r['preds'], r['maxvals']], dim=2)
        return parallel.gather_contiguous(out)        # the path's only collective: final result gather

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        device_step(i)
    barrier()
    sys.stdout.flush()
    os.dup2(saved_stdout, 1)
    os.close(saved_stdout)
    # ---- timed region (device-resident inputs) -------------------------------------------------------
    L = _lib.lib()
    launches0 = L.vpb_launch_count()
    L.vpb_profile_enable(1)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.mark()
    ev0.record()
    for i in range(args.steps):
        device_step(i)
    ev1.record()
    barrier()
    sampler.mark()
    ms = ev0.elapsed_time(ev1) / args.steps
    records = _lib.profile_records()
    L.vpb_profile_enable(0)
    launches = L.vpb_launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None

    # ---- end to end through the reference-facing API: pinned host crops -> host keypoints ----------------
    for i in range(2):
        model(img=host[i & 1], img_metas=metas, return_loss=False)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        res = model(img=host[i & 1], img_metas=metas, return_loss=False)
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / args.steps
    t = torch.tensor([ms, e2e_s * 1e3], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_ms = t.tolist()

    if rank == 0:
        pk = peaks()
        # dominant kernel: the transformer GEMMs (fc1 carries the largest share); per-launch average inside the step
        by_tag = {}
        for tag, v in records:
            by_tag.setdefault(tag, []).append(v)
        bb = cfg['backbone']
        D, hidden, rows = bb['embed_dim'], int(bb['embed_dim'] * bb['mlp_ratio']), 2 * n * 192
        flops = {'gemm_qkv': 2.0 * rows * 3 * D * D, 'gemm_proj_ln': 2.0 * rows * D * D,
                 'gemm_fc1': 2.0 * rows * hidden * D, 'gemm_fc2_ln': 2.0 * rows * hidden * D}
        dom = 'gemm_fc1'
        dom_ms = float(np.mean(by_tag[dom]))
        achieved = flops[dom] / dom_ms / 1e9
        tgemm_ms = sum(float(np.sum(by_tag[k])) for k in flops) / args.steps
        tgemm_tf = sum(flops[k] * len(by_tag[k]) for k in flops) / args.steps / tgemm_ms / 1e9
        shares = {k: round(float(np.sum(v)) / args.steps / ms, 4) for k, v in by_tag.items()}
        # DRAM bytes of one fc1 launch from the committed ncu --set full capture of this command (B workload, 256 crops)
        traffic = None
        tpath = os.path.join(ROOT, 'profiles', 'r01_fc1_traffic.json')
        if args.workload == 'B-classic-17' and n == 256 and os.path.exists(tpath):
            traffic = json.load(open(tpath))['traffic_bytes']
        roofline = dict(bound='tensor', kernel='gemm_bf16_tn_kernel<256,GELU> (mlp.fc1)', achieved=achieved,
                        peak=pk['tf_sustained'], unit='TFLOP/s', frac=achieved / pk['tf_sustained'], traffic=traffic,
                        peak_source=f"{pk['source']} sustained bf16 (kernel timed inside a long step)",
                        transformer_gemms_tflops=tgemm_tf, transformer_gemms_frac=tgemm_tf / pk['tf_sustained'],
                        step_share_by_kernel=shares)
        total_crops = n * world
        value = total_crops / (ms / 1e3)
        gf = GFLOP_PER_CROP[args.workload] * 2
        line = dict(metric=METRIC, value=value, unit='crops/s', n_gpus=world, steps=args.steps, warmup=args.warmup,
                    ms_per_step=ms, higher_is_better=True, scaling='weak', vs_baseline=None, dtype='bf16',
                    data='synthetic',
                    config=dict(workload=args.workload, crops_per_gpu=n, global_crops=total_crops, flip_test=True,
                                decode={3: 'udp_dark', 2: 'unbiased', 1: 'default', 0: 'none'}[mode],
                                parallelism=f'dp{world}', l2='activations per step >> 126 MB L2; inputs ping-pong '
                                                              'between two buffers',
                                weights='random scaled-init (no checkpoints offline)'),
                    model_tflops=value * gf / 1e3, model_frac_of_peak=value * gf / 1e3 / world / pk['tf_sustained'],
                    roofline=roofline, clocks=clocks, gpu_launches=int(launches),
                    e2e=dict(value=total_crops / (e2e_ms / 1e3), unit='crops/s',
                             h2d_bytes_per_step=int(host[0].numel() * 4 + n * 16),
                             d2h_bytes_per_step=int(n * K * 3 * 4), ms_per_step=e2e_ms,
                             api='TopDown.forward_test(img=<pinned host fp32>, img_metas=...)'))
        if not args.no_cpu_baseline and world == 1:     # rank 0 at N = 1 only (other ranks would compete for the cores)
            threads = os.cpu_count() or 1
            rate, _ = oracle_crops_per_sec(cfg, sd, 4, K, threads, steps=1, warmup=1)
            ns = int(max(4, min(n, rate * 12)))
            cps, dt = oracle_crops_per_sec(cfg, sd, ns, K, threads, steps=1, warmup=0)
            line['cpu_baseline'] = dict(value=cps, unit='crops/s', cores=threads, kind='port',
                                        sample=f'{ns} crops, 1 pass ({dt:.1f} s), torch fp32 eager + numpy decode '
                                               f'(oracle port of the reference forward_test)')
        assert res['preds'].shape == (n, K, 3)
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
