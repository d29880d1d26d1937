#!/usr/bin/env python
"""Headline benchmark: crops/sec of ViTPose forward (+flip test) + decode on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload B-classic-17] [--crops 256]
  python bench.py --impl reference ...      # the reference's own CPU path on the host cores
  python bench.py --train ...               # BASELINE configs[4] alone (training step)

A step = one pass of the hot path over one batch of synthetic crops: backbone + head on the crops and their
horizontal flips, then the fused decode.  `value` times it with the crops already resident in HBM (CUDA events
on the launching stream); `e2e` times TopDown.forward_test — the call a user of the reference makes — from
pinned HOST crops to HOST keypoints (H2D + D2H inside the timed region).  One JSON line on stdout (rank 0).

The top-level record is BASELINE configs[1] (ViTPose-B classic, 256 crops per GPU, weak scaling).  Unless
`--no-extra` is given the same line carries `configs`: sub-records for BASELINE configs[2..4] measured in the same
process right after the headline —
  L-simple-17   1024 crops in total, split over the N GPUs (strong scaling), UDP-DARK decode;
  H-classic-133 256 crops per GPU (weak scaling; 2048 over 8 GPUs), quarter-offset decode + shifted flip;
  B-train       ViTPose-B training step, 64 crops per GPU, gradient all-reduce over NCCL when N > 1.
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
TRAIN_METRIC = 'crops/sec ViTPose-B training step (forward_train + backward + grad all-reduce + layer-decay AdamW) 256x192'
GFLOP_PER_CROP = {  # BASELINE.md §3, one forward, no flip
    'S-classic-17': 11.19, 'B-classic-17': 37.05, 'L-simple-17': 120.85, 'H-classic-133': 251.84}
DEFAULT_CROPS = {'S-classic-17': 256, 'B-classic-17': 256, 'L-simple-17': 128, 'H-classic-133': 64}
REF_DIR = os.path.join(ROOT, 'baseline', '_ref')


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d['hbm_gbs'], tf_burst=d['bf16_tflops'], tf_sustained=d['bf16_tflops_sustained'],
                    source='measured')
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source='fallback')


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms by one background process (it needs about a second
    to print its first line, so it is started before the first warm-up). ``mark()`` returns the current sample
    index; ``region(a, b)`` summarises the samples taken between two marks (plus the one straddling each end, so
    that a region shorter than the sampling period still has data)."""
    Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

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

    def mark(self):
        return len(self.rows)

    def region(self, a, b):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['nvidia-smi unavailable'])
        time.sleep(0.12)                      # let the sample that straddles the end arrive
        rows = self.rows[max(0, a - 1):b + 1]
        sm = [float(r[0]) for r in rows if len(r) >= 6 and r[0].replace('.', '').isdigit()]
        mx = [float(r[1]) for r in rows if len(r) >= 6 and r[1].replace('.', '').isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = sorted({names[i] for r in rows if len(r) >= 6 for i in range(4) if r[2 + i] == 'Active'})
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=reasons, samples=len(sm))

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()


# ---- reference arm ---------------------------------------------------------------------------------------------
def reference_available():
    return os.path.isfile(os.path.join(REF_DIR, 'mmpose', 'models', 'backbones', 'vit.py'))


def cpu_crops_per_sec(cfg, sd, n, K, threads, steps=1, warmup=1, use_reference=None):
    """The reference algorithm on the host CPU: the UNMODIFIED reference modules from baseline/_ref (copied there by
    __graft_entry__.build(), loaded by oracle/ref_loader.py; kind 'reference') when present, else the oracle port
    (oracle/vitpose_torch.py + oracle/decode_np.py; kind 'port'). Returns (crops/s, seconds per step, kind)."""
    torch.set_num_threads(threads)
    img = synthetic.synthetic_crops(n, 0)
    metas = synthetic.synthetic_metas(n, K, 0)
    if use_reference is None:
        use_reference = reference_available()
    if use_reference:
        os.environ['VITPOSE_REFERENCE_ROOT'] = REF_DIR
        from oracle import ref_loader
        ref_loader.REF_ROOT = REF_DIR
        model = ref_loader.build_reference_topdown(cfg)
        model.load_state_dict(sd, strict=True)
        model.eval()

        def run(i, m):
            with torch.no_grad():
                return model(img=i, img_metas=m, return_loss=False)
        kind = 'reference'
    else:
        from oracle import vitpose_torch as VT

        def run(i, m):
            return VT.forward_test(sd, i, m, cfg)
        kind = 'port'
    for _ in range(warmup):
        run(img[:min(n, 4)], metas[:min(n, 4)])
    t0 = time.perf_counter()
    for _ in range(steps):
        run(img, metas)
    dt = (time.perf_counter() - t0) / steps
    return n / dt, dt, kind


def cpu_sample_text(kind, n, steps, dt):
    what = ('the unmodified reference TopDown.forward_test (baseline/_ref, torch fp32 eager CPU + its NumPy/cv2 decode)'
            if kind == 'reference' else 'torch fp32 eager + numpy decode (oracle port of the reference forward_test)')
    return f'{n} crops/step x {steps} step(s) ({dt:.1f} s/step), {what}'


def run_reference(args, cfg, K):
    rank = int(os.environ.get('RANK', 0))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    sd = synthetic.scaled_init_state_dict(cfg, 0)
    # bounded sample: calibrate on 4 crops, then size each step for ~10 s of CPU work
    rate, _, _ = cpu_crops_per_sec(cfg, sd, 4, K, threads, steps=1, warmup=1)
    n = int(max(4, min(args.crops, rate * 10)))
    steps = max(1, min(args.steps, 6))
    cps, dt, kind = cpu_crops_per_sec(cfg, sd, n, K, threads, steps=steps, warmup=0)
    line = dict(metric=METRIC, value=cps, unit='crops/s', impl='reference', n_gpus=args.gpus, steps=steps,
                warmup=args.warmup, ms_per_step=dt * 1e3, higher_is_better=True, scaling='weak', vs_baseline=None,
                dtype='f32', data='synthetic',
                config=dict(workload=args.workload, crops_per_step=n, flip_test=True,
                            decode=cfg['test_cfg'].get('use_udp') and 'udp_dark' or 'default'),
                cpu_baseline=dict(value=cps, unit='crops/s', cores=threads, kind=kind,
                                  sample=cpu_sample_text(kind, n, steps, dt)),
                e2e=dict(value=cps, unit='crops/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line))


# ---- B200 arm ----------------------------------------------------------------------------------------------------
class Ctx:
    def __init__(self):
        import torch.distributed as dist
        self.dist = dist
        self.world = int(os.environ.get('WORLD_SIZE', 1))
        self.rank = int(os.environ.get('RANK', 0))
        self.local_rank = int(os.environ.get('LOCAL_RANK', 0))
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device('cuda', self.local_rank)
        self.sampler = ClockSampler(self.local_rank)   # started now: nvidia-smi needs ~1 s before its first sample
        if self.rank == 0:
            self.sampler.start()
        # NCCL prints its version banner on stdout at the first collective; keep stdout for the one JSON line
        sys.stdout.flush()
        self.saved_stdout = os.dup(1)
        os.dup2(2, 1)
        if self.world > 1:
            dist.init_process_group('nccl', device_id=self.dev)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(self, *vals):
        t = torch.tensor(vals, device=self.dev, dtype=torch.float64)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return t.tolist()

    def restore_stdout(self):
        sys.stdout.flush()
        os.dup2(self.saved_stdout, 1)
        os.close(self.saved_stdout)

    def finish(self):
        self.sampler.stop()
        if self.world > 1:
            self.dist.barrier()
            self.dist.destroy_process_group()


def measure_inference(ctx, workload, n, steps, warmup, profile=False, e2e_steps=None):
    """One inference workload, n crops per GPU: device-resident timing (CUDA events, max over ranks) and the
    end-to-end timing through TopDown.forward_test from pinned host memory. Returns a dict (rank 0 uses it)."""
    import vitpose_b200 as V
    from vitpose_b200 import _lib, parallel
    from vitpose_b200.core.post_processing import flip_index_from_pairs
    from vitpose_b200.engine import decode_mode_from_cfg

    cfg = configs.baseline_model_cfg(workload)
    K = cfg['keypoint_head']['out_channels']
    dev, world, rank = ctx.dev, ctx.world, ctx.rank
    sd = synthetic.scaled_init_state_dict(cfg, 0)
    model = V.build_posenet(cfg)
    model.load_state_dict(sd, strict=True)
    model = model.cuda().eval()
    test_cfg = cfg['test_cfg']
    # two distinct pinned host batches (ping-pong) so no step re-reads the previous step's input; one when a single
    # batch is already several times the 126 MB L2
    nbuf = 2 if n * 3 * 256 * 192 * 4 < (512 << 20) else 1
    host = [synthetic.synthetic_crops(n, seed=rank * 7 + i).pin_memory() for i in range(nbuf)]
    metas = synthetic.synthetic_metas(n, K, seed=rank)
    dev_img = [h.to(dev) for h in host]
    eng = model._engine()
    flip_index = torch.from_numpy(flip_index_from_pairs(K, metas[0]['flip_pairs'])).to(dev)
    center = torch.from_numpy(np.stack([m['center'] for m in metas])).to(dev)
    scale = torch.from_numpy(np.stack([m['scale'] for m in metas])).to(dev)
    mode = decode_mode_from_cfg(test_cfg)

    def device_step(i):
        hm, _ = eng.forward_heatmaps(dev_img[i % nbuf], flip=True)
        r = eng.decode(hm, n, True, flip_index, bool(test_cfg.get('shift_heatmap', False)), mode,
                       test_cfg.get('modulate_kernel', 11), bool(test_cfg.get('use_udp', False)), center, scale)
        out = torch.cat([r['preds'], r['maxvals']], dim=2)
        return parallel.gather_contiguous(out)        # the path's only collective: final result gather

    for i in range(warmup):
        device_step(i)
    ctx.barrier()
    # ---- timed region (device-resident inputs) -------------------------------------------------------
    L = _lib.lib()
    launches0 = L.vpb_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ctx.barrier()
    m0 = ctx.sampler.mark()
    ev0.record()
    for i in range(steps):
        device_step(i)
    ev1.record()
    ctx.barrier()
    m1 = ctx.sampler.mark()
    ms = ev0.elapsed_time(ev1) / steps
    launches = L.vpb_launch_count() - launches0
    clocks = ctx.sampler.region(m0, m1) if rank == 0 else None
    # ---- per-kernel durations: the SAME K steps once more, right away, with a CUDA-event pair around every launch ----
    # (on the launching stream, vpb_profile_enable). Recording the events inside the timed region above cost 0.3-0.6 ms
    # per step — an event between two launches takes the programmatic-dependent-launch overlap away (same-box A/B,
    # tools/profile_overhead.py: 20.61 / 20.96 ms without, 21.20 / 21.26 ms with) — so `value` is timed without them and
    # the roofline / step shares come from this second pass, whose own step time is reported next to them (it can be
    # lower or higher than the timed one: the power-capped SM clock drifts by a few percent between the two passes).
    records, ms_events = [], None
    if profile:
        L.vpb_profile_enable(1)
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ctx.barrier()
        p0.record()
        for i in range(steps):
            device_step(i)
        p1.record()
        ctx.barrier()
        ms_events = p0.elapsed_time(p1) / steps
        records = _lib.profile_records()
        L.vpb_profile_enable(0)

    # ---- end to end through the reference-facing API: pinned host crops -> host keypoints ----------------
    e2e_steps = e2e_steps or steps
    for i in range(2):
        model(img=host[i % nbuf], img_metas=metas, return_loss=False)
    ctx.barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        res = model(img=host[i % nbuf], img_metas=metas, return_loss=False)
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) / e2e_steps * 1e3
    assert res['preds'].shape == (n, K, 3)
    ms, e2e_ms = ctx.max_over_ranks(ms, e2e_ms)
    total = n * world
    out = dict(workload=workload, crops_per_gpu=n, global_crops=total, ms=ms, e2e_ms=e2e_ms,
               value=total / (ms / 1e3), e2e_value=total / (e2e_ms / 1e3), clocks=clocks, launches=int(launches),
               records=records, ms_events=ms_events, mode=mode, cfg=cfg, sd=sd, K=K,
               h2d=int(host[0].numel() * 4 + n * 16), d2h=int(n * K * 3 * 4))
    del model, eng, dev_img, host
    torch.cuda.empty_cache()
    return out


def measure_train(ctx, workload, n, steps, warmup):
    """BASELINE.json configs[4]: one step = TopDown.forward_train -> loss.backward() (gradient all-reduce over NCCL
    issued inside, overlapped, when N > 1) -> grad-norm clip + layer-decay AdamW; per-GPU batch fixed."""
    import vitpose_b200 as V
    from vitpose_b200 import _lib
    from vitpose_b200.optim import LayerDecayOptimizerConstructor

    cfg = configs.baseline_model_cfg(workload)
    K = cfg['keypoint_head']['out_channels']
    dev, world, rank = ctx.dev, ctx.world, ctx.rank
    drop = float(cfg['backbone'].get('drop_path_rate', 0.0))      # the config's stochastic depth (ViTPose-B: 0.3)
    sd = synthetic.scaled_init_state_dict(cfg, 0)
    model = V.build_posenet(cfg)
    model.load_state_dict(sd, strict=True)
    model = model.cuda().train()
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
        return out

    # Device-resident number: the logged values stay 0-dim device tensors (TopDown.log_vars_on_device, the documented
    # switch of detectors/top_down.py) — nothing is read back inside the timed region. The end-to-end loop below runs the
    # default API: train_step returns Python floats (one stacked device->host read inside train_step, as base.py:66-74's
    # .item() calls do), which is the step's result on the host.
    model.log_vars_on_device = True
    for i in range(warmup):
        step(devb[i & 1])
    ctx.barrier()
    calls0 = _lib.ABI_CALLS[0]
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ctx.barrier()
    m0 = ctx.sampler.mark()
    ev0.record()
    for i in range(steps):
        step(devb[i & 1])
    ev1.record()
    ctx.barrier()
    m1 = ctx.sampler.mark()
    ms = ev0.elapsed_time(ev1) / steps
    calls = _lib.ABI_CALLS[0] - calls0
    clocks = ctx.sampler.region(m0, m1) if rank == 0 else None
    # end to end: pinned host batch -> device every step, loss value back on the host
    # (as a DataLoader with pin_memory + a prefetching iterator does: the copy of batch i + 1 runs on a side stream
    # while step i computes; every step still pays its own H2D copy inside the timed region, the first one exposed)
    ctx.barrier()
    model.log_vars_on_device = False
    copy_stream = torch.cuda.Stream(dev)
    main_stream = torch.cuda.current_stream(dev)

    # Two resident device batches, refilled in turn by the copy stream (no per-step allocation: fresh tensors on a side
    # stream + record_stream made the caching allocator fall back to cudaMalloc / cudaFree in some runs: 21 ms per step
    # instead of 11.7). A buffer is overwritten only after the step that read it has finished (event on the main stream).
    dev_bufs = [tuple(torch.empty(t.shape, dtype=t.dtype, device=dev) for t in host[0]) for _ in range(2)]
    consumed = [None, None]

    def upload(k, hb):
        with torch.cuda.stream(copy_stream):
            if consumed[k] is not None:
                copy_stream.wait_event(consumed[k])
            for d, h in zip(dev_bufs[k], hb):
                d.copy_(h, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return dev_bufs[k], ev

    prefetch = os.environ.get('VPB_BENCH_PREFETCH', '1') != '0'      # 0: copy on the compute stream (A/B)
    def e2e_loop(count):
        nxt = upload(0, host[0]) if prefetch else None
        for i in range(count):
            k = i & 1
            if prefetch:
                batch, ev = nxt
                main_stream.wait_event(ev)
                if i + 1 < count:
                    nxt = upload(k ^ 1, host[k ^ 1])
            else:
                for d, h in zip(dev_bufs[k], host[k]):
                    d.copy_(h, non_blocking=True)
                batch = dev_bufs[k]
            out = step(batch)
            done = torch.cuda.Event()
            done.record(main_stream)
            consumed[k] = done
            loss_host = out['log_vars']['loss']        # a Python float: read back inside train_step
            assert isinstance(loss_host, float)
        torch.cuda.synchronize()
        return loss_host

    out_keys = step(devb[0])['log_vars']          # heatmap_loss, acc_pose, loss: the floats a step returns
    e2e_loop(2)                        # untimed: the copy stream's allocator pool, as the 2 warm-up calls of inference
    ctx.barrier()
    t0 = time.perf_counter()
    loss_host = e2e_loop(steps)
    e2e_ms = (time.perf_counter() - t0) / steps * 1e3
    ms, e2e_ms = ctx.max_over_ranks(ms, e2e_ms)
    total = n * world
    out = dict(workload=workload + '-train', crops_per_gpu=n, global_crops=total, ms=ms, e2e_ms=e2e_ms,
               value=total / (ms / 1e3), e2e_value=total / (e2e_ms / 1e3), clocks=clocks, launches=int(calls),
               final_loss=float(loss_host), drop_path=drop,
               h2d=int(sum(t.numel() * 4 for t in host[0])), d2h=4 * len(out_keys))
    del model, opt, devb, host
    torch.cuda.empty_cache()
    return out


def train_record(r, world, pk, steps, warmup, full=False):
    gf = GFLOP_PER_CROP['B-classic-17'] * 3          # forward + dgrad + wgrad
    tf = r['value'] * gf / 1e3 / world
    rec = dict(metric=TRAIN_METRIC, value=r['value'], unit='crops/s', n_gpus=world, steps=steps, warmup=warmup,
               ms_per_step=r['ms'], higher_is_better=True, scaling='weak', dtype='bf16', data='synthetic',
               config=dict(workload=r['workload'], crops_per_gpu=r['crops_per_gpu'], global_crops=r['global_crops'],
                           optimizer='AdamW lr 5e-4 wd 0.1, layer decay 0.75, grad clip 1.0',
                           drop_path=r['drop_path'], parallelism=f'dp{world}',
                           collective='gradient all-reduce (NCCL, overlapped with backward)' if world > 1 else None,
                           l2='activations per step >> 126 MB L2; inputs ping-pong between two buffers',
                           log_vars='value: left on the device (TopDown.log_vars_on_device), nothing read back in the '
                                    'timed region; e2e: Python floats from train_step (default API, one device->host '
                                    'read per step inside train_step)'),
               roofline=dict(bound='tensor', kernel='whole training step (3 x forward GEMM FLOPs)', achieved=tf,
                             peak=pk['tf_sustained'], unit='TFLOP/s', frac=tf / pk['tf_sustained'], traffic=None,
                             peak_source=f"{pk['source']} sustained bf16"),
               clocks=r['clocks'], gpu_launches=r['launches'], final_loss=r['final_loss'],
               e2e=dict(value=r['e2e_value'], unit='crops/s', h2d_bytes_per_step=r['h2d'], d2h_bytes_per_step=r['d2h'],
                        ms_per_step=r['e2e_ms'], api='TopDown.train_step + loss.backward() + LayerDecayAdamW.step'))
    if full:
        rec['vs_baseline'] = None
    return rec


def inference_record(r, world, pk, steps, warmup, scaling):
    gf = GFLOP_PER_CROP[r['workload']] * 2
    return dict(metric=METRIC, value=r['value'], unit='crops/s', n_gpus=world, steps=steps, warmup=warmup,
                ms_per_step=r['ms'], higher_is_better=True, scaling=scaling, dtype='bf16', data='synthetic',
                config=dict(workload=r['workload'], crops_per_gpu=r['crops_per_gpu'], global_crops=r['global_crops'],
                            flip_test=True, decode={3: 'udp_dark', 2: 'unbiased', 1: 'default', 0: 'none'}[r['mode']],
                            parallelism=f'dp{world}', l2='activations per step >> 126 MB L2',
                            weights='random scaled-init (no checkpoints offline)'),
                model_tflops=r['value'] * gf / 1e3,
                model_frac_of_peak=r['value'] * gf / 1e3 / world / pk['tf_sustained'],
                clocks=r['clocks'], gpu_launches=r['launches'],
                e2e=dict(value=r['e2e_value'], unit='crops/s', h2d_bytes_per_step=r['h2d'],
                         d2h_bytes_per_step=r['d2h'], ms_per_step=r['e2e_ms'],
                         api='TopDown.forward_test(img=<pinned host fp32>, img_metas=...)'))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--workload', default='B-classic-17', choices=sorted(configs.BASELINE_CONFIGS))
    ap.add_argument('--crops', type=int, default=0, help='crops per GPU per step (default: BASELINE batch)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-extra', action='store_true', help='skip the configs[2..4] sub-records')
    ap.add_argument('--train', action='store_true', help='measure the training step (BASELINE configs[4]) instead')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'b200' else args.warmup
    default_crops = args.crops <= 0
    if default_crops:
        args.crops = DEFAULT_CROPS[args.workload]
    cfg = configs.baseline_model_cfg(args.workload)
    K = cfg['keypoint_head']['out_channels']
    if args.impl == 'reference':
        run_reference(args, cfg, K)
        return
    pk = peaks()
    ctx = Ctx()
    world, rank = ctx.world, ctx.rank
    if args.train:
        r = measure_train(ctx, args.workload, 64 if default_crops else args.crops, args.steps, args.warmup)
        ctx.restore_stdout()
        if rank == 0:
            print(json.dumps(train_record(r, world, pk, args.steps, args.warmup, full=True)))
        ctx.finish()
        return

    n = args.crops
    r = measure_inference(ctx, args.workload, n, args.steps, args.warmup, profile=True)
    line = None
    if rank == 0:
        # dominant kernel: the transformer GEMMs (fc1 carries the largest share); per-launch average inside the step
        by_tag = {}
        for tag, v in r['records']:
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
        shares = {k: round(float(np.sum(v)) / args.steps / r['ms_events'], 4) for k, v in by_tag.items()}
        per_launch_ms = {k: round(float(np.mean(v)), 4) for k, v in by_tag.items()}
        # DRAM bytes of one fc1 launch: NOT measured by this run — read from the committed ncu --set full capture of
        # this same command (B workload, 256 crops), see traffic_source
        traffic, traffic_source = None, None
        tpath = os.path.join(ROOT, 'profiles', 'r02_fc1_traffic.json')
        if args.workload == 'B-classic-17' and n == 256 and os.path.exists(tpath):
            traffic = json.load(open(tpath))['traffic_bytes']
            traffic_source = ('constant from profiles/r02_fc1_traffic.json (ncu --set full capture of the fc1 launch '
                              'of `bench.py`, dram__bytes_read.sum + dram__bytes_write.sum); not re-measured live')
        roofline = dict(bound='tensor', kernel='gemm_bf16_tn_kernel<256,GELU> (mlp.fc1)', achieved=achieved,
                        peak=pk['tf_sustained'], unit='TFLOP/s', frac=achieved / pk['tf_sustained'], traffic=traffic,
                        traffic_source=traffic_source,
                        peak_source=f"{pk['source']} sustained bf16 (kernel timed inside a long step)",
                        frac_of_burst=achieved / pk['tf_burst'],
                        transformer_gemms_tflops=tgemm_tf, transformer_gemms_frac=tgemm_tf / pk['tf_sustained'],
                        transformer_gemms_frac_of_burst=tgemm_tf / pk['tf_burst'],
                        step_share_by_kernel=shares, ms_per_launch=per_launch_ms,
                        kernel_timing=('CUDA-event pair around every launch on the launching stream, recorded in a '
                                       'second pass of the same %d steps right after the timed region (events inside '
                                       'it cost 0.3-0.6 ms per step: no launch overlap across an event)' % args.steps),
                        ms_per_step_with_events=r['ms_events'])
        line = inference_record(r, world, pk, args.steps, args.warmup, 'weak')
        line['vs_baseline'] = None
        line['roofline'] = roofline
        line['config']['l2'] = 'activations per step >> 126 MB L2; inputs ping-pong between two buffers'

    # ---- BASELINE configs[2..4] in the same process (sub-records; the headline above stays comparable) -----------
    if not args.no_extra and args.workload == 'B-classic-17' and default_crops:
        extra = {}
        n_l = max(1, 1024 // world)
        rl = measure_inference(ctx, 'L-simple-17', n_l, 3, 3, e2e_steps=2)
        rh = measure_inference(ctx, 'H-classic-133', 256, 3, 3, e2e_steps=2)
        rt = measure_train(ctx, 'B-classic-17', 64, max(5, min(args.steps, 10)), 3)
        if rank == 0:
            extra['L-simple-17@1024/N'] = inference_record(rl, world, pk, 3, 3, 'strong')
            extra['H-classic-133@256/GPU'] = inference_record(rh, world, pk, 3, 3, 'weak')
            extra['B-train@64/GPU'] = train_record(rt, world, pk, max(5, min(args.steps, 10)), 3)
            line['configs'] = extra
    ctx.restore_stdout()
    if rank == 0:
        if not args.no_cpu_baseline and world == 1:     # rank 0 at N = 1 only (other ranks would compete for the cores)
            threads = os.cpu_count() or 1
            rate, _, _ = cpu_crops_per_sec(r['cfg'], r['sd'], 4, K, threads, steps=1, warmup=1)
            ns = int(max(4, min(n, rate * 12)))
            cps, dt, kind = cpu_crops_per_sec(r['cfg'], r['sd'], ns, K, threads, steps=1, warmup=0)
            line['cpu_baseline'] = dict(value=cps, unit='crops/s', cores=threads, kind=kind,
                                        sample=cpu_sample_text(kind, ns, 1, dt))
        print(json.dumps(line))
    ctx.finish()


if __name__ == '__main__':
    main()
