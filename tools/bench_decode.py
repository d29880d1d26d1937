"""Decode kernel alone against the HBM roofline (CUDA events, inputs larger than L2 by rotating buffer sets).

Algorithmic bytes per map = H*W*4 for each pass read (2 passes with the flip test) + 12 bytes of results."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vitpose_b200 import _lib, ops  # noqa: E402


def time_decode(n, k, mode, shift, sets=6, iters=30):
    dev = torch.device('cuda:0')
    hms = [(torch.rand(n, k, 64, 48, device=dev), torch.rand(n, k, 64, 48, device=dev)) for _ in range(sets)]
    fi = torch.arange(k, device=dev, dtype=torch.int32)
    c, s = torch.rand(n, 2, device=dev), torch.rand(n, 2, device=dev) + 0.5
    for i in range(sets):
        ops.decode(hms[i][0], hms[i][1], fi, shift, mode, 11, True, c, s)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(iters):
        h = hms[i % sets]
        ops.decode(h[0], h[1], fi, shift, mode, 11, True, c, s)
    b.record()
    torch.cuda.synchronize()
    us = a.elapsed_time(b) / iters * 1e3
    nbytes = n * k * (2 * 64 * 48 * 4 + 12)
    return dict(us=round(us, 1), gbs=round(nbytes / us / 1e3, 1), bytes=nbytes)


def main():
    peaks = json.load(open(os.path.join(os.path.dirname(__file__), '..', 'MEASURED_PEAKS.json')))
    hbm = float(peaks.get('hbm_gbs', peaks.get('hbm_copy_gbs', 6551.0))) if isinstance(peaks, dict) else 6551.0
    out = {}
    for name, n, k, mode, shift in (('H133_default_shift', 256, 133, _lib.DECODE_DEFAULT, True),
                                    ('H133_udp', 256, 133, _lib.DECODE_UDP_DARK, False),
                                    ('B17x1024_udp', 1024, 17, _lib.DECODE_UDP_DARK, False),
                                    ('B17x256_udp', 256, 17, _lib.DECODE_UDP_DARK, False),
                                    ('H133_unbiased', 256, 133, _lib.DECODE_UNBIASED, False)):
        r = time_decode(n, k, mode, shift)
        r['hbm_frac'] = round(r['gbs'] / hbm, 3)
        out[name] = r
        print(name, json.dumps(r))
    os.makedirs('gpurun_out', exist_ok=True)
    json.dump(dict(hbm_peak_gbs=hbm, results=out), open('gpurun_out/decode_roofline.json', 'w'), indent=1)


if __name__ == '__main__':
    main()
