"""A/B: does recording CUDA events around every launch (vpb_profile_enable, what bench.py does in its timed region to
get per-kernel durations) change the step time?  python tools/profile_overhead.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

ctx = bench.Ctx()
out = []
for prof in (False, True, False, True):
    r = bench.measure_inference(ctx, 'B-classic-17', 256, 10, 3, profile=prof, e2e_steps=2)
    out.append(f'profile events {"on " if prof else "off"}: {r["ms"]:.3f} ms per step, {r["value"]:.0f} crops/s')
ctx.restore_stdout()
print('\n'.join(out))
ctx.finish()
