#!/usr/bin/env python
"""Where does one training iteration go?  Runs the bench's train step under torch.profiler
(CUPTI kernel records, no ncu replay) and prints, per image: wall time, summed device time of
all kernels (= GPU-busy time if nothing overlapped), launch count, and the top kernels.

    python tools/step_profile.py [--images 8] [--no-materialize] [--graph] > gpurun_out/step_profile.md
"""
import argparse
import collections
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import bench  # noqa: E402
from wesup_b200 import synth  # noqa: E402
from wesup_b200.models import initialize_trainer  # noqa: E402
from wesup_b200.utils.metrics import accuracy, dice  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=8)
    ap.add_argument("--no-materialize", action="store_true")
    ap.add_argument("--no-pool-first", action="store_true")
    ap.add_argument("--graph", action="store_true")
    ap.add_argument("--host", action="store_true", help="also print where the host time goes")
    ap.add_argument("--size", type=int, nargs=2, default=[bench.H, bench.W])
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    h, w = a.size
    trainer = initialize_trainer("wesup", device=dev, pretrained=False, materialize_hypercolumn=not a.no_materialize,
                                 pool_first=not a.no_pool_first, cuda_graph=a.graph)
    trainer.optimizer, _ = trainer.get_default_optimizer()
    trainer.metric_funcs = [accuracy, dice]
    pool = 4
    data = [tuple(t.to(dev) for t in synth.sample(h, w, index=i)) for i in range(pool)]

    def run(n):
        for k in range(n):
            trainer.prefetch(*data[(k + 1) % pool])
            trainer.train_one_iteration("train", *data[k % pool])
        trainer.flush_metrics()

    run(6)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    run(a.images)
    enqueue = (time.perf_counter() - t0) / a.images * 1e3      # host time to enqueue (includes the per-image host waits)
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / a.images * 1e3
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        run(a.images)
        torch.cuda.synchronize()
    per = collections.defaultdict(lambda: [0.0, 0])
    total, launches = 0.0, 0
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA and ev.device_time > 0:
            per[ev.name][0] += ev.device_time
            per[ev.name][1] += 1
            total += ev.device_time
            launches += 1
    n = a.images
    print(f"# step profile {h}x{w}, {n} images, materialize={not a.no_materialize} pool_first={not a.no_pool_first}\n")
    print(f"host enqueue {enqueue:.3f} ms/image; wall (unprofiled) {wall:.3f} ms/image; summed device time {total / n / 1e3:.3f} ms/image; "
          f"{launches / n:.0f} device activities/image\n")
    print("| us/image | share | n/image | kernel |\n|---:|---:|---:|---|")
    for name, (us, cnt) in sorted(per.items(), key=lambda kv: -kv[1][0])[:45]:
        print(f"| {us / n:.1f} | {100 * us / total:.1f} % | {cnt / n:.1f} | `{name[:110]}` |")


    if a.host:
        print("\n## host side (self CPU time per image)\n\n| us/image | calls/image | op |\n|---:|---:|---|")
        for ev in sorted(prof.key_averages(), key=lambda e: -e.self_cpu_time_total)[:40]:
            print(f"| {ev.self_cpu_time_total / n:.1f} | {ev.count / n:.1f} | `{ev.key[:90]}` |")
        import cProfile
        import pstats
        pr = cProfile.Profile()
        pr.enable()
        run(a.images)
        torch.cuda.synchronize()
        pr.disable()
        print("\n## cProfile (cumulative, python frames)\n\n```")
        import io
        buf = io.StringIO()
        pstats.Stats(pr, stream=buf).sort_stats("cumulative").print_stats(45)
        print(buf.getvalue()[:9000])
        print("```")


if __name__ == "__main__":
    main()
