#!/usr/bin/env python
"""Device-memory microbench used to sanity-check the roofline denominators:
write-only (fill), read-only (sum), copy bandwidth of plain torch ops on 2 GiB,
CUDA events, L2-sized working sets excluded."""
import json
import torch


def t(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(n):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); e.synchronize()
        best = min(best, s.elapsed_time(e))
    return best


def main():
    n = 1 << 29                                    # 2 GiB of fp32
    a = torch.empty(n, device="cuda")
    b = torch.empty(n, device="cuda")
    a.normal_()
    out = {"bytes": n * 4}
    ms = t(lambda: a.zero_());           out["memset_write_gbs"] = n * 4 / ms / 1e6
    ms = t(lambda: a.fill_(1.5));        out["fill_write_gbs"] = n * 4 / ms / 1e6
    ms = t(lambda: b.copy_(a));          out["copy_rw_gbs"] = 2 * n * 4 / ms / 1e6
    ms = t(lambda: a.sum());             out["sum_read_gbs"] = n * 4 / ms / 1e6
    ms = t(lambda: torch.mul(a, 2.0, out=b)); out["scale_rw_gbs"] = 2 * n * 4 / ms / 1e6
    print(json.dumps(out))


if __name__ == "__main__":
    main()
