#!/usr/bin/env python
"""BASELINE config 5: standalone fake-quant / min-max sweep, GB/s vs the HBM roofline.

    python bench_sweep.py [--max-log2 30] [--reps 20] [--out profiles/sweep.json]

Every kernel is timed with CUDA events on torch's current stream after warm-up; between timed
iterations a 512 MiB buffer is rewritten to flush the 126 MB L2 (unless --no-flush) and a second 512 MiB
buffer is then read, so that the ~100 MB of DIRTY lines the rewrite leaves in L2 are written back before the
clock starts instead of inside the timed kernel (which would add up to 100 MB of write traffic to a kernel
that moves 4 MB ... 16 GB: +37 % at 2^26 elements for a 4 B/element kernel).  GB/s is
ALGORITHMIC bytes / time: 4 B/elem range, 8 B/elem forward, 12 B/elem online,
4 B/elem histogram, 12 B/elem masked STE backward.
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from quantization.mxnet_b200 import ops  # noqa: E402


def peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


def timeit(fn, reps, flush):
    fn()
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        if flush is not None:
            flush[0].add_(1)          # rewrite 512 MiB: nothing of the tensor under test survives in L2
            flush[1].max()            # read 512 MiB: the dirty lines of the rewrite are evicted (written back) now
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b) * 1e-3)
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--min-log2", type=int, default=20)
    ap.add_argument("--max-log2", type=int, default=30)
    ap.add_argument("--step", type=int, default=2)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--no-flush", action="store_true")
    ap.add_argument("--out", default="")
    ap.add_argument("--kernels", default="")
    args = ap.parse_args()
    torch.cuda.set_device(0)
    peak, peak_kind = peak_gbs()
    flush = None if args.no_flush else (torch.zeros(128 * 1024 * 1024, dtype=torch.float32, device="cuda"),
                                        torch.zeros(128 * 1024 * 1024, dtype=torch.float32, device="cuda"))
    rows_out = []
    want = set(k for k in args.kernels.split(",") if k)
    for lg in range(args.min_log2, args.max_log2 + 1, args.step):
        n = 1 << lg
        g = torch.Generator(device="cuda").manual_seed(7)
        x = torch.randn(n, device="cuda", generator=g).abs_()
        y = torch.empty_like(x)
        qp = ops.scale_from_max(torch.tensor([3.0], device="cuda"), 8, False, ops.LO_ZERO)
        mx = torch.tensor([4.0], device="cuda")
        counts = torch.zeros(2049, dtype=torch.int64, device="cuda")
        cur = torch.empty(1, device="cuda")
        qp2 = torch.empty(4, device="cuda")
        s64 = torch.full((64,), 0.01, device="cuda")
        s1k = torch.full((1024,), 0.01, device="cuda")
        rowbuf = torch.empty(1024, device="cuda")
        cases = {
            "torch_copy": (8, lambda: y.copy_(x)),
            "absmax_layer": (4, lambda: ops.absmax_rows(x, 1, out=rowbuf[:1])),
            "absmax_rows1024": (4, lambda: ops.absmax_rows(x, 1024, out=rowbuf)),
            "input_range_n128": (4, lambda: ops.input_range(x, 128, cur_max=cur)),
            "minmax": (4, lambda: ops.minmax(x, out=rowbuf[:2])),
            "fwd_scalar_u8": (8, lambda: ops.forward_scalar(x, qp, out=y)),
            "fwd_rows64": (8, lambda: ops.forward_rows(x, s64, out=y)),
            "fwd_rows1024": (8, lambda: ops.forward_rows(x, s1k, out=y)),
            "fwd_online_n128": (12, lambda: ops.forward_online(x, 8, False, ops.LO_ZERO, n_samples=128, out=y, cur_max=cur, qparams=qp2)),
            "fwd_offline_track_n128": (8, lambda: ops.forward_online(x, 8, False, ops.LO_ZERO, input_max=mx, n_samples=128, out=y, cur_max=cur, qparams=qp2)),
            "ste_mask": (12, lambda: ops.ste_backward(x, y, qp, mode=ops.STE_CLIP_MASK)),
            "hist2048": (4, lambda: ops.hist_nonzero(x, mx, 2048, counts)),
        }
        for name, (bpe, fn) in cases.items():
            if want and name not in want:
                continue
            med, best = timeit(fn, args.reps, flush)
            rec = {"kernel": name, "log2n": lg, "bytes_per_elem": bpe, "median_us": med * 1e6, "best_us": best * 1e6,
                   "gbs_median": bpe * n / med / 1e9, "gbs_best": bpe * n / best / 1e9,
                   "frac_of_%s_peak" % peak_kind: bpe * n / med / 1e9 / peak, "frac_of_8000": bpe * n / med / 1e9 / 8000.0}
            rows_out.append(rec)
            print(json.dumps(rec), flush=True)
        del x, y
    if args.out:
        os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
        json.dump({"peak_gbs": peak, "peak_kind": peak_kind, "l2_flush": not args.no_flush, "rows": rows_out},
                  open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
