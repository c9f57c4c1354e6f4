#!/usr/bin/env python
"""Headline benchmark: KL calibration of MobileNet-1.0 (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

One *step* = one calibration batch of 128 synthetic 224x224 images: the 27 quantised-layer inputs
(639,172,608 fp32 elements, 2.56 GB -- far larger than the 126 MB L2, so no flush is needed) are
histogrammed into 2048(+1) bins and folded into the running float32 histograms.  A calibration is
K such steps followed by ONE KL threshold search over the 27 layers and the threshold update; that
closing search runs INSIDE the timed region.  The K-step calibration is timed REPEATS times back to back
(each bracketed by a barrier + synchronize, CUDA events, max over ranks) and the MEDIAN repetition is
reported, so that one scheduler hiccup in an 18 ms region cannot move the headline.  ``value`` times this
with the layer inputs resident in HBM; ``e2e`` times the user-facing calls (``collect_feature_maps`` +
``kl_calibrate_all``) fed from pinned host images, network forward, H2D and D2H included.

The same JSON line also carries
  parity   cross-rank equality of the histograms / chosen bins (hash all-reduced with MIN and MAX), equality of a
           data-parallel calibration with a single-process replay over the same shards, and the data-parallel
           fake-BN statistics against the single-process ones;
  sweep    BASELINE config 5 in brief: every north-star kernel at 2^28 and 2^30 elements, L2-flushed, median of 15;
  configs  BASELINE configs 1, 3 (notebook converters) and 4 through the drop-in API at this N (bench_configs.py).

``--impl reference`` times the reference's own CPU path for the same work on a bounded sample: the reference's
quantize/distribution_calibrate.py itself when build() has staged it under oracle/_ref/ (kind "reference"), else
the oracle port (kind "port"); all host cores, same amortisation as the GPU arm (one KL search per K x 128 images).
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BINS, LEVELS = 2048, 256
RING = 8                             # batches per sum-all-reduce of the integer counts
BATCH = 128
REPEATS = 5                          # timed K-step calibrations; the median one is reported
MODEL = "mobilenet1.0"
METRIC = "kl_calibration_images_per_sec"
N_LAYERS = 27
ELEMS_PER_IMAGE = 4_993_536          # sum of the 27 layer-input sizes for one 224x224 image


def workload_config(n_gpus):
    return {
        "workload": "mobilenet1.0 KL calibration: per-channel int8 weights, offline uint8 inputs, 2048-bin "
                    "histograms of the 27 quantised-layer inputs of a synthetic 128x3x224x224 batch per GPU, "
                    "one KL threshold search (levels=256) closing every K-step calibration inside the timed region",
        "batch_per_gpu": BATCH, "bins": BINS, "levels": LEVELS, "layers": N_LAYERS,
        "elements_per_step_per_gpu": ELEMS_PER_IMAGE * BATCH,
        "l2": "inputs (2.56 GB per step) exceed the 126 MB L2; no flush needed",
        "parallelism": "batch sharded over %d GPU(s); NCCL max-all-reduce of first-batch ranges, one asynchronous "
                       "sum-all-reduce of the 32-bit counts per %d steps (float32 adds replayed per step in batch "
                       "order)" % (n_gpus, RING),
    }


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,clocks.mem,power.draw")

    def __init__(self, index=0):
        self.index = index
        self.samples = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.time(), line.strip()))

    def stop(self):
        if self.proc is None:
            return
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        self.proc = None

    def summary(self, t0=None, t1=None):
        """Median clocks / throttle reasons of the samples read between host times t0 and t1 (all if None).  A
        sample is what nvidia-smi saw during the 100 ms before it was read."""
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi produced no sample"], "samples": 0}
        sm, mx, mem, pw, reasons = [], [], [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ts, s in self.samples:
            if (t0 is not None and ts < t0) or (t1 is not None and ts > t1 + 0.12):
                continue
            parts = [p.strip() for p in s.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
            try:
                mem.append(float(parts[6]))
                pw.append(float(parts[7]))
            except (ValueError, IndexError):
                pass
        sm.sort()
        mem.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "mem_mhz": mem[len(mem) // 2] if mem else None,
                "power_w_max": max(pw) if pw else None}


# ------------------------------------------------------------------------------------------------
# CPU side (the reference's own file, or the oracle port): only used for cpu_baseline and --impl reference
# ------------------------------------------------------------------------------------------------
def layer_shapes():
    """(C, H, W) of the 27 quantised-layer inputs of mobilenet1.0 at 224x224 (first conv excluded)."""
    dw = [32, 64, 128, 128, 256, 256] + [512] * 6 + [1024]
    ch = [64, 128, 128, 256, 256, 512] + [512] * 5 + [1024, 1024]
    st = [1, 2, 1, 2, 1, 2, 1, 1, 1, 1, 1, 2, 1]
    hw = 112
    shapes = []
    for d, c, s in zip(dw, ch, st):
        shapes.append((d, hw, hw))          # depthwise input
        hw //= s
        shapes.append((d, hw, hw))          # pointwise input
    shapes.append((1024, 1, 1))             # dense input
    assert sum(c * h * w for c, h, w in shapes) == ELEMS_PER_IMAGE
    return shapes


REF_FILE = os.path.join(ROOT, "oracle", "_ref", "distribution_calibrate.py")
_REF_MOD = []


def reference_module():
    """The reference's quantize/distribution_calibrate.py, staged verbatim under oracle/_ref/ by build()
    (oracle/stage_ref.py; git-ignored, travels with the tree), or None."""
    if not _REF_MOD:
        mod = None
        if os.path.exists(REF_FILE):
            import importlib.util
            spec = importlib.util.spec_from_file_location("fq_reference_distribution_calibrate", REF_FILE)
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
        _REF_MOD.append(mod)
    return _REF_MOD[0]


_CPU_FMS = []        # per-layer sample activations, generated once in the parent and inherited by fork


def _cpu_layer_hist(args):
    i, fm_max = args
    ref = reference_module()
    if ref is not None:
        h, m = ref._discrete_histogram(_CPU_FMS[i], BINS, fm_max)      # distribution_calibrate.py:31-47, verbatim
        return h, m
    from oracle import fq_oracle as O
    return O.discrete_histogram(_CPU_FMS[i], BINS, fm_max)


def _cpu_layer_kl(hist):
    ref = reference_module()
    if ref is not None:
        return ref.kl_calibrate(hist, LEVELS, LEVELS, BINS)            # distribution_calibrate.py:117-171, verbatim
    from oracle import fq_oracle as O
    return O.kl_calibrate(hist, LEVELS, LEVELS, BINS)


def cpu_calibration(sample_images, steps, procs, warmup=0):
    """`steps` sample batches through the CPU path, then one KL search of the 27 layers.
    Returns (histogram seconds, KL seconds, best bins) of wall clock with `procs` worker processes; the
    synthetic layer inputs are generated before the clock starts (the GPU arm's are resident too)."""
    import multiprocessing as mp
    import numpy as np
    del _CPU_FMS[:]
    for i, shp in enumerate(layer_shapes()):
        r = np.random.RandomState(100 + i)
        _CPU_FMS.append(np.maximum(r.standard_normal((sample_images,) + shp), 0).astype(np.float32))
    hists, maxes = [0] * N_LAYERS, [None] * N_LAYERS
    hist_wall = 0.0
    with mp.get_context("fork").Pool(procs) as pool:
        pool.map(_cpu_layer_hist, [(N_LAYERS - 1, None)] * procs)     # import + page-in, untimed
        for s in range(warmup + steps):
            t0 = time.perf_counter()
            res = pool.map(_cpu_layer_hist, [(i, maxes[i]) for i in range(N_LAYERS)], chunksize=1)
            if s >= warmup:
                hist_wall += time.perf_counter() - t0
            for i, (h, m) in enumerate(res):
                if maxes[i] is None:
                    maxes[i] = m
                if s >= warmup:
                    hists[i] = hists[i] + h
        t0 = time.perf_counter()
        best = pool.map(_cpu_layer_kl, hists, chunksize=1)
        kl_wall = time.perf_counter() - t0
    return hist_wall, kl_wall, [int(b) for b in best]


def cpu_calibration_c(sample_images, steps):
    """The same work through the C + OpenMP restatement (oracle/fq_oracle.c): every layer uses all cores."""
    import numpy as np
    from oracle import build_c as C
    from oracle import fq_oracle as O
    C.lib()
    fms = []
    for i, shp in enumerate(layer_shapes()):
        r = np.random.RandomState(100 + i)
        fms.append(np.maximum(r.standard_normal((sample_images,) + shp), 0).astype(np.float32))
    hists, maxes = [0] * N_LAYERS, [None] * N_LAYERS
    hist_wall = 0.0
    for s in range(steps):
        t0 = time.perf_counter()
        for i, fm in enumerate(fms):
            if maxes[i] is None:
                maxes[i] = np.float32(fm.max())
            c = C.histogram_counts(fm, BINS, maxes[i], O.hist_scale(maxes[i], BINS, "nep50"))
            hists[i] = hists[i] + c[:BINS].astype(np.float32)
        hist_wall += time.perf_counter() - t0
    t0 = time.perf_counter()
    best = [C.kl_calibrate(h, LEVELS, LEVELS, BINS, "nep50")[0] for h in hists]
    kl_wall = time.perf_counter() - t0
    return hist_wall, kl_wall, best


def cpu_arm(K, steps, warmup, procs, sample_images):
    """The CPU path's images/s on the GPU arm's terms: histogram seconds per image measured on `steps` sample
    batches of `sample_images` images, ONE KL search of the 27 layers amortised over K batches of 128."""
    hist_wall, kl_wall, best = cpu_calibration(sample_images, steps, procs, warmup)
    per_image = hist_wall / (sample_images * steps)
    total_for_k = per_image * BATCH * K + kl_wall
    kind = "reference" if reference_module() is not None else "port"
    what = ("the reference's own quantize/distribution_calibrate.py (_discrete_histogram + kl_calibrate, executed "
            "verbatim from oracle/_ref/)" if kind == "reference" else
            "NumPy oracle port of distribution_calibrate.py (its vectorised KL is ~10x faster than the reference's "
            "pure-Python loop)")
    return {"value": BATCH * K / total_for_k, "unit": "images/s", "cores": procs, "kind": kind,
            "sample": "%s; 27 layers spread over %d process(es); histograms timed on %d step(s) of %d synthetic images "
                      "(%.3f s per image), one KL search of the 27 layers (%.2f s) amortised over K=%d batches of %d "
                      "as in the GPU arm" % (what, procs, steps, sample_images, per_image, kl_wall, K, BATCH),
            "hist_s_per_image": per_image, "kl_s": kl_wall, "ms_per_step": 1e3 * total_for_k / K,
            "best_bins_head": best[:4]}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    procs = os.cpu_count() or 1
    sample_images = 4
    steps = max(1, min(args.steps, 8))             # bounded: 8 x 4 images x 27 layers is ~10 s of CPU work
    cpu = cpu_arm(args.steps, steps, min(args.warmup, 1), procs, sample_images)
    # For transparency: the same work as hand-written C + OpenMP (oracle/fq_oracle.c), a much stronger CPU
    # implementation than the reference has.
    try:
        cpu_calibration_c(1, 1)
        ch, ck, _ = cpu_calibration_c(sample_images, 2)
        per = ch / (sample_images * 2)
        cpu["c_openmp"] = {"value": BATCH * args.steps / (per * BATCH * args.steps + ck), "unit": "images/s",
                           "cores": procs, "what": "C + OpenMP restatement (oracle/fq_oracle.c), same amortisation; "
                                                   "hist %.4f s/image, KL %.2fs" % (per, ck)}
    except Exception as e:       # the C checker is optional here
        cpu["c_openmp"] = {"unavailable": str(e)[:200]}
    value = cpu["value"]
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": cpu["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": cpu,
        "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------
# GPU side
# ------------------------------------------------------------------------------------------------
def build_net(device):
    import torch
    from torch import nn
    from quantization.mxnet_b200 import model_zoo as Z
    from quantization.mxnet_b200.quantize import convert
    from quantization.mxnet_b200.quantize.initialize import qparams_init
    torch.manual_seed(7)
    net = Z.get_model(MODEL, classes=1000).to(device).eval()
    fn = {nn.Conv2d: convert.gen_conv2d_converter(quant_type="channel"),
          nn.Linear: convert.gen_dense_converter(quant_type="channel"), nn.ReLU: None, nn.BatchNorm2d: None}
    convert.convert_model(net, exclude=Z.default_exclusions(net, MODEL), convert_fn=fn)
    qparams_init(net)
    return net


def capture_layer_inputs(net, X):
    import torch
    blocks = net.collect_quantized_blocks()
    acts, hooks = [], []
    for b in blocks:
        hooks.append(b.register_forward_hook(lambda m, x, y: acts.append(x[0].detach().contiguous())))
    with torch.no_grad():
        net(X)
    for h in hooks:
        h.remove()
    return acts


def tensor_hash(t):
    """Order-sensitive 62-bit hash of a tensor's bytes, computed on the device (int64 arithmetic wraps)."""
    import torch
    v = t.contiguous().view(torch.int32).reshape(-1).to(torch.int64)
    w = torch.arange(1, v.numel() + 1, device=v.device, dtype=torch.int64) * 0x9E3779B1 + 0x7F4A7C15
    return int(((v * w).sum() & ((1 << 62) - 1)).item())


def parity_block(world, rank, dev):
    """Small data-parallel calibration (2 global batches, 16 synthetic images per rank per batch, the 27 layer
    shapes of config 2) checked three ways: every rank ends with the same histograms and bins (hash all-reduced
    with MIN and MAX); rank 0 replays the same global batches in a single process (every rank's shard regenerated
    from its seed, counts of all shards summed in one slot) and must get the identical bits; and the fake-BN batch
    statistics combined from the ranks' records equal the single-process statistics of the concatenated batch."""
    import torch
    import torch.distributed as dist
    from quantization.mxnet_b200 import dist as fqdist
    from quantization.mxnet_b200 import ops
    n_img, n_batches = 16, 2
    shapes = layer_shapes()

    def shard(r, b):
        g = torch.Generator(device=dev).manual_seed(1000 + 97 * r + b)
        out = []
        for c, h, w in shapes:
            x = torch.randn(n_img, c, h, w, device=dev, generator=g)
            out.append(torch.relu_(x).mul_(1.0 + 0.5 * b))          # batch 1 exceeds batch 0's frozen max
        return out

    def calibrate(ranks, exchange):
        """`ranks`: whose shards this process histograms; `exchange`: all-reduce with the other processes."""
        hist = torch.zeros(N_LAYERS, BINS + 1, dtype=torch.float32, device=dev)
        minmax = torch.zeros(N_LAYERS, 2, dtype=torch.float32, device=dev)
        bad = torch.zeros(N_LAYERS, dtype=torch.int32, device=dev)
        ring = fqdist.CountsRing(N_LAYERS, BINS + 1, dev, slots=2, local=not exchange,
                                 accumulate=lambda c, first: ops.hist_accumulate(c.reshape(-1), hist.view(-1), first),
                                 max_count=n_img * max(c * h * w for c, h, w in shapes))
        for b in range(n_batches):
            if b == 0:              # the first global batch's max is frozen (distribution_calibrate.py:97-101)
                for r in ranks:
                    mm = torch.stack([ops.minmax(a) for a in shard(r, 0)])
                    minmax[:, 1] = torch.maximum(minmax[:, 1], mm[:, 1])
                if exchange and world > 1:
                    mx = minmax[:, 1].contiguous()
                    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
                    minmax[:, 1].copy_(mx)
            for r in ranks:         # counts of every shard of the global batch land in the same slot
                ops.hist_nonzero_multi(shard(r, b), minmax, 2, 1, BINS, ring.slot(), promotion="nep50", bad_flags=bad)
            ring.commit()
        ring.flush()
        margin = torch.empty(N_LAYERS, dtype=torch.float64, device=dev)
        best, _ = ops.kl_search(hist[:, :BINS].contiguous(), LEVELS, LEVELS, BINS, promotion="nep50", margin=margin)
        return hist, best, margin, bad

    hist, best, margin, bad = calibrate([rank], True)
    h_hist, h_best = tensor_hash(hist), tensor_hash(best)
    lo = torch.tensor([h_hist, h_best], dtype=torch.int64, device=dev)
    hi = lo.clone()
    if world > 1:
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    out = {"ranks_identical": bool(torch.equal(lo, hi)), "hist_hash": "%016x" % h_hist,
           "best_bins": [int(v) for v in best.cpu()], "min_kl_margin": float(margin.min()),
           "batches": n_batches, "images_per_rank_per_batch": n_img, "assert_flags_raised": int(bad.sum())}

    # fake-BN statistics: per-rank records -> all-gather -> finish, against one launch over the concatenated batch
    g = torch.Generator(device=dev)

    def conv_out(r):
        g.manual_seed(5000 + r)
        return torch.randn(32, 96, 28, 28, device=dev, generator=g) * 1.7 + torch.linspace(-3, 3, 96, device=dev).view(1, -1, 1, 1)
    rec = torch.empty(96, 4, dtype=torch.float64, device=dev)
    ops.channel_stats(conv_out(rank), parts=rec, finish=False)
    allrec = torch.empty(world, 96, 4, dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_gather_into_tensor(allrec.view(-1), rec.view(-1))
    else:
        allrec.copy_(rec.unsqueeze(0))
    dp_mean, dp_var = ops.channel_stats_finish(allrec)
    stats_hash = torch.tensor([tensor_hash(dp_mean), tensor_hash(dp_var)], dtype=torch.int64, device=dev)
    s_lo, s_hi = stats_hash.clone(), stats_hash.clone()
    if world > 1:
        dist.all_reduce(s_lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(s_hi, op=dist.ReduceOp.MAX)
    out["fake_bn_stats_ranks_identical"] = bool(torch.equal(s_lo, s_hi))

    if rank == 0:
        one_hist, one_best, _, _ = calibrate(list(range(world)), False)
        out["equals_single_process"] = bool(torch.equal(hist.view(torch.int32), one_hist.view(torch.int32)) and
                                            torch.equal(best, one_best))
        one_mean, one_var = ops.channel_stats(torch.cat([conv_out(r) for r in range(world)]))

        def ulps(a, b):
            return int((a.view(torch.int32).to(torch.int64) - b.view(torch.int32).to(torch.int64)).abs().max())
        out["fake_bn_stats_max_ulp_vs_single_process"] = {"mean": ulps(dp_mean, one_mean), "var": ulps(dp_var, one_var)}
    return out


def sweep_block(dev, peak):
    """BASELINE config 5 in brief (bench_sweep.py is the full sweep): each kernel at 2^28 and 2^30 fp32 elements,
    L2 flushed between iterations (512 MiB rewrite + 512 MiB read so the dirty lines are written back before the
    clock starts), CUDA events, median of 15.  GB/s = ALGORITHMIC bytes / time."""
    import torch
    from quantization.mxnet_b200 import ops
    flush = (torch.zeros(128 << 20, dtype=torch.float32, device=dev), torch.zeros(128 << 20, dtype=torch.float32, device=dev))
    rows = []
    for lg in (28, 30):
        n = 1 << lg
        g = torch.Generator(device=dev).manual_seed(7)
        x = torch.randn(n, device=dev, generator=g)
        y = torch.empty_like(x)
        codes = torch.empty(n, dtype=torch.int8, device=dev)
        mx3 = torch.tensor([3.0], device=dev)
        qp_u8 = ops.scale_from_max(mx3, 8, False, ops.LO_ZERO)
        qp_i4 = ops.scale_from_max(mx3, 4, True, ops.LO_NEG_MAX)
        qp_i16 = ops.scale_from_max(mx3, 16, True, ops.LO_NEG_MAX)
        mx4 = torch.tensor([4.0], device=dev)
        counts = torch.zeros(2049, dtype=torch.int64, device=dev)
        cur, qp2 = torch.empty(1, device=dev), torch.empty(4, device=dev)
        s64, s1k = torch.full((64,), 0.01, device=dev), torch.full((1024,), 0.01, device=dev)
        rowbuf = torch.empty(1024, device=dev)
        cases = [
            ("torch_copy (reference point)", 8, lambda: y.copy_(x)),
            ("K1 absmax per-layer", 4, lambda: ops.absmax_rows(x, 1, out=rowbuf[:1])),
            ("K1 absmax per-channel rows=1024", 4, lambda: ops.absmax_rows(x, 1024, out=rowbuf)),
            ("K1 input range (per-sample absmax + Kahan mean) N=128", 4, lambda: ops.input_range(x, 128, cur_max=cur)),
            ("K1 minmax", 4, lambda: ops.minmax(x, out=rowbuf[:2])),
            ("K2 forward scalar uint8", 8, lambda: ops.forward_scalar(x, qp_u8, out=y)),
            ("K2 forward scalar int4 signed", 8, lambda: ops.forward_scalar(x, qp_i4, out=y)),
            ("K2 forward scalar int16 signed", 8, lambda: ops.forward_scalar(x, qp_i16, out=y)),
            ("K2 forward scalar int4 + int8 codes", 9, lambda: _fwd_codes(ops, x, qp_i4, y, codes)),
            ("K2 forward rows=64 (per-group)", 8, lambda: ops.forward_rows(x, s64, out=y)),
            ("K2 forward rows=1024 (per-channel)", 8, lambda: ops.forward_rows(x, s1k, out=y)),
            ("K2 online (range + quantise) N=128", 12, lambda: ops.forward_online(x, 8, False, ops.LO_ZERO, n_samples=128, out=y, cur_max=cur, qparams=qp2)),
            ("K2 online int4 signed N=128", 12, lambda: ops.forward_online(x, 4, True, ops.LO_NEG_MAX, n_samples=128, out=y, cur_max=cur, qparams=qp2)),
            ("K2 offline + range tracking N=128", 8, lambda: ops.forward_online(x, 8, False, ops.LO_ZERO, input_max=mx4, n_samples=128, out=y, cur_max=cur, qparams=qp2)),
            ("K3 STE backward with clip mask", 12, lambda: ops.ste_backward(x, y, qp_u8, mode=ops.STE_CLIP_MASK)),
            ("K5 histogram 2048 bins", 4, lambda: ops.hist_nonzero(x, mx4, 2048, counts)),
            ("fake-BN channel statistics [128,64,.,.] one pass", 4, lambda: ops.channel_stats(x.view(128, 64, -1, 64), mean=rowbuf[:64], var=rowbuf[64:128])),
        ]
        for name, bpe, fn in cases:
            fn()
            fn()
            torch.cuda.synchronize()
            ts = []
            for _ in range(15):
                flush[0].add_(1)
                flush[1].max()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                fn()
                b.record()
                b.synchronize()
                ts.append(a.elapsed_time(b) * 1e-3)
            ts.sort()
            med = ts[len(ts) // 2]
            gbs = bpe * n / med / 1e9
            rows.append({"kernel": name, "log2n": lg, "bytes_per_elem": bpe, "median_us": round(med * 1e6, 1),
                         "gbs": round(gbs, 1), "frac_of_measured_peak": round(gbs / peak, 4),
                         "frac_of_8000": round(gbs / 8000.0, 4)})
        del x, y, codes
        torch.cuda.empty_cache()
    return {"peak_gbs": peak, "l2_flush": "512 MiB rewrite + 512 MiB read before every timed launch", "reps": 15,
            "stat": "median", "rows": rows}


def qconv_block(dev):
    """nn.Conv2D(quantized=True) on the int8 tensor cores (tcgen05, csrc/fq_qconv_mma.cu), two ResNet-like 3x3 layers
    at batch 32: GPU time of the implicit-GEMM kernel and of the whole layer (range + pack + GEMM), CUDA events around
    a graph replay of 20 calls, beside a plain cuDNN fp32 convolution of the same shape.  The roofline of the GEMM
    kernel is the tensor pipe: nominal dense int8 is 4.5 POP/s; MEASURED_PEAKS' cuBLAS bf16 figure x 2 is what a
    library GEMM reaches on this part."""
    import torch
    from quantization.mxnet_b200 import ops
    from quantization.mxnet_b200.nn import Conv2D

    def gpu_ms(fn, reps=20):
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(3):
                fn()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, capture_error_mode="thread_local"):      # other threads (NCCL watchdog) may call CUDA
            for _ in range(reps):
                fn()
        g.replay()
        torch.cuda.synchronize(dev)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        g.replay()
        b.record()
        b.synchronize()
        return a.elapsed_time(b) / reps

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    bf16 = float(json.load(open(peaks_path)).get("bf16_tflops", 0.0)) if os.path.exists(peaks_path) else 0.0
    rows = []
    saved_benchmark = torch.backends.cudnn.benchmark
    torch.backends.cudnn.benchmark = True             # the cuDNN comparison gets its autotuned algorithm
    for n, c, hw, co in ((32, 256, 56, 256), (32, 128, 28, 128)):
        conv = Conv2D(co, 3, 1, 1, in_channels=c, quantized=True, input_dtype="int8", weight_dtype="int8").to(dev)
        x = torch.randn(n, c, hw, hw, device=dev)
        with torch.no_grad():
            in_rng, unsigned, w_rng = conv._tensor_core_ranges(x)
            xq, s_in = ops.qconv_pack_input(x, in_rng, 1, 1)
            wq, s_w = conv._weight_codes(w_rng)
            t_mm = gpu_ms(lambda: ops.qconv_igemm(xq, wq, None, s_in, s_w, (1, 1), 1))
            t_layer = gpu_ms(lambda: conv(x))
            t_cudnn = gpu_ms(lambda: torch.nn.functional.conv2d(x, conv.weight, None, 1, 1))
        tops = 2.0 * n * hw * hw * co * c * 9 / (t_mm * 1e-3) / 1e12
        rows.append({"layer": "N%d %dx%dx%d -> %d, 3x3" % (n, c, hw, hw, co), "igemm_us": round(t_mm * 1e3, 1),
                     "igemm_tops_int8": round(tops, 1), "layer_us": round(t_layer * 1e3, 1),
                     "cudnn_fp32_conv_us": round(t_cudnn * 1e3, 1),
                     "roofline": {"bound": "tensor", "achieved": round(tops, 1), "peak": 4500.0, "unit": "TOP/s",
                                  "frac": round(tops / 4500.0, 3), "peak_kind": "nominal dense int8",
                                  "frac_of_2x_measured_bf16": round(tops / (2 * bf16), 3) if bf16 else None}})
        del conv, x, xq, wq
    torch.backends.cudnn.benchmark = saved_benchmark
    return {"rows": rows, "timing": "CUDA events around a CUDA-graph replay of 20 calls"}


def _fwd_codes(ops, x, qp, y, codes):
    from quantization.mxnet_b200._ffi import check_call, current_stream, dl
    a, q, o, c = dl(x), dl(qp), dl(y), dl(codes)
    check_call(ops._lib().fq_forward_scalar(a.ptr, q.ptr, o.ptr, c.ptr, current_stream()))


def run_b200(args):
    import torch
    import torch.distributed as dist
    from quantization.mxnet_b200 import dist as fqdist
    from quantization.mxnet_b200 import ops
    from quantization.mxnet_b200.quantize.distribution_calibrate import collect_feature_maps, kl_calibrate_all

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    torch.backends.cudnn.allow_tf32 = False          # the framework convolutions stay plain fp32
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.benchmark = not args.no_e2e      # autotune the framework forward of the e2e leg only
    K, W = args.steps, max(args.warmup, 1)
    t_wall0 = time.time()

    # everything after the headline is best effort under a deadline: if an optional block hangs (a captured NCCL
    # graph, a box hiccup) the line is still printed with what has been measured, and the process leaves
    line = {}
    printed = threading.Event()

    def emit():
        if rank == 0 and not printed.is_set():
            printed.set()
            print(json.dumps(line), flush=True)

    def watchdog():
        emit_at = t_wall0 + args.deadline
        while time.time() < emit_at:
            if printed.is_set():
                return
            time.sleep(0.5)
        line.setdefault("notes", []).append("deadline of %d s reached: optional blocks cut short" % args.deadline)
        emit()
        sys.stdout.flush()
        os._exit(0)
    threading.Thread(target=watchdog, daemon=True).start()

    # one sampler per rank, started before the set-up so that nvidia-smi (whose start-up takes up to seconds on an
    # 8-GPU box) is already producing a sample every 100 ms when the timed region begins; windows are cut by host time
    clocks = ClockSampler(local).start()
    net = build_net(dev)
    net.disable_quantize()           # calibrate with fp32 inputs and weights (simulate_quantization.py:298)
    g = torch.Generator(device="cpu").manual_seed(7 + rank)
    X_host = torch.randn(BATCH, 3, 224, 224, generator=g).pin_memory()
    X = X_host.to(dev)
    acts = capture_layer_inputs(net, X)
    n_elems = sum(a.numel() for a in acts)
    assert len(acts) == N_LAYERS and n_elems == ELEMS_PER_IMAGE * BATCH, (len(acts), n_elems)

    hist = torch.zeros(N_LAYERS, BINS + 1, dtype=torch.float32, device=dev)
    launches = [0]

    def fold(c, first):
        ops.hist_accumulate(c.reshape(-1), hist.view(-1), first)
        launches[0] += 1
    # integer counts of up to RING batches share ONE asynchronous sum-all-reduce; the float32 adds are replayed in
    # batch order.  32-bit on the wire: one rank's count per bin is bounded by its largest layer input.
    ring = fqdist.CountsRing(N_LAYERS, BINS + 1, dev, accumulate=fold, slots=RING,
                             max_count=max(a.numel() for a in acts) if world > 1 else None)
    minmax = torch.zeros(N_LAYERS, 2, dtype=torch.float32, device=dev)
    div = torch.empty(N_LAYERS, BINS, dtype=torch.float64, device=dev)
    margin = torch.empty(N_LAYERS, dtype=torch.float64, device=dev)
    thresholds = torch.empty(N_LAYERS, dtype=torch.float32, device=dev)
    bad = torch.zeros(N_LAYERS, dtype=torch.int32, device=dev) if args.check_inputs else None

    def step(first, ev=None):
        if first:
            for i, a in enumerate(acts):
                ops.minmax(a, out=minmax[i])
            if world > 1:
                mx = minmax[:, 1].contiguous()
                dist.all_reduce(mx, op=dist.ReduceOp.MAX)
                minmax[:, 1].copy_(mx)
        if ev is not None:
            ev[0].record()
        ops.hist_nonzero_multi(acts, minmax, 2, 1, BINS, ring.slot(), promotion="nep50", bad_flags=bad)   # 27 layers, one launch
        if ev is not None:
            ev[1].record()
        launches[0] += 1
        ring.commit()

    def kl_close():
        ring.flush()
        best, _ = ops.kl_search(hist[:, :BINS], LEVELS, LEVELS, BINS, promotion="nep50", divergence=div, margin=margin)
        ops.kl_threshold(best, minmax[:, 1].contiguous(), BINS, out=thresholds)
        launches[0] += 3
        return best

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for w in range(W):
        step(w == 0)
    kl_close()
    # warm-up also covers the collective: NCCL connects an algorithm the first time a message size selects it
    ring.prime([min(RING, K), K % RING])
    barrier()

    t_clk0 = time.time()
    reps = []
    hist_ms_all = 0.0
    launches[0] = 0
    for rep in range(REPEATS):
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
        t_start, t_kl, t_end = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        barrier()
        t_start.record()
        for k in range(K):
            step(False, evs[k])
        t_kl.record()
        best = kl_close()
        t_end.record()
        barrier()
        hist_ms = sum(a.elapsed_time(b) for a, b in evs)
        reps.append((t_start.elapsed_time(t_end), t_kl.elapsed_time(t_end), hist_ms))
        hist_ms_all += hist_ms
    gpu_launches = launches[0]
    # per repetition: max over ranks of the total; the median repetition is the headline
    tot = torch.tensor([r[0] for r in reps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.MAX)
    tot_sorted = sorted(float(t) for t in tot)
    total_ms_max = tot_sorted[len(tot_sorted) // 2]
    value = world * BATCH * K / (total_ms_max * 1e-3)
    mine = torch.tensor([hist_ms_all / (REPEATS * K), sorted(r[1] for r in reps)[REPEATS // 2],
                         sorted(r[0] for r in reps)[REPEATS // 2]], device=dev, dtype=torch.float64)
    per_rank = [mine.clone() for _ in range(world)]
    if world > 1:
        dist.all_gather(per_rank, mine)
    t_clk1 = time.time()
    # a short timed region (40 steps x 5 repeats ~ 80 ms) can fall between two 100 ms samples: keep the GPU under the
    # same load for a number of extra, untimed repetitions that every rank derives from the SAME reduced time (the
    # ring's collectives must stay in step), so that the window holds >= 4 samples
    extra = max(0, int(math.ceil((450.0 - sum(tot_sorted)) / max(total_ms_max, 1e-3))))
    for _ in range(min(extra, 200)):
        for k in range(K):
            step(False)
        kl_close()
        torch.cuda.synchronize()
    t_clk1 = time.time()
    clock_info = clocks.summary(t_clk0, t_clk1)
    clock_info["window_s"] = round(t_clk1 - t_clk0, 3)
    clock_info["window"] = "the timed repetitions + %d untimed repetitions of the same step" % min(extra, 200)
    clock_all = [None] * world
    if world > 1:
        dist.all_gather_object(clock_all, clock_info)
    else:
        clock_all = [clock_info]
    t_clk2 = time.time()                                                 # a second window: parity + e2e legs
    min_margin = float(margin.min())

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_kind = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_kind = 6650.0, "fallback"
    per_launch_bytes = 4.0 * n_elems              # one multi-tensor launch reads every layer input once
    per_launch_s = float(per_rank[0][0]) * 1e-3
    achieved = per_launch_bytes / per_launch_s / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "hist_kernel_traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
    reasons = sorted({r for c in clock_all if c for r in c.get("reasons", [])})
    line.update({
        "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": total_ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": workload_config(world),
        "repeats": {"n": REPEATS, "stat": "median", "total_ms_each_max_over_ranks": [round(float(t), 4) for t in tot]},
        "roofline": {"bound": "hbm", "kernel": "fq::hist_multi_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "traffic_source": "ncu --set full capture under profiles/",
                     "peak_kind": peak_kind, "frac_of_nominal_8000": achieved / 8000.0,
                     "algorithmic_bytes_per_launch": per_launch_bytes, "avg_launch_us": per_launch_s * 1e6,
                     "launches_timed": REPEATS * K, "input_checks_in_kernel": bool(args.check_inputs)},
        "breakdown_ms": {"hist_per_step": float(per_rank[0][0]), "kl_search_once": float(per_rank[0][1]),
                         "total": total_ms_max},
        "per_rank": [{"rank": r, "hist_us_per_launch": round(float(p[0]) * 1e3, 2), "kl_close_ms": round(float(p[1]), 4),
                      "total_ms": round(float(p[2]), 4),
                      "sm_mhz": (clock_all[r] or {}).get("sm_mhz"), "mem_mhz": (clock_all[r] or {}).get("mem_mhz"),
                      "power_w_max": (clock_all[r] or {}).get("power_w_max")} for r, p in enumerate(per_rank)],
        "cpu_baseline": None, "e2e": None,
        "clocks": dict(clock_info, reasons=reasons), "gpu_launches": gpu_launches,
        "kl_best_bins": [int(b) for b in best.cpu()], "kl_min_margin": min_margin,
    })
    del acts
    torch.cuda.empty_cache()

    # ---- parity (all ranks) -----------------------------------------------------------------------
    if not args.no_parity:
        try:
            p = parity_block(world, rank, dev)
            line["parity"] = p
        except Exception as e:
            line["parity"] = {"error": "%s: %s" % (type(e).__name__, str(e)[:300])}
        torch.cuda.empty_cache()
        barrier()

    # ---- e2e through the public API with host buffers -------------------------------------------
    if not args.no_e2e:
        class Loader:
            def __init__(self, n):
                self.n = n

            def __len__(self):
                return self.n

            def __iter__(self):
                for _ in range(self.n):
                    yield X_host, None

        def calibrate(nb):
            hc, mc = collect_feature_maps(net, BINS, Loader(nb), ctx=dev, tqdm_desc="e2e")
            b, th = kl_calibrate_all(hc, LEVELS, LEVELS, BINS, fm_max=mc)
            return th.cpu()
        os.environ.setdefault("TQDM_DISABLE", "1")
        calibrate(2)
        barrier()
        t0 = time.perf_counter()
        th = calibrate(K)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        line["e2e"] = {"value": world * BATCH * K / float(tt), "unit": "images/s",
                       "h2d_bytes_per_step": X_host.numel() * 4,
                       "d2h_bytes_per_step": (N_LAYERS * (BINS + 1) * 4 + N_LAYERS * 12 + N_LAYERS * 4) / K,
                       "ms_per_step": 1e3 * float(tt) / K,
                       "api": "quantize.distribution_calibrate.collect_feature_maps + kl_calibrate_all on the torch "
                              "mobilenet1.0 (fp32 cuDNN forward included), pinned host images"}
        barrier()
    if rank == 0:
        c2 = clocks.summary(t_clk2, time.time())
        line["clocks"]["e2e_window"] = c2
        line["clocks"]["reasons"] = sorted(set(line["clocks"]["reasons"]) | set(c2.get("reasons", [])))
    del net, X
    torch.cuda.empty_cache()

    # ---- config 5 in brief (rank 0; the other ranks wait at the barrier) ---------------------------
    if not args.no_sweep:
        if rank == 0:
            try:
                line["sweep"] = sweep_block(dev, peak)
            except Exception as e:
                line["sweep"] = {"error": "%s: %s" % (type(e).__name__, str(e)[:300])}
        torch.cuda.empty_cache()
        barrier()

    # ---- QConv2D on the tensor cores (rank 0, a few hundred ms) ---------------------
    if not args.no_qconv:
        if rank == 0:
            try:
                line["qconv"] = qconv_block(dev)
            except Exception as e:
                line["qconv"] = {"error": "%s: %s" % (type(e).__name__, str(e)[:300])}
        torch.cuda.empty_cache()
        barrier()

    # ---- configs 1 / 3 / 4 at this N --------------------------------------------------------------
    graphs_live = False
    if not args.no_configs:
        import bench_configs as BC
        cfgs = {}
        for cid, kw in ((1, dict(steps=30, warmup=5, graph=True)), (3, dict(steps=20, warmup=5, graph=True)),
                        (4, dict(steps=6, warmup=2, graph=False))):
            try:
                cfgs["config%d" % cid] = BC.run_config(cid, dev, world, rank, **kw)
            except Exception as e:
                cfgs["config%d" % cid] = {"error": "%s: %s" % (type(e).__name__, str(e)[:300])}
            torch.cuda.empty_cache()
        line["configs"] = cfgs
        graphs_live = BC.graphs_captured_collectives() and world > 1

    # ---- CPU baseline on a bounded sample (rank 0, N=1 only) ---------------------------------------
    if rank == 0 and world == 1 and not args.no_cpu:
        procs = os.cpu_count() or 1
        line["cpu_baseline"] = cpu_arm(K, 2, 0, procs, 4)

    line["wall_s"] = round(time.time() - t_wall0, 1)
    clocks.stop()
    emit()
    if world > 1:
        if graphs_live:
            # a live CUDA graph that captured NCCL work keeps the communicator busy and destroy_process_group()
            # never returns: drop the graphs, line the ranks up and leave without the teardown
            BC.release_graphs()
            torch.cuda.synchronize()
            dist.barrier()
            sys.stdout.flush()
            os._exit(0)
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-sweep", action="store_true")
    ap.add_argument("--no-configs", action="store_true")
    ap.add_argument("--no-qconv", action="store_true")
    ap.add_argument("--check-inputs", type=int, default=1,
                    help="1: the histogram kernel also evaluates the reference's per-batch asserts (>= 0, no NaN)")
    ap.add_argument("--deadline", type=int, default=420, help="seconds after which the line is printed as is")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
