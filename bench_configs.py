#!/usr/bin/env python
"""The other BASELINE.json configurations, end to end through the drop-in Python API (torch networks,
cuDNN fp32 convolutions as the framework calls, libfq_b200 kernels for everything on the hot path).

    python bench_configs.py --config 1 [--graph]     cifar_resnet20_v1 online uint8 / int8 inference forward, N=128
    python bench_configs.py --config 3 [--graph]     mobilenetv2_1.0 CIFAR QAT step with the notebook's converters
    python bench_configs.py --config 4               resnet50_v1 4-bit per-group + fake-BN, EMA calibration step, N=256/GPU
    torchrun --nproc-per-node N bench_configs.py --config 3 --gpus N      data parallel (NCCL)

``run_config()`` is also what ``bench.py`` calls for its ``configs`` block, so the driver-run JSON line carries
these numbers at every N.  Each run reports images/s of the quantised step, of the same step with quantisation
disabled (framework only; the fake-BN fold and its statistics hook stay active, as in the reference), and their
ratio -- what the fake-quant path costs on top of the network.  CUDA events on torch's stream, W warm-up + K timed
steps, barrier + synchronize on both sides, max over ranks.  Inputs are synthetic N(0,1) images resident on the GPU
(a 128x3x32x32 batch is 1.5 MB: the kernels, not PCIe, are under test here; bench.py carries the host-buffer e2e).
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist
from torch import nn

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from quantization.mxnet_b200 import dist as fqdist  # noqa: E402
from quantization.mxnet_b200 import model_zoo as Z  # noqa: E402
from quantization.mxnet_b200.quantize import convert  # noqa: E402
from quantization.mxnet_b200.quantize.initialize import qparams_init  # noqa: E402

# config 3 uses the converters of examples/quantize_aware_training_cifar10.ipynb (cell 6, :109-112):
#   gen_conv2d_converter(quant_type="channel", fake_bn=True, input_width=4, weight_width=4),
#   gen_dense_converter(quant_type="channel", input_width=4, weight_width=4), BatchNorm -> bypass_bn
NOTEBOOK = dict(quant_type="channel", fake_bn=True, input_width=4, weight_width=4)
CONFIGS = {
    1: dict(model="cifar_resnet20_v1", classes=10, shape=(128, 3, 32, 32), conv={}, kind="infer_online",
            name="cifar_resnet20_v1 simulate_quantization, per-layer int8 weights / uint8 online inputs, batch 128 of 32x32"),
    3: dict(model="mobilenetv2_1.0", classes=10, shape=(128, 3, 32, 32), conv=NOTEBOOK, kind="qat",
            name="mobilenetv2_1.0 CIFAR-10 QAT step with the notebook's converters (per-channel 4-bit weights, 4-bit "
                 "inputs, fake-BN + bypass_bn): fake-quant fwd + identity-STE bwd + EMA of input_max and of the "
                 "fake-BN running statistics + Adam lr 1e-6, offline inputs with range tracking (the phase after "
                 "the notebook's offline switch), batch 128 per GPU"),
    4: dict(model="resnet50_v1", classes=1000, shape=(256, 3, 224, 224),
            conv=dict(weight_width=4, quant_type="group", fake_bn=True), kind="ema_calib",
            name="resnet50_v1 ImageNet per-group 4-bit weights with merge-BN (fake-BN), EMA calibration step "
                 "(online uint8 inputs, update_ema of input_max and the fake-BN statistics), batch 256 per GPU"),
}


def build(cfg, dev):
    torch.manual_seed(7)
    net = Z.get_model(cfg["model"], classes=cfg["classes"]).to(dev)
    ck = cfg["conv"]
    dk = {k: v for k, v in ck.items() if k in ("weight_width", "quant_type", "input_width")}
    fn = {nn.Conv2d: convert.gen_conv2d_converter(**ck), nn.Linear: convert.gen_dense_converter(**dk), nn.ReLU: None,
          nn.BatchNorm2d: convert.bypass_bn if ck.get("fake_bn") else None}
    convert.convert_model(net, exclude=Z.default_exclusions(net, cfg["model"]), convert_fn=fn)
    qparams_init(net)
    return net


def timed(step, K, W, world):
    for _ in range(W):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(K):
        step()
    b.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b)], device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t) / K


_GRAPHS = []


def capture(step, warm=3):
    """Record `step` as a CUDA graph after `warm` eager iterations on a side stream -> (graph, static result)."""
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(warm):
            step()
    torch.cuda.current_stream().wait_stream(s)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=s):
        out = step()
    _GRAPHS.append(graph)
    return graph, out


def graphs_captured_collectives():
    return bool(_GRAPHS)


def release_graphs():
    for g_ in _GRAPHS:
        g_.reset()
    del _GRAPHS[:]


def run_config(config, dev, world=1, rank=0, steps=20, warmup=5, graph=False):
    """One configuration on this rank's GPU (all ranks call it together).  Returns the result dict (every rank)."""
    cfg = CONFIGS[config]
    saved = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.benchmark = True
    try:
        return _run_config(config, cfg, dev, world, rank, steps, warmup, graph)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark = saved


def _run_config(config, cfg, dev, world, rank, steps, warmup, graph):
    net = build(cfg, dev)
    fqdist.broadcast_parameters(net)
    g = torch.Generator(device=dev).manual_seed(7 + rank)
    X = torch.randn(*cfg["shape"], device=dev, generator=g)
    y = torch.randint(0, cfg["classes"], (cfg["shape"][0],), device=dev, generator=g)
    batch = cfg["shape"][0]
    extra = {}

    if cfg["kind"] == "infer_online":
        net.eval()
        net.fix_params()
        net.quantize_input(enable=True, online=True)           # simulate_quantization.py:346-347
        if world > 1:
            fqdist.enable_data_parallel(net)

        def step():
            with torch.no_grad():
                return net(X)
        step()                                                  # caches the quantised weights (fixed_params -> 1)
        t_q = timed(step, steps, warmup, world)
        if graph:
            # with ranks the per-layer NCCL all-gathers of the per-sample maxima are captured with the rest
            try:
                gr, out = capture(step)
                ref = step()
                gr.replay()
                torch.cuda.synchronize()
                extra["graph_output_equals_eager"] = bool(torch.equal(out, ref))
                extra["graph_ms_per_step"] = timed(gr.replay, steps, warmup, world)
                extra["graph_images_per_sec"] = world * batch / (extra["graph_ms_per_step"] * 1e-3)
            except Exception as e:
                if world == 1:
                    raise
                extra["graph_error"] = str(e)[:300]
        net.disable_quantize()
        if "graph_ms_per_step" in extra:
            graph_f, _ = capture(step)
            extra["graph_framework_only_ms_per_step"] = timed(graph_f.replay, steps, warmup, world)
        t_f = timed(step, steps, warmup, world)
    elif cfg["kind"] == "ema_calib":
        net.eval()
        net.quantize_input(enable=True, online=True)           # simulate_quantization.py:322
        if world > 1:
            fqdist.enable_data_parallel(net)

        def step():
            with torch.no_grad():
                out = net(X)
            net.update_ema()                                    # evaluate(..., update_ema=True), :133
            return out
        t_q = timed(step, steps, warmup, world)
        net.disable_quantize()

        def step_f():
            with torch.no_grad():
                return net(X)
        t_f = timed(step_f, steps, warmup, world)
    else:       # QAT step of the notebook (cell 15), after the switch to offline inputs
        net.train()
        for m in net.modules():
            if isinstance(m, nn.BatchNorm2d):
                m.eval()
        net.quantize_input(enable=True, online=True)
        with torch.no_grad():
            net(X)
        net.update_ema()                                        # a non-zero input_max to quantise against
        net.quantize_input(enable=True, online=False)
        if world > 1:
            fqdist.enable_data_parallel(net)
        params = [p for p in net.parameters() if p.requires_grad]
        opt = torch.optim.Adam(params, lr=1e-6, capturable=graph, fused=True)       # one multi-tensor kernel, like MXNet's adam_update
        bucket = fqdist.GradBucket(params, net=net if world > 1 else None)   # input ranges ride with the gradients
        loss_fn = nn.CrossEntropyLoss()

        def step():
            opt.zero_grad(set_to_none=False)
            loss = loss_fn(net(X), y)
            net.update_ema()                                    # before backward, as in the notebook
            loss.backward()
            bucket.all_reduce_mean()
            opt.step()
            return loss
        t_q = timed(step, steps, warmup, world)
        if graph:
            # the whole QAT step (forward, EMA, backward, gradient all-reduce, Adam) as one CUDA graph: the 52
            # layers' launches replay back to back with no Python in between.  With ranks, the NCCL collectives
            # (gradient bucket with the per-sample maxima in its tail, fake-BN statistics records) are captured too.
            try:
                gr, loss_g = capture(step)
                gr.replay()
                torch.cuda.synchronize()
                extra["graph_loss_finite"] = bool(torch.isfinite(loss_g).item())
                extra["graph_ms_per_step"] = timed(gr.replay, steps, warmup, world)
                extra["graph_images_per_sec"] = world * batch / (extra["graph_ms_per_step"] * 1e-3)
            except Exception as e:          # capture of the collectives is the only part that can refuse
                if world == 1:
                    raise
                extra["graph_error"] = str(e)[:300]
        net.disable_quantize()
        t_f = timed(step, steps, warmup, world)
        if "graph_ms_per_step" in extra:
            graph_f, _ = capture(step)
            extra["graph_framework_only_ms_per_step"] = timed(graph_f.replay, steps, warmup, world)

    line = {"config": config, "workload": cfg["name"], "n_gpus": world, "steps": steps, "warmup": warmup,
            "metric": "images_per_sec", "value": world * batch / (t_q * 1e-3), "ms_per_step": t_q,
            "framework_only_images_per_sec": world * batch / (t_f * 1e-3), "framework_only_ms_per_step": t_f,
            "quantisation_overhead": t_q / t_f - 1.0, "scaling": "weak", "dtype": "f32", "data": "synthetic",
            "conv": "cuDNN fp32 (TF32 off)", "converters": cfg["conv"] or "defaults (8-bit per-layer)"}
    line.update(extra)
    if "graph_ms_per_step" in extra and "graph_framework_only_ms_per_step" in extra:
        line["graph_quantisation_overhead"] = extra["graph_ms_per_step"] / extra["graph_framework_only_ms_per_step"] - 1.0
    del net
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, required=True, choices=sorted(CONFIGS))
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--graph", action="store_true", help="also replay the step as a CUDA graph")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    line = run_config(args.config, dev, world, rank, args.steps, args.warmup, args.graph)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        if graphs_captured_collectives():
            # a live CUDA graph that captured NCCL work keeps the communicator busy: destroy_process_group() never
            # returns (seen on 2 GPUs).  Drop the graphs, line the ranks up and leave without the teardown.
            release_graphs()
            torch.cuda.synchronize()
            dist.barrier()
            sys.stdout.flush()
            os._exit(0)
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
