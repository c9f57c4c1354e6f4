"""Where does a data-parallel QAT step (config 3, notebook converters, CUDA graph) spend its time?

    python tools/qat_dp_probe.py                                   1 GPU
    torchrun --nproc-per-node N tools/qat_dp_probe.py              N GPUs

Replays the captured step under torch.profiler (kineto) and prints, for rank 0: the step time, GPU-busy time of the
default-stream kernels, the NCCL kernels' durations and the largest gaps -- the numbers behind DESIGN "QAT scaling"."""
import json
import os
import sys

import torch
import torch.distributed as dist
from torch import nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench_configs as BC  # noqa: E402
from quantization.mxnet_b200 import dist as fqdist  # noqa: E402


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.benchmark = True
    cfg = BC.CONFIGS[3]
    net = BC.build(cfg, dev)
    fqdist.broadcast_parameters(net)
    g = torch.Generator(device=dev).manual_seed(7 + rank)
    X = torch.randn(*cfg["shape"], device=dev, generator=g)
    y = torch.randint(0, cfg["classes"], (cfg["shape"][0],), device=dev, generator=g)
    net.train()
    for m in net.modules():
        if isinstance(m, nn.BatchNorm2d):
            m.eval()
    net.quantize_input(enable=True, online=True)
    with torch.no_grad():
        net(X)
    net.update_ema()
    net.quantize_input(enable=True, online=False)
    if world > 1:
        fqdist.enable_data_parallel(net)
    params = [p for p in net.parameters() if p.requires_grad]
    opt = torch.optim.Adam(params, lr=1e-6, capturable=True, fused=True)
    align = int(os.environ.get("FQ_BUCKET_ALIGN", "32"))
    bucket = fqdist.GradBucket(params, net=net if world > 1 else None, align=align)
    loss_fn = nn.CrossEntropyLoss()

    def step():
        opt.zero_grad(set_to_none=False)
        loss = loss_fn(net(X), y)
        net.update_ema()
        loss.backward()
        bucket.all_reduce_mean()
        opt.step()
        return loss
    graph, _ = BC.capture(step)
    t = BC.timed(graph.replay, 30, 5, world)
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(3):
            graph.replay()
        torch.cuda.synchronize()
    if rank == 0:
        evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
        evs.sort(key=lambda e: e.time_range.start)
        t0, t1 = evs[0].time_range.start, max(e.time_range.end for e in evs)
        busy = sum(e.time_range.end - e.time_range.start for e in evs)
        nccl = [e for e in evs if "nccl" in e.name.lower()]
        fq = [e for e in evs if e.name.startswith("fq::") or "fq::" in e.name]
        by = {}
        for e in evs:
            k = e.name[:60]
            d = by.setdefault(k, [0, 0.0])
            d[0] += 1
            d[1] += e.time_range.end - e.time_range.start
        top = sorted(by.items(), key=lambda kv: -kv[1][1])[:14]
        full = {}
        for e in evs:
            d = full.setdefault(e.name[:300], [0, 0.0])
            d[0] += 1
            d[1] += e.time_range.end - e.time_range.start
        out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
        os.makedirs(out_dir, exist_ok=True)
        with open(os.path.join(out_dir, "r2_qat_kernels_n%d_align%d.json" % (world, align)), "w") as f:
            json.dump({k: [v[0] / 3, round(v[1] / 3, 2)] for k, v in sorted(full.items(), key=lambda kv: -kv[1][0])}, f, indent=0)
        gaps = sorted(((evs[i + 1].time_range.start - evs[i].time_range.end, evs[i].name[:40], evs[i + 1].name[:40])
                       for i in range(len(evs) - 1)), reverse=True)[:8]
        print(json.dumps({"n_gpus": world, "bucket_align": align, "graph_ms_per_step": t, "profiled_replays": 3,
                          "span_us_per_replay": (t1 - t0) / 3, "kernel_busy_us_per_replay": busy / 3,
                          "kernels_per_replay": len(evs) / 3,
                          "nccl_us_per_replay": sum(e.time_range.end - e.time_range.start for e in nccl) / 3,
                          "nccl_kernels": [(e.name[:50], round(e.time_range.end - e.time_range.start, 1)) for e in nccl[:6]],
                          "fq_us_per_replay": sum(e.time_range.end - e.time_range.start for e in fq) / 3,
                          "fq_kernels_per_replay": len(fq) / 3,
                          "top_kernels_us_per_replay": [(k, v[0] / 3, round(v[1] / 3, 1)) for k, v in top],
                          "largest_gaps_us": [(round(g_, 1), a, b) for g_, a, b in gaps]}), flush=True)
    if world > 1:
        BC.release_graphs()
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
