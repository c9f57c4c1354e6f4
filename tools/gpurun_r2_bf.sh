# run BF (2 GPUs): driver-form bench at N=1 and N=2 with the qconv block
set -x
cd $GRAFT_REPO_ROOT
CUDA_VISIBLE_DEVICES=0 timeout 600 python bench.py > gpurun_out/r2_bench_bf_n1.json 2> gpurun_out/r2_bench_bf_n1.err
timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 40 --warmup 3 > gpurun_out/r2_bench_bf_n2.json 2> gpurun_out/r2_bench_bf_n2.err
python - <<'PY'
import json
for n in (1, 2):
    try:
        d=json.loads(open("gpurun_out/r2_bench_bf_n%d.json" % n).read().strip().splitlines()[-1])
    except Exception as e:
        print("ERR", n, e); continue
    print("N=%d" % n, d["value"], d["ms_per_step"], d["roofline"]["frac"], d.get("wall_s"), d.get("notes"))
    print(d.get("clocks", {}).get("reasons"), d.get("parity", {}).get("ranks_identical"), d.get("e2e", {}).get("value"))
    print([ (r["igemm_us"], r["layer_us"], r["cudnn_fp32_conv_us"], r["roofline"]["frac"]) for r in d.get("qconv", {}).get("rows", [])], d.get("qconv", {}).get("error"))
    print({k: (v.get("value"), v.get("graph_images_per_sec"), v.get("error")) for k, v in d.get("configs", {}).items()})
PY
