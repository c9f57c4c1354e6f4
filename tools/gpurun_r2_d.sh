# run D (2 GPUs): NCCL tests, full GPU suite, config 3 at N=1 (fused fold backward), bench at N=2
set -x
cd $GRAFT_REPO_ROOT
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_dist.py -x -q 2>&1 | tail -25 > gpurun_out/r2_tests_d_dist.log
cat gpurun_out/r2_tests_d_dist.log
timeout 1500 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_dist.py 2>&1 | tail -25 > gpurun_out/r2_tests_d.log
cat gpurun_out/r2_tests_d.log
CUDA_VISIBLE_DEVICES=0 timeout 300 python bench_configs.py --config 3 --graph --steps 30 > gpurun_out/r2_config3_d_n1.json 2> gpurun_out/r2_d.err
cat gpurun_out/r2_config3_d_n1.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 40 --warmup 3 > gpurun_out/r2_bench_d_n2.json 2> gpurun_out/r2_bench_d_n2.err
tail -c 1500 gpurun_out/r2_bench_d_n2.err
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r2_bench_d_n2.json").read().strip().splitlines()[-1])
    print("N=2", d["value"], d["ms_per_step"], d["roofline"]["frac"], d.get("wall_s"))
    print(d.get("parity")); print(d.get("per_rank")); print(d.get("e2e")); print(d.get("notes"))
    for k,v in d.get("configs",{}).items(): print(k, {a:b for a,b in v.items() if a!="workload"})
except Exception as e:
    print("ERR", e)
PY
