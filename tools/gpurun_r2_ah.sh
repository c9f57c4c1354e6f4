# run AH (2 GPUs): NCCL tests on the current tree, then the driver-form bench line at N=2
set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_dist.py -q 2>&1 | tail -6 > gpurun_out/r2_tests_ah_dist.log
cat gpurun_out/r2_tests_ah_dist.log
timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 40 --warmup 3 > gpurun_out/r2_bench_ah_n2.json 2> gpurun_out/r2_bench_ah_n2.err
tail -c 300 gpurun_out/r2_bench_ah_n2.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_bench_ah_n2.json").read().strip().splitlines()[-1])
print("N=2", d["value"], d["ms_per_step"], d["roofline"]["frac"], d.get("wall_s"), d.get("notes"))
print(d.get("clocks")); print(d.get("parity")); print(d.get("e2e"))
for r in d.get("per_rank", []): print(r)
for k,v in d.get("configs",{}).items(): print(k, {a:b for a,b in v.items() if a not in ("workload","converters","conv","dtype","data","scaling","metric","name")})
PY
