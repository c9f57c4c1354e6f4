# run B (2 GPUs): multi-GPU tests, kernels touched since run A, the online-path A/B, bench at N=2
set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_dist.py -x -q 2>&1 | tail -15 > gpurun_out/r2_tests_b_dist.log
cat gpurun_out/r2_tests_b_dist.log
timeout 1500 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_dist.py 2>&1 | tail -15 > gpurun_out/r2_tests_b.log
cat gpurun_out/r2_tests_b.log
for f in 2 1 0; do
  echo "FQ_ONLINE_MODE=$f" >> gpurun_out/r2_graph_probe_b.txt
  FQ_ONLINE_MODE=$f CUDA_VISIBLE_DEVICES=0 timeout 300 python tools/graph_probe.py >> gpurun_out/r2_graph_probe_b.txt 2>&1
  FQ_ONLINE_MODE=$f CUDA_VISIBLE_DEVICES=0 timeout 300 python bench_configs.py --config 1 --graph --steps 50 >> gpurun_out/r2_config1_fused_ab.jsonl 2>> gpurun_out/r2_b.err
done
cat gpurun_out/r2_graph_probe_b.txt
CUDA_VISIBLE_DEVICES=0 timeout 300 python bench_sweep.py --min-log2 26 --max-log2 30 --step 2 --reps 15 --kernels fwd_offline_track_n128,fwd_online_n128,fwd_scalar_u8,hist2048 --out gpurun_out/r2_sweep_b.json > gpurun_out/r2_sweep_b.log 2>&1
tail -20 gpurun_out/r2_sweep_b.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 40 --warmup 3 > gpurun_out/r2_bench_b_n2.json 2> gpurun_out/r2_bench_b_n2.err
tail -c 1500 gpurun_out/r2_bench_b_n2.err
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r2_bench_b_n2.json").read().strip().splitlines()[-1])
    print("N=2", d["value"], d["ms_per_step"], d["roofline"]["frac"], d.get("parity"), d.get("wall_s"), d.get("per_rank"))
    for k,v in d.get("configs",{}).items(): print(k, {a:b for a,b in v.items() if a!="workload"})
    print(d.get("e2e"))
except Exception as e:
    print("ERR", e)
PY
