# run H (8 GPUs): the headline bench with every block at N=8, and the QAT data-parallel probe
set -x
cd $GRAFT_REPO_ROOT
nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 40 --warmup 3 > gpurun_out/r2_bench_h_n8.json 2> gpurun_out/r2_bench_h_n8.err
tail -c 800 gpurun_out/r2_bench_h_n8.err
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r2_bench_h_n8.json").read().strip().splitlines()[-1])
    print("N=8", d["value"], d["ms_per_step"], d["roofline"]["frac"], d.get("wall_s"), d.get("notes"))
    print(d.get("parity")); print(d.get("e2e"))
    for r in d.get("per_rank", []): print(r)
    for k,v in d.get("configs",{}).items(): print(k, {a:b for a,b in v.items() if a not in ("workload","converters","conv","dtype","data","scaling","metric")})
except Exception as e:
    print("ERR", e)
PY
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 tools/qat_dp_probe.py 2>> gpurun_out/r2_h.err | cut -c1-700
