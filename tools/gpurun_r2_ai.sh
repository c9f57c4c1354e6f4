# run AI (8 GPUs): the driver-form bench line at N=8 and N=4 on the final tree
set -x
cd $GRAFT_REPO_ROOT
nvidia-smi -L | wc -l
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 40 --warmup 3 > gpurun_out/r2_bench_ai_n8.json 2> gpurun_out/r2_bench_ai_n8.err
tail -c 300 gpurun_out/r2_bench_ai_n8.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 4 --steps 40 --warmup 3 > gpurun_out/r2_bench_ai_n4.json 2> gpurun_out/r2_bench_ai_n4.err
python - <<'PY'
import json
for n in (8, 4):
    try:
        d=json.loads(open("gpurun_out/r2_bench_ai_n%d.json" % n).read().strip().splitlines()[-1])
    except Exception as e:
        print("ERR", n, e); continue
    print("N=%d" % n, d["value"], d["ms_per_step"], d["roofline"]["frac"], d.get("wall_s"), d.get("notes"))
    print(d.get("clocks")); print(d.get("parity")); print(d.get("e2e"))
    for r in d.get("per_rank", []): print(r)
    for k,v in d.get("configs",{}).items(): print(k, {a:b for a,b in v.items() if a not in ("workload","converters","conv","dtype","data","scaling","metric","name")})
PY
