set -x
cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem --format=csv
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2_tests_a.log
cat gpurun_out/r2_tests_a.log
timeout 600 python bench.py --steps 40 --warmup 3 > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err
tail -c 3000 gpurun_out/r2_bench_a.err
timeout 300 python bench.py --steps 40 --warmup 3 --check-inputs 0 --no-e2e --no-sweep --no-configs --no-cpu --no-parity > gpurun_out/r2_bench_a_nocheck.json 2>> gpurun_out/r2_bench_a.err
python - <<'PY'
import json
for f in ("gpurun_out/r2_bench_a.json","gpurun_out/r2_bench_a_nocheck.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["avg_launch_us"], d.get("parity"), d.get("wall_s"))
    except Exception as e:
        print(f, "ERR", e)
PY
