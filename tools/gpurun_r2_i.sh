# run I (1 GPU): the whole GPU suite on the current tree, then the driver-form bench line
set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -30 > gpurun_out/r2_tests_i.log
tail -5 gpurun_out/r2_tests_i.log
timeout 600 python bench.py > gpurun_out/r2_bench_i_n1.json 2> gpurun_out/r2_bench_i_n1.err
tail -c 600 gpurun_out/r2_bench_i_n1.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_bench_i_n1.json").read().strip().splitlines()[-1])
print("N=1", d["value"], d["ms_per_step"], d["roofline"]["frac"], d.get("wall_s"), d.get("notes"))
print(d.get("clocks")); print(d.get("e2e")); print(d.get("cpu_baseline"))
PY
