# run AG (1 GPU): whole GPU suite, then the driver-form bench line
set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -q 2>&1 | grep -v "^E   *+\|^E   *and" | tail -8 | cut -c1-250 > gpurun_out/r2_tests_ag.log
tail -4 gpurun_out/r2_tests_ag.log
timeout 600 python bench.py > gpurun_out/r2_bench_ag_n1.json 2> gpurun_out/r2_bench_ag_n1.err
tail -c 400 gpurun_out/r2_bench_ag_n1.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_bench_ag_n1.json").read().strip().splitlines()[-1])
print("N=1", d["value"], d["ms_per_step"], d["roofline"]["frac"], d.get("wall_s"), d.get("notes"))
print(d.get("clocks")); print(d.get("e2e")); print({k:v for k,v in d.get("cpu_baseline",{}).items() if k!="sample"})
for r in d.get("sweep",{}).get("rows",[]): print(r.get("kernel"), r.get("log2n"), round(r.get("gbs",0)), r.get("frac_of_8000"))
for k,v in d.get("configs",{}).items(): print(k, {a:b for a,b in v.items() if a not in ("workload","converters","conv","dtype","data","scaling","metric","name")})
PY
