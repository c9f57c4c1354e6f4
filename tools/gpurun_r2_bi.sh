# run BI (1 GPU): final tree -- whole GPU suite, driver-form bench, full QConv probe
set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/r2_tests_bi.log
cat gpurun_out/r2_tests_bi.log
timeout 600 python bench.py > gpurun_out/r2_bench_bi_n1.json 2> gpurun_out/r2_bench_bi_n1.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_bench_bi_n1.json").read().strip().splitlines()[-1])
print("N=1", d["value"], d["ms_per_step"], d["roofline"]["frac"], d.get("wall_s"), d.get("notes"), d["clocks"]["reasons"], d["e2e"]["value"], d["cpu_baseline"]["value"])
print([(r["igemm_us"], r["layer_us"], r["roofline"]["frac"]) for r in d["qconv"]["rows"]])
for r in d["sweep"]["rows"]:
    if r["log2n"] == 30: print(r["kernel"], round(r["gbs"]), r["frac_of_8000"])
PY
timeout 300 python tools/qconv_probe.py 2>&1 | tee gpurun_out/r2_qconv_probe_bi.txt
