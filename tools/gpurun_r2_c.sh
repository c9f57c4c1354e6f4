# run C (1 GPU): kernels touched since run B -- online-path modes, offline-track variants, channel stats
set -x
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2_tests_c.log
cat gpurun_out/r2_tests_c.log
rm -f gpurun_out/r2_graph_probe_c.txt gpurun_out/r2_config1_modes_c.jsonl
for f in 2 1 0; do
  echo "FQ_ONLINE_MODE=$f" >> gpurun_out/r2_graph_probe_c.txt
  FQ_ONLINE_MODE=$f timeout 300 python tools/graph_probe.py >> gpurun_out/r2_graph_probe_c.txt 2>&1
  FQ_ONLINE_MODE=$f timeout 300 python bench_configs.py --config 1 --graph --steps 50 >> gpurun_out/r2_config1_modes_c.jsonl 2>> gpurun_out/r2_c.err
done
cat gpurun_out/r2_graph_probe_c.txt
for v in 0 1 2; do
  FQ_TRACK_VARIANT=$v timeout 300 python bench_sweep.py --min-log2 26 --max-log2 30 --step 2 --reps 15 --kernels fwd_offline_track_n128 --out gpurun_out/r2_sweep_c_track$v.json 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l); print('track variant $v', d['log2n'], round(d['gbs_median']), round(d['median_us'], 1))
    except Exception: pass"
done
timeout 300 python bench.py --steps 40 --warmup 3 --no-e2e --no-configs --no-cpu --no-parity > gpurun_out/r2_bench_c.json 2>> gpurun_out/r2_c.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_bench_c.json").read().strip().splitlines()[-1])
print(d["value"], d["roofline"]["frac"], d["clocks"])
for r in d["sweep"]["rows"]:
    if "channel" in r["kernel"] or "offline" in r["kernel"] or "online" in r["kernel"]: print(r)
PY
tail -5 gpurun_out/r2_c.err
