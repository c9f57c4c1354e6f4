# run G (1 GPU): full GPU suite, 2^32 sweep of the row kernels, ncu launch list + full captures, final N=1 bench
set -x
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2_tests_g.log
cat gpurun_out/r2_tests_g.log
timeout 300 python tools/ncu_kernels_probe.py 2>&1 | tail -3 | tee gpurun_out/r2_qconv_timing.txt
timeout 400 python bench_sweep.py --min-log2 32 --max-log2 32 --reps 10 --kernels fwd_scalar_u8,fwd_rows64,fwd_rows1024,fwd_offline_track_n128,fwd_online_n128 --out gpurun_out/r2_sweep_g_2p32.json 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l); print(d['kernel'], d['log2n'], round(d['gbs_median']), round(d['median_us'], 1))
    except Exception: pass"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 4 --warmup 3 --no-e2e --no-sweep --no-configs --no-cpu --no-parity > gpurun_out/r2_bench_under_ncu.json 2> gpurun_out/r2_g.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hist_multi_kernel -s 3 -c 1 -f -o gpurun_out/r2_prof_hist python bench.py --steps 2 --warmup 3 --no-e2e --no-sweep --no-configs --no-cpu --no-parity > /dev/null 2>> gpurun_out/r2_g.err
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:offline_track|channel_stats_kernel|online_quant_small|input_path_kernel|fold_backward|qconv_igemm|qconv_pack_input" -s 8 -c 8 -f -o gpurun_out/r2_prof_others python tools/ncu_kernels_probe.py > /dev/null 2>> gpurun_out/r2_g.err
ls -la gpurun_out/*.ncu-rep
timeout 600 python bench.py --steps 40 --warmup 3 > gpurun_out/r2_bench_g_n1.json 2>> gpurun_out/r2_g.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_bench_g_n1.json").read().strip().splitlines()[-1])
print("N=1", d["value"], d["ms_per_step"], d["roofline"]["frac"], d.get("wall_s"), d["e2e"]["value"], d["cpu_baseline"]["value"], d["clocks"])
for k,v in d.get("configs",{}).items(): print(k, {a:b for a,b in v.items() if a not in ("workload","converters","conv","dtype","data","scaling","metric")})
PY
tail -5 gpurun_out/r2_g.err
