set -x
cd $GRAFT_REPO_ROOT
for m in 0 1; do
FQ_FORWARD_BULK=$m timeout 300 python bench_sweep.py --min-log2 24 --max-log2 30 --step 2 --reps 15 --kernels fwd_scalar_u8 --out gpurun_out/r2_sweep_fwd_bulk$m.json 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l); print('bulk=$m', d['kernel'], d['log2n'], round(d['gbs_median']), round(d['median_us'], 1))
    except Exception: pass"
done
FQ_FORWARD_BULK=1 timeout 200 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "forward_scalar" 2>&1 | tail -2
timeout 300 python bench_sweep.py --min-log2 32 --max-log2 32 --reps 8 --kernels fwd_offline_track_n128,fwd_scalar_u8 --out gpurun_out/r2_sweep_2p32_final.json 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l); print(d['kernel'], d['log2n'], round(d['gbs_median']), round(d['median_us'], 1))
    except Exception: pass"
