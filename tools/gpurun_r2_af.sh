set -x
cd $GRAFT_REPO_ROOT
for m in 0 1; do
FQ_TRACK_MODE=$m timeout 300 python bench_sweep.py --min-log2 24 --max-log2 30 --step 2 --reps 15 --kernels fwd_offline_track_n128 --out gpurun_out/r2_sweep_track_mode$m.json 2>&1 | tail -4 | cut -c1-200
done
FQ_TRACK_MODE=1 timeout 200 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_api.py -m gpu -q -k "online or offline or track or config3 or qat or randomised" 2>&1 | tail -2
