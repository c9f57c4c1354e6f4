set -x
cd $GRAFT_REPO_ROOT
for m in 0 1; do
FQ_TRACK_MODE=$m timeout 300 python bench_sweep.py --min-log2 26 --max-log2 30 --step 2 --reps 15 --kernels fwd_offline_track_n128,fwd_scalar_u8 --out gpurun_out/r2_sweep_track_mode$m.json 2>&1 | tail -8
done
FQ_TRACK_MODE=1 timeout 200 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "online or offline or track" 2>&1 | tail -2
for m in 0 1; do
FQ_TRACK_MODE=$m timeout 300 ncu --set full --clock-control none -k "regex:offline_track" -s 2 -c 1 -f -o gpurun_out/r2_prof_track_mode$m python bench_sweep.py --min-log2 28 --max-log2 28 --reps 3 --kernels fwd_offline_track_n128 > /dev/null 2>&1
done
ls gpurun_out/r2_prof_track_mode*
