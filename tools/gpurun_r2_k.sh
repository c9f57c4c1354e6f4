# run K (1 GPU): GPU suite, eager host profile with the call plans, framework forward layout probe
set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r2_tests_k.log
tail -5 gpurun_out/r2_tests_k.log
timeout 300 python tools/eager_profile.py 1 > gpurun_out/r2_eager_profile_k.txt 2>&1
grep -E "quantised|disabled|us$" gpurun_out/r2_eager_profile_k.txt
timeout 300 python tools/eager_profile.py 3 > gpurun_out/r2_eager_profile_k3.txt 2>&1
grep -E "quantised|disabled" gpurun_out/r2_eager_profile_k3.txt
timeout 300 python tools/forward_layout_probe.py > gpurun_out/r2_forward_layout_k.txt 2>&1
cat gpurun_out/r2_forward_layout_k.txt | tail -5
