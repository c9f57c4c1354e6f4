# run J (1 GPU): whole GPU suite without -x; host-side profile of the eager config-1 forward
set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r2_tests_j.log
tail -5 gpurun_out/r2_tests_j.log
timeout 300 python tools/eager_profile.py > gpurun_out/r2_eager_profile_j.txt 2>&1
tail -60 gpurun_out/r2_eager_profile_j.txt
