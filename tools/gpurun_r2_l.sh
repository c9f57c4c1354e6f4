set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_api.py tests/test_gpu_qconv_mma.py -m gpu -q -x -k "eager_host or tensor_core" 2>&1 | grep -v "^E   *+\|^E   *and" | tail -80 > gpurun_out/r2_tests_l.log
cat gpurun_out/r2_tests_l.log | cut -c1-400
timeout 300 python tools/eager_profile.py 1 2>&1 | grep -E "quantised|disabled|us$"
