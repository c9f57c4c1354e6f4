set -x
cd $GRAFT_REPO_ROOT
timeout 200 python -m pytest tests/test_gpu_qconv_mma.py tests/test_gpu_api.py -m gpu -q -k "qconv or tensor_core" 2>&1 | grep -v "^E   *+\|^E   *and" | tail -5 | cut -c1-250
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:qconv_igemm" -s 3 -c 1 -f -o gpurun_out/r2_prof_qconv_q python tools/qconv_probe.py --quick > /dev/null 2> gpurun_out/r2_q.err
ls -la gpurun_out/r2_prof_qconv_q.ncu-rep
