set -x
cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_gpu_qconv_mma.py tests/test_gpu_api.py -m gpu -q -k "qconv or tensor_core" 2>&1 | tail -3
timeout 300 python tools/qconv_probe.py --quick 2>&1 | tee gpurun_out/r2_qconv_probe_o.txt
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:qconv_igemm|qconv_pack_input" -c 4 -f -o gpurun_out/r2_prof_qconv_o python tools/qconv_probe.py --quick > /dev/null 2> gpurun_out/r2_o.err
ls -la gpurun_out/r2_prof_qconv_o.ncu-rep
