set -x
cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_gpu_qconv_mma.py tests/test_gpu_api.py -m gpu -q -k "qconv or tensor_core" 2>&1 | tail -2
timeout 200 python tools/qconv_probe.py 2>&1 | tee gpurun_out/r2_qconv_probe_ab.txt
for tool in memcheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool python tools/sanitizer_probe.py 2>&1 | tail -6
done > gpurun_out/r2_sanitizer_ab.txt 2>&1
cat gpurun_out/r2_sanitizer_ab.txt
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:qconv_igemm|qconv_pack_input" -c 3 -f -o gpurun_out/r2_prof_qconv_ab python tools/qconv_profile_probe.py > /dev/null 2> gpurun_out/r2_ab.err
ls -la gpurun_out/r2_prof_qconv_ab.ncu-rep
