set -x
cd $GRAFT_REPO_ROOT
for tool in memcheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool python tools/sanitizer_probe.py 2>&1 | tail -4
done > gpurun_out/r2_sanitizer_an.txt 2>&1
cat gpurun_out/r2_sanitizer_an.txt
timeout 300 python tools/qconv_probe.py 2>&1 | tee gpurun_out/r2_qconv_probe_an.txt
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:qconv_igemm" -c 2 -f -o gpurun_out/r2_prof_qconv_an python tools/qconv_profile_probe.py > /dev/null 2> gpurun_out/r2_an.err
FQ_QCONV_2CTA=0 timeout 600 ncu --set full --clock-control none --import-source on -k "regex:qconv_igemm" -c 1 -f -o gpurun_out/r2_prof_qconv_an_1sm python tools/qconv_profile_probe.py > /dev/null 2>> gpurun_out/r2_an.err
ls -la gpurun_out/r2_prof_qconv_an*.ncu-rep
