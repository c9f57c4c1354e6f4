# run AU (1 GPU): final tree -- whole GPU suite, full QConv probe, smoke
set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/r2_tests_au.log
cat gpurun_out/r2_tests_au.log
timeout 300 python tools/qconv_probe.py 2>&1 | tee gpurun_out/r2_qconv_probe_au.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
