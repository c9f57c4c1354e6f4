# run F (2 GPUs): raw-NCCL C-ABI test (fixed), tcgen05 conv tests, QAT DP kernel diff with aligned / packed gradient views
set -x
cd $GRAFT_REPO_ROOT
CUDA_VISIBLE_DEVICES=0 timeout 300 python -m pytest tests/test_gpu_qconv_mma.py -x -q 2>&1 | tail -30 > gpurun_out/r2_tests_f_mma.log
cat gpurun_out/r2_tests_f_mma.log
timeout 400 python -m pytest tests/test_gpu_dist.py -x -q -k "c_abi" 2>&1 | tail -30 > gpurun_out/r2_tests_f_dist.log
cat gpurun_out/r2_tests_f_dist.log
CUDA_VISIBLE_DEVICES=0 timeout 200 python -m pytest tests/test_gpu_kernels.py -x -q -k "from_maxima or every_execution_mode" 2>&1 | tail -5
CUDA_VISIBLE_DEVICES=0 timeout 200 python tools/qat_dp_probe.py 2> gpurun_out/r2_f.err | cut -c1-400
for al in 32 1; do
FQ_BUCKET_ALIGN=$al timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$al tools/qat_dp_probe.py 2>> gpurun_out/r2_f.err | cut -c1-400
done
python - <<'PY'
import json
a=json.load(open("gpurun_out/r2_qat_kernels_n1_align32.json"))
for al in (32,1):
    b=json.load(open("gpurun_out/r2_qat_kernels_n2_align%d.json"%al))
    rows=[(b.get(k,[0,0])[0]-a.get(k,[0,0])[0], b.get(k,[0,0])[1]-a.get(k,[0,0])[1], k) for k in set(a)|set(b)]
    rows.sort(key=lambda r:-abs(r[1]))
    print("=== N=2 align %d minus N=1: count diff, us diff, kernel"%al)
    for r in rows[:14]: print(round(r[0],1), round(r[1],1), r[2][:170])
PY
tail -3 gpurun_out/r2_f.err
