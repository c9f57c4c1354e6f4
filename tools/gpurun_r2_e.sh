# run E (2 GPUs): raw-NCCL C-ABI test, QAT data-parallel profile at N=1 and N=2, config 1 graph at N=2
set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_dist.py -x -q 2>&1 | tail -25 > gpurun_out/r2_tests_e_dist.log
cat gpurun_out/r2_tests_e_dist.log
CUDA_VISIBLE_DEVICES=0 timeout 300 python tools/qat_dp_probe.py > gpurun_out/r2_qat_dp_probe_n1.json 2> gpurun_out/r2_e.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/qat_dp_probe.py > gpurun_out/r2_qat_dp_probe_n2.json 2>> gpurun_out/r2_e.err
cat gpurun_out/r2_qat_dp_probe_n1.json gpurun_out/r2_qat_dp_probe_n2.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench_configs.py --config 1 --gpus 2 --graph --steps 50 > gpurun_out/r2_config1_e_n2.json 2>> gpurun_out/r2_e.err
cat gpurun_out/r2_config1_e_n2.json
tail -5 gpurun_out/r2_e.err
