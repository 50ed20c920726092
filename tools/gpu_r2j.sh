#!/bin/bash
# two GPUs: multi-rank NCCL test, then the driver-style bench launch
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2j_gpus.txt
timeout 900 python -m pytest tests/test_gpu_multirank.py -m gpu -q > gpurun_out/r2j_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2j_pytest.log
tail -5 gpurun_out/r2j_pytest.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2j_bench2.json 2> gpurun_out/r2j_bench2.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/r2j_bench2.err
