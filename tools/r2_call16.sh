#!/bin/bash
# GPU call 16 (2 GPUs): the two-phase image-set path (image-sharded extraction -> NCCL all_gather of the float16 feature blocks -> pair-sharded
# matching -> gather) and the cfg5 / headline workloads on two ranks.
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 500 $TR --master-port 29511 bench.py --gpus 2 --mode exhaustive > gpurun_out/r2_mode_exhaustive_2gpu.json 2> gpurun_out/r2_mode_exhaustive_2gpu.err; tail -c 400 gpurun_out/r2_mode_exhaustive_2gpu.err; head -c 1500 gpurun_out/r2_mode_exhaustive_2gpu.json
timeout 300 $TR --master-port 29512 bench.py --gpus 2 --mode nn > gpurun_out/r2_mode_nn_2gpu.json 2> gpurun_out/r2_mode_nn_2gpu.err; tail -c 300 gpurun_out/r2_mode_nn_2gpu.err; head -c 700 gpurun_out/r2_mode_nn_2gpu.json
timeout 300 $TR --master-port 29513 bench.py --gpus 2 --quick --steps 10 --warmup 3 > gpurun_out/r2_quick_2gpu.json 2> gpurun_out/r2_quick_2gpu.err; tail -c 300 gpurun_out/r2_quick_2gpu.err; cat gpurun_out/r2_quick_2gpu.json
