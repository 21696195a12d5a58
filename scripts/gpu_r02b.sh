#!/bin/bash
# round 2, 2-GPU visit: multi-GPU parity against the oracle, then the bench with the partitioned block
out=gpurun_out/r02b
mkdir -p $out
nvidia-smi --query-gpu=index,name --format=csv > $out/smi.txt 2>&1
nvidia-smi topo -m > $out/topo.txt 2>&1
echo "== multi test"; TL_DEBUG_SHARD=1 timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -30 | tee $out/pytest_multi.txt
echo "== bench 2 gpus"; TL_DEBUG_SHARD=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 2>$out/bench2.err | tee $out/bench2.json
tail -5 $out/bench2.err
echo "== bench 2 gpus nccl transport"; TL_SHARD_TRANSPORT=nccl timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 --partitioned-oracle-check 0 2>$out/bench2_nccl.err | tee $out/bench2_nccl.json
