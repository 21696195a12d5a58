#!/bin/bash
# 2 GPUs: teardown-order probe for a context with NCCL + peer mailboxes, then the multi-GPU parity test
out=gpurun_out/r02l
mkdir -p $out
for v in no_close close_after close_first; do
  echo "== teardown $v"
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 scripts/teardown_probe.py $v > $out/teardown_$v.log 2>&1
  echo "rc=$?" | tee -a $out/teardown_$v.log
  grep -E "done|closed|Fatal|Segmentation|File \"|rc=" $out/teardown_$v.log | head -20
done
echo "== multi test"; timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -15 | tee $out/pytest_multi.txt
