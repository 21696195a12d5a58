#!/bin/bash
# 4 GPUs: multi-GPU parity test against the oracle (world 2 and 4), then the bench with the partitioned block
out=gpurun_out/r04d
mkdir -p $out
echo "== multi test"; timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -5 | tee $out/pytest_multi.txt
echo "== bench 4 gpus"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 20 --warmup 3 2>$out/bench4.err > $out/bench4.json
tail -3 $out/bench4.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r04d/bench4.json').read().strip().splitlines()[-1])
print('value', d['value'], 'n_gpus', d['n_gpus'], 'e2e', d['e2e']['value'])
p=d['partitioned']
print(json.dumps(p['config4_100k_sharded_triangle'].get('sharded'), indent=1))
print(json.dumps(p['config5_1024_tours'].get('sharded'), indent=1))
PY
