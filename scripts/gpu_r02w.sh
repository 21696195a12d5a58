#!/bin/bash
# N GPUs (default 8): the bench with the partitioned block
N=${1:-8}
out=gpurun_out/r02w
mkdir -p $out
nvidia-smi topo -m > $out/topo.txt 2>&1
echo "== bench $N gpus"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 2>$out/bench$N.err > $out/bench$N.json
tail -3 $out/bench$N.err
python - $N <<'PY'
import json,sys
N=sys.argv[1]
d=json.loads(open(f'gpurun_out/r02w/bench{N}.json').read().strip().splitlines()[-1])
print('value', d['value'], 'n_gpus', d['n_gpus'], 'e2e', d['e2e']['value'])
p=d['partitioned']
print(json.dumps(p['config4_100k_sharded_triangle'].get('sharded'), indent=1))
print(json.dumps(p['config5_1024_tours'], indent=1))
PY
