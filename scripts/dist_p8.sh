#!/bin/bash
cd "$(dirname "$0")/.."
P=${1:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $P --master-addr 127.0.0.1 --master-port 29515 bench_dist.py --iters 5 2>/dev/null | grep '^{' > gpurun_out/bench_dist_fused_p$P.json
python -c "
import json; d=json.load(open('gpurun_out/bench_dist_fused_p$P.json'))
print('P=$P fused %.3f ms nccl %.3f ms speedup %.2f bitwise %s' % (d['modes']['fused']['ms_per_iteration'], d['modes']['nccl']['ms_per_iteration'], d['speedup_fused_vs_nccl'], d['modes_agree_bitwise']))"
for cfg in "1 0" "4 64" "4 32"; do
  set -- $cfg
  MILB_DSLAB_CHUNKS=$1 MILB_DSLAB_SIDE_CTAS=$2 python -m torch.distributed.run --nnodes=1 --nproc-per-node $P --master-addr 127.0.0.1 \
     --master-port 29513 bench_dist.py --iters 5 --modes fused 2>/dev/null | grep '^{' | \
     python -c "import sys,json; d=json.loads(sys.stdin.read()); print('chunks $1 side $2 ms/iter %.3f' % d['ms_per_iteration'])"
done
