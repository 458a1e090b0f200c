#!/bin/bash
# sweep of the fused plane-stage knobs at P = $1 GPUs: chunks x side CTAs -> ms per iteration
cd "$(dirname "$0")/.."
P=${1:-2}
for cfg in "1 0" "2 48" "4 32" "4 48" "4 64" "8 48"; do
  set -- $cfg
  MILB_DSLAB_CHUNKS=$1 MILB_DSLAB_SIDE_CTAS=$2 python -m torch.distributed.run --nnodes=1 --nproc-per-node $P --master-addr 127.0.0.1 \
     --master-port 29513 bench_dist.py --iters 4 --modes fused 2>/dev/null | grep '^{' | \
     python -c "import sys,json; d=json.loads(sys.stdin.read()); print('chunks $1 side $2 ms/iter %.3f' % d['ms_per_iteration'])"
done
