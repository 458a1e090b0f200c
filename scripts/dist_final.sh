#!/bin/bash
cd "$(dirname "$0")/.."
P=${1:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $P --master-addr 127.0.0.1 --master-port 29515 bench_dist.py --iters 5 2>/dev/null | grep '^{' > gpurun_out/bench_dist_fused_p$P.json
python -c "
import json; d=json.load(open('gpurun_out/bench_dist_fused_p$P.json'))
print('P=$P fused %.3f ms nccl %.3f ms speedup %.2f bitwise %s value %.3e' % (d['modes']['fused']['ms_per_iteration'], d['modes']['nccl']['ms_per_iteration'], d['speedup_fused_vs_nccl'], d['modes_agree_bitwise'], d['value']))"
