#!/bin/bash
# Round 2: the fused plane stage (k_planes_fused) vs three launches: ms/iteration, per-kernel times, checksum, and DRAM bytes.
cd "$(dirname "$0")/.."
for SH in "256,512,512" "512,512,512"; do
  export AB_SHAPE=$SH
  AB_TAG="unfused $SH" MILB_PLANES_FUSED=0 timeout 300 python scripts/ab_iter.py
  for G in 2 4 8; do
    AB_TAG="fused G=$G $SH" MILB_FUSE_GROUP=$G timeout 300 python scripts/ab_iter.py
  done
done
export PROBE_ITERS=2 PROBE_SHAPE=256,512,512
for G in 4; do
MILB_FUSE_GROUP=$G timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none \
    -k regex:'k_planes_fused|k_xpassP' --csv --log-file gpurun_out/fused_g$G.csv python scripts/prof_run.py > gpurun_out/fused_g$G.log 2>&1
grep -c . gpurun_out/fused_g$G.csv
done
