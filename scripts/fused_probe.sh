#!/bin/bash
# Round 2: the fused plane stage (k_planes_fused) vs three launches: ms/iteration, per-kernel times, checksum, and DRAM bytes.
cd "$(dirname "$0")/.."
for SH in "256,512,512"; do
  export AB_SHAPE=$SH
  AB_TAG="unfused $SH" MILB_PLANES_FUSED=0 timeout 300 python scripts/ab_iter.py
  for CFG in "16 0.285,0.46" "8 0.285,0.46" "32 0.285,0.46" "16 0.30,0.44" "16 0.27,0.48" "16 0.25,0.50" "16 0.32,0.40"; do
    set -- $CFG
    AB_TAG="fused ring=$1 split=$2 $SH" MILB_FUSE_RING=$1 MILB_FUSE_SPLIT=$2 timeout 300 python scripts/ab_iter.py
  done
done
AB_SHAPE=512,512,512 AB_TAG="fused 512^3" timeout 300 python scripts/ab_iter.py
AB_SHAPE=512,512,512 AB_TAG="unfused 512^3" MILB_PLANES_FUSED=0 timeout 300 python scripts/ab_iter.py
AB_SHAPE=128,128,128 AB_TAG="fused 128^3" timeout 300 python scripts/ab_iter.py
AB_SHAPE=128,128,128 AB_TAG="unfused 128^3" MILB_PLANES_FUSED=0 timeout 300 python scripts/ab_iter.py
export PROBE_ITERS=2 PROBE_SHAPE=256,512,512
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none \
    -k regex:'k_planes_fused|k_xpassP' --csv --log-file gpurun_out/fused_roles.csv python scripts/prof_run.py > gpurun_out/fused_roles.log 2>&1
grep -c . gpurun_out/fused_roles.csv
