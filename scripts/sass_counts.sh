#!/bin/bash
# Static SASS evidence of the hot kernels (no GPU needed): per kernel, how many TMA / cp.async / texture / tensor-core
# instructions the sm_100a code contains.  Usage: scripts/sass_counts.sh [tag]   -> profiles/<tag>_sass_counts.txt
cd "$(dirname "$0")/.."
TAG=${1:-r02}
OUT=profiles/${TAG}_sass_counts.txt
{
echo "# cuobjdump -sass of microimagelib_b200/build/*.o (sm_100a), instruction counts per kernel"
echo "# UTMALDG = TMA tile load (cp.async.bulk.tensor), SYNCS = mbarrier, LDGSTS = cp.async, TEX = texture fetch, BAR = CTA barrier,"
echo "# MEMBAR/ATOMG/REDG = release / acquire counters of the fused plane stage and of the plane pipeline, UBLKCP = bulk copy (cp.async.bulk) of k_zrow.  No UTC*MMA / HMMA: nothing on this path is a contraction."
printf "%-78s %6s %7s %6s %5s %6s %4s %5s %6s %5s %5s %5s\n" kernel instr UTMALDG UBLKCP SYNCS LDGSTS TEX BAR MEMBAR ATOM MUFU HMMA
for f in microimagelib_b200/build/decon_fast_n512.cu.o microimagelib_b200/build/decon_fast_n256.cu.o microimagelib_b200/build/decon_fast_n320.cu.o microimagelib_b200/build/decon_fast_n1024.cu.o microimagelib_b200/build/reg.cu.o microimagelib_b200/build/geom.cu.o; do
  cuobjdump -sass $f 2>/dev/null | awk -v file=$(basename $f) '
    /Function :/ { name=$3 }
    /^ +\/\*[0-9a-f]+\*\/ / { n[name]++; if ($0 ~ /UTMALDG/) a[name]++; if ($0 ~ /UBLKCP/) u[name]++; if ($0 ~ /SYNCS/) b[name]++; if ($0 ~ /LDGSTS/) c[name]++; if ($0 ~ / TEX| TLD/) d[name]++;
                               if ($0 ~ / BAR\./) e[name]++; if ($0 ~ /MEMBAR/) g[name]++; if ($0 ~ /ATOMG|REDG|ATOM\./) h[name]++; if ($0 ~ /MUFU/) m[name]++; if ($0 ~ /HMMA|UTC.MMA/) t[name]++ }
    END { for (k in n) if (n[k] > 300) printf "%-78s %6d %7d %6d %5d %6d %4d %5d %6d %5d %5d %5d\n", substr(file ":" k, 1, 78), n[k], a[k], u[k], b[k], c[k], d[k], e[k], g[k], h[k], m[k], t[k] }' | sort
done
} > $OUT
c++filt < $OUT > $OUT.tmp && mv $OUT.tmp $OUT
wc -l $OUT
