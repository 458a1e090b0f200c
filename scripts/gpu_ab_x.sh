#!/bin/bash
# A/B of two builds of the library (default vs lib/libapi_<tag>.so) with scripts/ab_iter.py
cd "$(dirname "$0")/.."
TAG=${1:-xnarrow}
for s in 512,512,512 256,512,512; do
  AB_SHAPE=$s AB_TAG=default timeout 120 python scripts/ab_iter.py 2>&1 | tail -1
  AB_SHAPE=$s AB_TAG=$TAG MILB_LIBAPI=$PWD/microimagelib_b200/lib/libapi_$TAG.so timeout 120 python scripts/ab_iter.py 2>&1 | tail -1
done
