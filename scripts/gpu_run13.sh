#!/bin/bash
cd "$(dirname "$0")/.."
for V in default narrow p32x16; do
  L=$PWD/microimagelib_b200/lib_alt/libapi_$V.so
  [ $V = default ] && L=$PWD/microimagelib_b200/lib/libapi.so
  AB_SHAPE=512,512,512 AB_TAG="$V 512^3" MILB_LIBAPI=$L timeout 300 python scripts/ab_iter.py
  AB_SHAPE=256,512,512 AB_TAG="$V 512x512x256" MILB_LIBAPI=$L timeout 300 python scripts/ab_iter.py
done
