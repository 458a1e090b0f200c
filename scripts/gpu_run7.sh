#!/bin/bash
cd "$(dirname "$0")/.."
echo "--- gpu tests"
python -m pytest tests -m gpu -q -x 2>&1 | tail -30 > gpurun_out/gpu_tests7.log; tail -15 gpurun_out/gpu_tests7.log
echo "--- fusion, 1 GPU, 16 points"
python bench_fusion.py --points 16 --iters 10 --gpus 1 > gpurun_out/fusion7_p1.json 2> gpurun_out/fusion7_p1.err; tail -c 2500 gpurun_out/fusion7_p1.json; tail -3 gpurun_out/fusion7_p1.err
echo "--- fusion, dispim geometry + 3-D MIPs, 8 points"
python bench_fusion.py --points 8 --iters 10 --gpus 1 --dispim --mip3d --modes resident,sequential_like_reference > gpurun_out/fusion7_dispim.json 2> gpurun_out/fusion7_dispim.err; tail -c 2000 gpurun_out/fusion7_dispim.json; tail -3 gpurun_out/fusion7_dispim.err
