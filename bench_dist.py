#!/usr/bin/env python
"""bench_dist.py -- BASELINE config 3: deconDualView joint RL on ONE 1024x1024x512 pair, slab-decomposed
distributed 3-D FFT with NCCL all-to-all at P = 2/4/8 B200 (strong scaling of a single volume).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node P --master-addr 127.0.0.1 \
        --master-port 29500 bench_dist.py --iters 5

Prints one JSON line on rank 0.  Headline: the FUSED exchange (the all-to-alls folded into the X-pass
and Y-inverse kernels' stores over NVLink peer memory) -- ms per dual-view iteration and
voxel-iterations/s.  Beside it the NCCL baseline (all_to_all_single + re-layout copies): its ms per
iteration, the all-to-all's bytes, stand-alone time and achieved NVLink GB/s per direction per GPU
(against the measured 770 GB/s peer-copy figure of B200_PROFILING.md) and its share of the iteration.
Timing: CUDA events on the launching stream, max over ranks.  Data: synthetic (uniform noise on a
background, generated on the device: the loop's cost does not depend on the voxel values).
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="512,1024,1024", help="slices,H,W of the FFT box")
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--views", type=int, default=2)
    ap.add_argument("--modes", default="fused,nccl", help="comma list of exchange implementations to time")
    ap.add_argument("--profile", action="store_true", help="rank 0 prints a per-kernel time table (torch.profiler / CUPTI) of 2 iterations per mode")
    args = ap.parse_args()
    import numpy as np
    import torch
    import torch.distributed as dist
    from microimagelib_b200 import synth
    from microimagelib_b200.dist_decon import DistDecon

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    shape = tuple(int(v) for v in args.shape.split(","))
    psf_a = synth.gaussian_psf((65, 65, 65), (4, 2, 2))
    psf_b = synth.gaussian_psf((65, 65, 65), (2, 2, 4))

    def timed(fn, reps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b) / reps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def build(fused):
        dd = DistDecon(shape, args.views, fused=fused)
        dd.set_psf(0, psf_a)
        if args.views == 2:
            dd.set_psf(1, psf_b)
        g = torch.Generator(device=dev)
        g.manual_seed(20260 + rank)
        for v in range(args.views):
            dd.set_image(v, torch.rand((dd.L.X, dd.L.ny, dd.L.Z), generator=g, device=dev) * 50 + 100)
        return dd

    res = {}
    checks = {}
    for mode in [m for m in args.modes.split(",") if m]:
        dd = build(mode == "fused")
        if mode == "fused" and not dd.fused:
            dd.close()
            continue
        L = dd.L
        dd.run(args.warmup)
        ms_iter = timed(lambda: dd.run(args.iters), 1) / args.iters
        res[mode] = {"ms_per_iteration": ms_iter}
        checks[mode] = dd.E.double().sum().item()      # same inputs, same kernels: the two modes must agree exactly
        if args.profile:
            from torch.profiler import ProfilerActivity, profile
            if world > 1:
                dist.barrier()
            with profile(activities=[ProfilerActivity.CUDA]) as prof:
                dd.run(2)
                torch.cuda.synchronize()
            if rank == 0:
                print(f"---- {mode}: kernels of 2 iterations on rank 0", flush=True)
                print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=60), flush=True)
        if mode == "nccl":
            def exchange_only():
                L.to_planes(dd.slab, dd.planes, dd.scratch)
                L.to_slabs(dd.planes, dd.slab, dd.scratch)

            exchange_only()
            ms_pair = timed(exchange_only, 3)             # two all-to-alls + the two local re-layout copies

            def a2a_only():
                if world > 1:
                    dist.all_to_all_single(dd.scratch[: world * L.np * L.row], dd.slab.reshape(-1), L.plane_splits(), L.slab_splits())

            a2a_only()
            ms_a2a = timed(a2a_only, 5) if world > 1 else 0.0
            nconv = 2 * args.views
            sent = dd.a2a_bytes_per_gpu()
            res[mode]["all_to_all"] = {
                "per_iteration": 2 * nconv, "bytes_sent_per_gpu_each": sent, "ms_each_standalone": ms_a2a,
                "GBps_per_direction_per_gpu": (sent / (ms_a2a * 1e-3) / 1e9) if ms_a2a else None,
                "nvlink_peak_GBps": 770.0, "frac_of_peak": (sent / (ms_a2a * 1e-3) / 1e9 / 770.0) if ms_a2a else None,
                "ms_exchange_pair_with_relayout": ms_pair, "share_of_iteration": nconv * ms_pair / ms_iter}
        else:
            nconv = 2 * args.views
            sent = dd.a2a_bytes_per_gpu()
            res[mode]["peer_store_bytes_per_gpu_per_exchange"] = sent
            res[mode]["exchanges_per_iteration"] = 2 * nconv
            # lower bound of an iteration if NVLink egress were the only cost: every exchange at the peer-copy peak
            res[mode]["nvlink_floor_ms_per_iteration"] = 2 * nconv * sent / 770e9 * 1e3
        dd.close()
        del dd
        torch.cuda.empty_cache()
    if rank == 0:
        nfft = float(np.prod(shape))
        head = "fused" if "fused" in res else "nccl"
        ms_iter = res[head]["ms_per_iteration"]
        line = {
            "metric": "RL voxel-iters/sec (one volume, slab-decomposed distributed FFT)", "value": nfft / (ms_iter * 1e-3),
            "unit": "voxel-iters/s", "n_gpus": world, "scaling": "strong", "ms_per_iteration": ms_iter, "iterations": args.iters,
            "exchange": head,
            "config": {"workload": f"deconDualView joint RL {shape[2]}x{shape[1]}x{shape[0]} pair (BASELINE config 3)" if args.views == 2
                       else f"deconSingleView RL {shape[2]}x{shape[1]}x{shape[0]}", "views": args.views,
                       "decomposition": "real volumes by rows (y), spectrum by whole kx-planes; 2 exchanges per convolution",
                       "fused": "exchange folded into the X-pass / Y-inverse kernels' stores (NVLink peer memory via CUDA IPC); "
                                "1-element NCCL all-reduce as phase barrier",
                       "nccl": "all_to_all_single + re-layout copies"},
            "modes": res,
            "modes_agree_bitwise": (len(set(checks.values())) == 1) if len(checks) > 1 else None,
            "speedup_fused_vs_nccl": (res["nccl"]["ms_per_iteration"] / res["fused"]["ms_per_iteration"]) if len(res) == 2 else None,
            "dtype": "f32", "data": "synthetic (device-generated noise on background; timing only)",
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
