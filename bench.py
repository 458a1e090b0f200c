#!/usr/bin/env python
"""bench.py -- Richardson-Lucy voxel-iterations/s on B200 (BASELINE.json metric).

    python bench.py --gpus 1 --steps 3 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the CPU arm (oracle port of the reference's FFTW path)

One "step" = one pass of the hot path over one synthetic volume: the full iteration loop of
deconSingleView (BASELINE config 2: 512x512x256 float32 beads + Gaussian PSF, 50 iterations).
  value : N_fft * iterations * steps * ranks / device time, inputs resident in HBM (CUDA events)
  e2e   : the same metric through the reference-facing call libapi.decon_singleview with HOST
          buffers: H2D of the image, the loop, D2H of the result inside the timed region
  roofline : algorithmic bytes of the loop (56 * N_fft per single-view iteration)
          / measured loop time, against the measured HBM copy peak (MEASURED_PEAKS.json)
  cpu_baseline : the oracle's numpy/pocketfft port of decon_singleview_OTF0 on the host cores
N > 1: every rank deconvolves its own volume (time points of spimFusionBatch shard with no
data-path collective) -> weak scaling; time = max over ranks.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALG_BYTES_PER_VOXEL_ITER = 56.0  # single-view RL iteration, float32


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--shape", default="256,512,512", help="slices,H,W")
    ap.add_argument("--iters", type=int, default=50)
    ap.add_argument("--psf", type=int, default=65)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-yardstick", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--chunk-planes", type=int, default=-1)
    return ap.parse_args()


def make_inputs(shape, psf_n, rank):
    import numpy as np
    from microimagelib_b200 import synth
    psf = synth.gaussian_psf((psf_n,) * 3, (4.0, 2.0, 2.0))
    img = synth.bead_image(shape, psf, seed=synth.SEED_A + 1000 * rank, noise_seed=synth.SEED_NOISE + 1000 * rank)
    return img, psf


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx = float(f[2])
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def cpu_port_rate(img, psf, iters, threads=None):
    """The oracle's port of the reference CPU loop, timed on the host: voxel-iters/s."""
    import numpy as np
    from oracle import decon_oracle as do
    if threads:
        os.environ["MILB_ORACLE_THREADS"] = str(threads)
    fshape = do.fft_shape_for(img.shape)
    otf, otf_bp = do.gen_otf_pair(psf, fshape)
    A = np.maximum(img, do.SMALLVALUE)
    t0 = time.perf_counter()
    do.rl_single(A, otf, otf_bp, iters)
    dt = time.perf_counter() - t0
    return float(np.prod(fshape)) * iters / dt, dt


def run_reference(args, shape, rank, world):
    """--impl reference: the reference's CPU implementation of the path.  Its FFTW host path cannot
    be built here, so this is the oracle port (kind "port") with all host threads."""
    if rank != 0:
        return
    import numpy as np
    cores = os.cpu_count() or 1
    img, psf = make_inputs(shape, args.psf, 0)
    sample_iters = 1
    for _ in range(min(args.warmup, 1)):
        cpu_port_rate(img, psf, sample_iters, cores)
    rates, times = [], []
    for _ in range(args.steps):
        r, dt = cpu_port_rate(img, psf, sample_iters, cores)
        rates.append(r)
        times.append(dt)
    n_fft = float(np.prod(shape))
    value = n_fft * sample_iters * len(times) / sum(times)
    line = {
        "impl": "reference", "metric": "RL voxel-iters/sec", "value": value, "unit": "voxel-iters/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"deconSingleView RL {shape[2]}x{shape[1]}x{shape[0]} float32, {args.iters} iterations (BASELINE config 2)",
                   "psf": f"{args.psf}^3 Gaussian", "sampled": f"{sample_iters} of the {args.iters} iterations per step (rate does not depend on the iteration index)"},
        "cpu_baseline": {"value": value, "unit": "voxel-iters/s", "cores": cores, "kind": "port",
                         "sample": f"{sample_iters} RL iteration(s) of the same volume per step, numpy + scipy.fft (pocketfft) float32"},
        "e2e": {"value": value, "unit": "voxel-iters/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def bind_to_gpu_numa_node(index):
    """Pin this process to the CPUs next to its GPU (what `numactl` would do for a user), so that the
    pinned host buffers of the end-to-end leg are allocated on the NUMA node the GPU hangs off."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        pynvml.nvmlDeviceSetCpuAffinity(h)
        return sorted(os.sched_getaffinity(0))[:2] + [len(os.sched_getaffinity(0))]
    except Exception as e:  # best effort
        return str(e)


def main():
    args = parse()
    shape = tuple(int(s) for s in args.shape.split(","))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, shape, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    from microimagelib_b200 import device, libapi

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    numa = bind_to_gpu_numa_node(local_rank)   # the end-to-end leg moves 2 x 256 MiB over PCIe per step
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    img, psf = make_inputs(shape, args.psf, rank)
    n_img = float(np.prod(shape))
    stream = torch.cuda.current_stream()
    d = device.Decon(shape, 1)
    if args.chunk_planes >= 0:
        d.set_chunk_planes(args.chunk_planes)
    n_fft = float(np.prod(d.fft_shape))
    d.set_psf(0, psf)
    d_img = torch.from_numpy(img).cuda()
    d.set_image(0, d_img, stream)

    # ---- device-resident loop ------------------------------------------------------------------
    for _ in range(args.warmup):
        d.run(args.iters, stream=stream)
    barrier()
    clocks = ClockSampler(local_rank)
    clocks.start()
    l0 = device.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    for _ in range(args.steps):
        d.run(args.iters, stream=stream)
    ev1.record(stream)
    barrier()
    launches = device.launch_count() - l0
    ms = ev0.elapsed_time(ev1)
    clk = clocks.stop()
    t = torch.tensor([ms], dtype=torch.float64, device=torch.device("cuda", local_rank))
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = n_fft * args.iters * args.steps * world / (ms_max * 1e-3)

    # ---- end to end through the reference-facing API (host buffers) -----------------------------
    e2e = None
    if not args.no_e2e:
        h_img = torch.from_numpy(img).pin_memory().numpy()
        h_out = torch.empty(shape, dtype=torch.float32).pin_memory().numpy()
        for _ in range(2):
            libapi.decon_singleview(h_img, psf, args.iters, deviceNum=local_rank, out=h_out)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            out, st, rec = libapi.decon_singleview(h_img, psf, args.iters, deviceNum=local_rank, out=h_out)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=torch.device("cuda", local_rank))
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        xt = torch.empty(shape, dtype=torch.float32, device="cuda")
        hp = torch.from_numpy(h_img)
        xt.copy_(hp)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        xt.copy_(hp, non_blocking=True)
        torch.cuda.synchronize()
        h2d_gbps = n_img * 4 / (time.perf_counter() - t0) / 1e9
        del xt
        e2e = {"value": n_fft * args.iters * args.steps * world / float(tt.item()), "unit": "voxel-iters/s",
               "pcie_h2d_GBps_pinned": h2d_gbps, "cpu_affinity": numa,
               "last_call_records_ms": {"init": float(rec[6]) * 1e3, "cache_check_h2d_pad": float(rec[7]) * 1e3,
                                        "loop_crop_d2h": float(rec[8]) * 1e3, "total": float(rec[9]) * 1e3},
               "h2d_bytes_per_step": int(n_img * 4), "d2h_bytes_per_step": int(n_img * 4),
               "ms_per_step": 1e3 * float(tt.item()) / args.steps,
               "call": "libapi.decon_singleview(host float32 image, host PSF) -> host float32 volume; OTFs cached across calls"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- yardstick + CPU baseline (rank 0, N = 1 only) ------------------------------------------
    peak, peak_src = peak_hbm()
    ms_iter = ms_max / (args.steps * args.iters)
    achieved = ALG_BYTES_PER_VOXEL_ITER * n_fft / (ms_iter * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get("dram_bytes_per_iteration")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "peak_source": peak_src, "launch": "one single-view RL iteration = 6 plane-pass launches + 2 fused X-pass launches",
                "algorithmic_bytes_per_launch": ALG_BYTES_PER_VOXEL_ITER * n_fft, "ms_per_launch": ms_iter}
    if world == 1:
        try:                                             # per-kernel break-down: kernel-level bytes / measured launch time
            kms = d.time_kernels(5, stream=stream)
            nspec = float(d.fft_shape[0] // 2 + 1) * d.fft_shape[1] * d.fft_shape[2]
            rows = [("k_ypassT (Y forward, transposing)", 16 * nspec, kms[0], 2), ("k_zconvT (Z forward * OTF, Z inverse)", 24 * nspec, kms[1], 2),
                    ("k_ypassF (Y inverse)", 16 * nspec, kms[2], 2), ("k_xpassP ratio (C2R, A/T, R2C)", 16 * nspec + 4 * n_fft, kms[3], 1),
                    ("k_xpassP update (C2R, E*T clamp, R2C)", 16 * nspec + 8 * n_fft, kms[4], 1)]
            roofline["kernels"] = [{"kernel": k, "launches_per_iteration": n, "bytes_per_launch": b, "ms_per_launch": float(ms),
                                    "GBps": b / (float(ms) * 1e-3) / 1e9, "frac_of_peak": b / (float(ms) * 1e-3) / 1e9 / peak}
                                   for k, b, ms, n in rows]
            roofline["kernels_note"] = "kernel-level HBM bytes (what each launch must read + write), CUDA events around every launch"
        except Exception as e:
            roofline["kernels"] = {"error": str(e)}
    registration = None
    if world == 1:
        try:   # the other kernel of the hot path: fused warp + ZNCC cost, same volume size (8 N bytes per evaluation)
            m = np.array([0.9994, 0.0349, 0, -5.1, -0.0349, 0.9994, 0, 6.3, 0, 0, 1, 1.75], np.float32)   # 2 deg about z + shift
            r = device.Reg(shape)
            r.set_images(d_img, d_img)
            r.prepare()
            registration = {"what": "k_zncc: trilinear warp + ZNCC sums (double accumulation), K candidate matrices per launch",
                            "algorithmic_bytes_per_evaluation": 8 * n_img}
            for K in (1, 8):
                mats = np.stack([m] * K)
                mats[:, 3] += 0.1 * np.arange(K, dtype=np.float32)
                r.cost(mats)
                torch.cuda.synchronize()
                ta, tb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                ta.record(stream)
                for _ in range(5):
                    r.cost(mats, stream=stream)
                tb.record(stream)
                torch.cuda.synchronize()
                ms_eval = ta.elapsed_time(tb) / 5 / K
                registration[f"K{K}"] = {"ms_per_evaluation": ms_eval, "GBps": 8 * n_img / (ms_eval * 1e-3) / 1e9,
                                         "frac_of_peak": 8 * n_img / (ms_eval * 1e-3) / 1e9 / peak}
            registration["note"] = "issue-bound (ncu: 78 % issue slots, DRAM 18 %); includes the per-launch D2H of the 2K sums"
            r.close()
        except Exception as e:
            registration = {"error": str(e)}
    yard = None
    if world == 1 and not args.no_yardstick:
        try:
            yi = max(5, args.iters // 5)
            yms = d.run_cufft_yardstick(yi, stream=stream) / yi     # CUDA events around the loop only
            yard = {"what": "cuFFT R2C/C2R + unfused element-wise kernels (the reference's launch structure)", "ms_per_iteration": yms,
                    "voxel_iters_per_s": n_fft / (yms * 1e-3), "speedup_vs_yardstick": yms / ms_iter}
        except Exception as e:  # the yardstick is informative only
            yard = {"error": str(e)}
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        si = 2
        rate, dt = cpu_port_rate(img, psf, si, cores)
        cpu = {"value": rate, "unit": "voxel-iters/s", "cores": cores, "kind": "port",
               "sample": f"{si} RL iterations of the same {shape[2]}x{shape[1]}x{shape[0]} volume ({dt:.1f} s), numpy + scipy.fft "
                         "(pocketfft) float32 restatement of decon_singleview_OTF0; the reference's FFTW path cannot be built here"}
    line = {
        "metric": "RL voxel-iters/sec", "value": value, "unit": "voxel-iters/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"deconSingleView RL {shape[2]}x{shape[1]}x{shape[0]} float32, {args.iters} iterations (BASELINE config 2)",
                   "fft_box": list(d.fft_shape), "psf": f"{args.psf}^3 Gaussian", "per_rank": "one volume per rank, no data-path collective",
                   "l2": "inputs larger than L2 (256 MiB volume, 126 MB L2); no explicit flush"},
        "clocks": clk, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu, "yardstick": yard, "registration": registration,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
