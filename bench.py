#!/usr/bin/env python
"""bench.py -- Richardson-Lucy voxel-iterations/s on B200 (BASELINE.json metric).

    python bench.py --gpus 1 --steps 3 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the CPU arm (oracle port of the reference's FFTW path)

One "step" = one pass of the hot path over one synthetic volume: the full iteration loop of
deconSingleView on the metric's own configuration, 512^3 float32 beads + Gaussian PSF, 50 iterations
(BASELINE.json metric "RL voxel-iters/sec at 512^3"); BASELINE config 2 (512x512x256) rides along as a
second record ("config2") of the same JSON line.
  value : N_fft * iterations * steps * ranks / device time, inputs resident in HBM (CUDA events)
  e2e   : the same metric through the reference-facing call libapi.decon_singleview with HOST
          buffers: H2D of the image, the loop, D2H of the result inside the timed region
  roofline : algorithmic bytes of the loop (56 * N_fft per single-view iteration)
          / measured loop time, against the measured HBM copy peak (MEASURED_PEAKS.json)
  cpu_baseline : the oracle's numpy/pocketfft port of decon_singleview_OTF0 on the host cores; its
          output doubles as the in-run parity check ("parity": rel-L2 of the CUDA result against it)
  traffic : DRAM bytes per iteration measured in this run by an ncu side-run of the same loop
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
    ap.add_argument("--shape", default="512,512,512", help="slices,H,W")
    ap.add_argument("--shape2", default="256,512,512", help="second record (BASELINE config 2); empty = skip")
    ap.add_argument("--no-config3", action="store_true", help="skip the config-3 record (one 1024x1024x512 dual-view pair: single GPU at N=1, "
                    "slab-decomposed distributed FFT over all ranks at N>1)")
    ap.add_argument("--config3-shape", default="512,1024,1024", help="slices,H,W of the config-3 FFT box")
    ap.add_argument("--config3-iters", type=int, default=5)
    ap.add_argument("--no-traffic", action="store_true", help="skip the ncu side-run that measures DRAM bytes per iteration")
    ap.add_argument("--no-refgpu", action="store_true", help="skip the reference's own GPU path (oracle/_ref) as an in-run yardstick")
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


def snap(n):
    """snapTransformSize (reference src/api_subfunc.cu:57-87): the FFT extent of an image extent."""
    n = (n + 15) // 16 * 16
    low = 1 << (n.bit_length() - 1)
    if low == n:
        return n
    return low * 2 if low * 2 <= 128 else (n + 63) // 64 * 64


def workload_config(shape, iters, psf_n):
    """The `config` object both arms print (same keys, same values)."""
    tag = "the metric's 512^3 configuration" if tuple(shape) == (512, 512, 512) else \
        "BASELINE config 2" if tuple(shape) == (256, 512, 512) else "custom shape"
    return {"workload": f"deconSingleView RL {shape[2]}x{shape[1]}x{shape[0]} float32, {iters} iterations ({tag})",
            "fft_box": [snap(s) for s in shape], "psf": f"{psf_n}^3 Gaussian", "iterations": iters,
            "per_rank": "one volume per rank, no data-path collective",
            "l2": "inputs larger than L2 (>= 256 MiB per volume, 126 MB L2); no explicit flush"}


def cpu_port_rate(img, psf, iters, threads=None, want_result=False):
    """The oracle's port of the reference CPU loop, timed on the host: voxel-iters/s."""
    import numpy as np
    from oracle import decon_oracle as do
    if threads:
        os.environ["MILB_ORACLE_THREADS"] = str(threads)
    fshape = do.fft_shape_for(img.shape)
    otf, otf_bp = do.gen_otf_pair(psf, fshape)
    A = np.maximum(do.pad_stack(img, fshape) if tuple(fshape) != tuple(img.shape) else img, do.SMALLVALUE)
    t0 = time.perf_counter()
    E = do.rl_single(A, otf, otf_bp, iters)
    dt = time.perf_counter() - t0
    rate = float(np.prod(fshape)) * iters / dt
    if want_result:
        return rate, dt, (do.crop_stack(E, img.shape) if tuple(fshape) != tuple(img.shape) else E)
    return rate, dt


def run_reference(args, shape, rank, world):
    """--impl reference: the reference's CPU implementation of the path.  Its FFTW host path cannot
    be built here, so this is the oracle port (kind "port") with all host threads."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    # torch.distributed.run exports OMP_NUM_THREADS=1 to its workers, which makes numpy / pocketfft 2.7x slower than
    # the same call from a shell; the CPU arm always uses every host core (set before numpy is imported)
    for k in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS"):
        os.environ[k] = str(cores)
    import numpy as np
    img, psf = make_inputs(shape, args.psf, 0)
    sample_iters = 1
    for _ in range(min(args.warmup, 1)):
        cpu_port_rate(img, psf, sample_iters, cores)
    rates, times = [], []
    for _ in range(args.steps):
        r, dt = cpu_port_rate(img, psf, sample_iters, cores)
        rates.append(r)
        times.append(dt)
    n_fft = float(np.prod([snap(s) for s in shape]))
    value = n_fft * sample_iters * len(times) / sum(times)
    line = {
        "impl": "reference", "metric": "RL voxel-iters/sec", "value": value, "unit": "voxel-iters/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(shape, args.iters, args.psf),
        "cpu_baseline": {"value": value, "unit": "voxel-iters/s", "cores": cores, "kind": "port",
                         "sample": f"{sample_iters} of the {args.iters} RL iterations of the same volume per step (the rate does not depend on the "
                                   "iteration index), numpy + scipy.fft (pocketfft) float32; one host, all cores: does not scale with --gpus"},
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


def measure_traffic(shape, timeout_s=240):
    """DRAM bytes per RL iteration of THIS build, measured now: an ncu side-run (dram__bytes_read.sum +
    dram__bytes_write.sum per launch) of scripts/prof_run.py, 3 iterations, the middle one is reported."""
    import csv
    import shutil
    import tempfile
    ncu = shutil.which("ncu") or "/usr/local/cuda/bin/ncu"
    if not os.path.exists(ncu):
        return None, "ncu not found"
    tmp = tempfile.mkdtemp(prefix="milb_traffic_")
    log = os.path.join(tmp, "t.csv")
    env = dict(os.environ, PROBE_ITERS="3", PROBE_SHAPE=",".join(str(s) for s in shape), PROBE_NOISE="0")
    cmd = [ncu, "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum", "--clock-control", "none",
           "-k", "regex:k_planes_fused|k_ypassT|k_zconvT|k_zrow|k_ypassF|k_xpassP|k_xpass|k_ypass|k_zpass", "--csv", "--log-file", log,
           sys.executable, os.path.join(ROOT, "scripts", "prof_run.py")]
    try:
        r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=timeout_s)
    except Exception as e:
        return None, f"ncu side-run failed: {e}"
    if r.returncode != 0 or not os.path.exists(log):
        return None, "ncu side-run failed: " + (r.stderr or r.stdout)[-200:]
    rows = []
    with open(log) as f:
        lines = [ln for ln in f if ln.startswith('"')]
    per = {}
    order = []
    for row in csv.DictReader(lines):
        i = int(row["ID"])
        if i not in per:
            per[i] = {"kernel": row["Kernel Name"], "bytes": 0.0, "us": None}
            order.append(i)
        v = float(row["Metric Value"].replace(",", ""))
        unit = row.get("Metric Unit", "")
        if row["Metric Name"].startswith("dram__bytes"):
            per[i]["bytes"] += v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
        else:
            per[i]["us"] = v * {"ns": 1e-3, "us": 1, "ms": 1e3, "nsecond": 1e-3, "usecond": 1, "msecond": 1e3}.get(unit, 1e-3)
    # the loop kernels: drop the OTF-generation / first forward launches by keeping the last 3 * L launches, L per iteration
    loop = [per[i] for i in order if "k_xpassF" not in per[i]["kernel"] and "fwd" not in per[i]["kernel"]]
    n_x = sum(1 for k in loop if "xpass" in k["kernel"])
    if n_x < 6:
        return None, f"unexpected launch list ({len(loop)} loop launches)"
    # iterations end with an X pass; 2 X passes per iteration -> split at every second one, walking backwards
    iters, curk, seen = [], [], 0
    for k in reversed(loop):
        if "xpass" in k["kernel"]:
            if seen == 2:
                iters.append(list(reversed(curk)))
                curk, seen = [], 0
            seen += 1
        curk.append(k)
        if len(iters) == 2:
            break
    mid = iters[1] if len(iters) > 1 else iters[0]
    try:
        shutil.rmtree(tmp)
    except Exception:
        pass
    return {"dram_bytes_per_iteration": sum(k["bytes"] for k in mid),
            "kernels": [{"kernel": k["kernel"][:60], "dram_bytes": k["bytes"], "us_under_ncu": k["us"]} for k in mid]}, \
        "ncu side-run in this bench run: dram__bytes_read.sum + dram__bytes_write.sum of the 2nd of 3 iterations"


def run_config3(args, rank, local_rank, world, barrier):
    """BASELINE config 3: deconDualView joint RL on ONE 1024x1024x512 pair.  N = 1: the single-GPU loop (the strong-scaling
    base).  N > 1: the volume is slab-decomposed over all ranks (csrc/dslab.cu + dist_decon.py), the exchange of the
    distributed FFT folded into the kernels' stores over NVLink peer memory; before timing, a small box is run both ways
    and compared bit for bit with the single-GPU result (rank 0).  Every rank calls this; rank 0 returns the record."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from microimagelib_b200 import device, synth
    shape = tuple(int(v) for v in args.config3_shape.split(","))
    iters = args.config3_iters
    dev = torch.device("cuda", local_rank)
    psf_a = synth.gaussian_psf((65, 65, 65), (4, 2, 2))
    psf_b = synth.gaussian_psf((65, 65, 65), (2, 2, 4))
    nfft = float(np.prod(shape))
    rec = {"config": {"workload": f"deconDualView joint RL {shape[2]}x{shape[1]}x{shape[0]} pair, one volume on {world} GPU(s) (BASELINE config 3)",
                      "iterations": iters, "views": 2}, "unit": "voxel-iters/s", "scaling": "strong", "n_gpus": world}

    def timed(fn):
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    g = torch.Generator(device=dev)
    if world == 1:
        d = device.Decon(shape, 2)
        d.set_psf(0, psf_a)
        d.set_psf(1, psf_b)
        g.manual_seed(20260)
        for v in range(2):
            d.set_image(v, torch.rand(shape, generator=g, device=dev) * 50 + 100)
        d.run(1)
        ms = timed(lambda: d.run(iters)) / iters
        d.close()
        rec.update({"value": nfft / (ms * 1e-3), "ms_per_iteration": ms, "implementation": "single GPU: the same fused kernels, no exchange",
                    "roofline": {"bound": "hbm", "achieved": 2 * ALG_BYTES_PER_VOXEL_ITER * nfft / (ms * 1e-3) / 1e9, "unit": "GB/s",
                                 "peak": peak_hbm()[0], "frac": 2 * ALG_BYTES_PER_VOXEL_ITER * nfft / (ms * 1e-3) / 1e9 / peak_hbm()[0]}})
        torch.cuda.empty_cache()
        return rec
    from microimagelib_b200.dist_decon import DistDecon
    # ---- correctness first: a small box through the distributed path vs the single-GPU path, bit for bit
    small = (64, 128, 128)
    pa, pb = synth.gaussian_psf((17, 17, 17), (3, 2, 2)), synth.gaussian_psf((17, 17, 17), (2, 2, 3))
    g.manual_seed(7)
    va, vb = (torch.rand(small, generator=g, device=dev) * 50 + 100 for _ in range(2))    # same seed on every rank: same volumes
    dd = DistDecon(small, 2)
    L = dd.L
    dd.set_psf(0, pa)
    dd.set_psf(1, pb)
    dd.set_image(0, va[:, L.y0:L.y0 + L.ny, :])
    dd.set_image(1, vb[:, L.y0:L.y0 + L.ny, :])
    E = dd.run(3)
    parts = [torch.empty_like(E) for _ in range(world)]
    dist.all_gather(parts, E)
    same = None
    if rank == 0:
        s = device.Decon(small, 2)
        s.set_psf(0, pa)
        s.set_psf(1, pb)
        s.set_image(0, va)
        s.set_image(1, vb)
        s.run(3)
        ref = torch.empty(small, dtype=torch.float32, device=dev)
        s.result(ref)
        same = bool(torch.equal(torch.cat(parts, dim=1), ref))
        s.close()
    rec["bit_identical_to_single_gpu"] = {"box": list(small), "iterations": 3, "equal": same, "exchange": "fused" if dd.fused else "nccl"}
    dd.close()
    del dd, parts
    torch.cuda.empty_cache()
    # ---- the timed volume
    dd = DistDecon(shape, 2)
    L = dd.L
    dd.set_psf(0, psf_a)
    dd.set_psf(1, psf_b)
    g.manual_seed(20260 + rank)
    for v in range(2):
        dd.set_image(v, torch.rand((L.X, L.ny, L.Z), generator=g, device=dev) * 50 + 100)
    dd.run(1)
    ms = timed(lambda: dd.run(iters)) / iters
    sent = dd.a2a_bytes_per_gpu()                  # bytes one GPU sends in one exchange; 8 exchanges per dual-view iteration
    per_iter = 8 * sent
    rec.update({"value": nfft / (ms * 1e-3), "ms_per_iteration": ms,
                "implementation": ("exchange fused into the X-pass / Y-inverse kernels' stores over NVLink peer memory (CUDA IPC)" if dd.fused
                                   else "NCCL all_to_all_single + re-layout copies"),
                "roofline": {"bound": "nvlink", "achieved": per_iter / (ms * 1e-3) / 1e9, "unit": "GB/s per direction per GPU", "peak": 770.0,
                             "frac": per_iter / (ms * 1e-3) / 1e9 / 770.0, "peak_source": "measured peer-copy figure, B200_PROFILING.md",
                             "bytes_sent_per_gpu_per_iteration": per_iter,
                             "note": "exchange bytes / WHOLE iteration time: the exchange rides on the kernels' stores, so the butterflies of "
                                     "1/N of the volume are inside the same time"}})
    dd.close()
    torch.cuda.empty_cache()
    return rec


def device_loop(d, iters, steps, warmup, stream, barrier, world, local_rank, clocks=None):
    """W warm-up runs, then `steps` timed runs of the full iteration loop (CUDA events on the launching stream)."""
    import torch
    import torch.distributed as dist
    from microimagelib_b200 import device
    for _ in range(warmup):
        d.run(iters, stream=stream)
    barrier()
    if clocks:
        clocks.start()
    l0 = device.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    for _ in range(steps):
        d.run(iters, stream=stream)
    ev1.record(stream)
    barrier()
    launches = device.launch_count() - l0
    ms = ev0.elapsed_time(ev1)
    clk = clocks.stop() if clocks else None
    t = torch.tensor([ms], dtype=torch.float64, device=torch.device("cuda", local_rank))
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()), launches, clk


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
    numa = bind_to_gpu_numa_node(local_rank)   # the end-to-end leg moves 2 x 512 MiB over PCIe per step
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

    # ---- the same loop timed ALONE, before the sustained run heats the board into its power cap: 3 untimed + 20 timed
    # iterations after the set-up's idle time.  Reported beside the headline (roofline.timed_alone), never instead of it.
    alone = None
    if world == 1:
        d.run(3, stream=stream)
        torch.cuda.synchronize()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record(stream)
        d.run(20, stream=stream)
        a1.record(stream)
        torch.cuda.synchronize()
        alone = a0.elapsed_time(a1) / 20.0

    # ---- device-resident loop ------------------------------------------------------------------
    ms_max, launches, clk = device_loop(d, args.iters, args.steps, args.warmup, stream, barrier, world, local_rank, ClockSampler(local_rank))
    value = n_fft * args.iters * args.steps * world / (ms_max * 1e-3)

    # ---- end to end through the reference-facing API (host buffers) -----------------------------
    e2e = None
    h_out = None
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

    fused = bool(d.plane_stage_fused())
    # ---- config 3 (one volume on all ranks): every rank takes part, after the headline's timed regions ------------
    config3 = None
    if not args.no_config3:
        try:
            if world > 1:
                d.close()
            torch.cuda.empty_cache()
            config3 = run_config3(args, rank, local_rank, world, barrier)
        except Exception as e:
            config3 = {"error": str(e)[:300]}
            try:
                barrier()
            except Exception:
                pass

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- everything below: rank 0 only, after the timed regions -----------------------------------
    peak, peak_src = peak_hbm()
    ms_iter = ms_max / (args.steps * args.iters)
    achieved = ALG_BYTES_PER_VOXEL_ITER * n_fft / (ms_iter * 1e-3) / 1e9
    launch_desc = ("one single-view RL iteration = 2 fused plane-stage launches (k_planes_fused) + 2 fused X-pass launches (k_xpassP)" if fused
                   else "one single-view RL iteration = 6 plane-pass launches + 2 fused X-pass launches")
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                "peak_source": peak_src, "frac_of_nominal_8TBps": achieved / 8000.0, "launch": launch_desc,
                "algorithmic_bytes_per_launch": ALG_BYTES_PER_VOXEL_ITER * n_fft, "ms_per_launch": ms_iter}
    if alone:
        roofline["timed_alone"] = {"ms_per_launch": alone, "frac": ALG_BYTES_PER_VOXEL_ITER * n_fft / (alone * 1e-3) / 1e9 / peak,
                                   "what": "20 iterations after 3 untimed ones at process start, CUDA events, before the sustained timed region "
                                           "(which runs under the board's power cap, see clocks)"}
    if world == 1 and not args.no_traffic:
        tr, how = measure_traffic(shape)
        if tr:
            roofline["traffic"] = tr["dram_bytes_per_iteration"]
            roofline["traffic_per_voxel"] = tr["dram_bytes_per_iteration"] / n_fft
            roofline["traffic_kernels"] = tr["kernels"]
        roofline["traffic_source"] = how
    if world == 1:
        try:                                             # per-kernel break-down: kernel-level bytes / measured launch time
            kms = d.time_kernels(5, stream=stream)
            nspec = float(d.fft_shape[0] // 2 + 1) * d.fft_shape[1] * d.fft_shape[2]
            if fused:
                rows = [("k_planes_fused (Y forward, Z forward * OTF, Z inverse, Y inverse; intermediates in L2)", 24 * nspec, kms[0], 2)]
            elif d.row_convolution():
                rows = [("k_ypassF (Y forward, in place)", 16 * nspec, kms[0], 2), ("k_zrow (Z forward * OTF, Z inverse, rows in place)", 24 * nspec, kms[1], 2),
                        ("k_ypassF (Y inverse)", 16 * nspec, kms[2], 2)]
            else:
                rows = [("k_ypassT (Y forward, transposing)", 16 * nspec, kms[0], 2), ("k_zconvT (Z forward * OTF, Z inverse)", 24 * nspec, kms[1], 2),
                        ("k_ypassF (Y inverse)", 16 * nspec, kms[2], 2)]
            rows += [("k_xpassP ratio (C2R, A/T, R2C)", 16 * nspec + 4 * n_fft, kms[3], 1),
                     ("k_xpassP update (C2R, E*T clamp, R2C)", 16 * nspec + 8 * n_fft, kms[4], 1)]
            roofline["kernels"] = [{"kernel": k, "launches_per_iteration": n, "bytes_per_launch": b, "ms_per_launch": float(ms),
                                    "GBps": b / (float(ms) * 1e-3) / 1e9, "frac_of_peak": b / (float(ms) * 1e-3) / 1e9 / peak}
                                   for k, b, ms, n in rows]
            roofline["kernels_note"] = "HBM bytes each launch must read + write (compulsory, L2-resident hand-overs excluded), CUDA events around every launch"
        except Exception as e:
            roofline["kernels"] = {"error": str(e)}

    # ---- CPU baseline (oracle port) and, from the same computation, the in-run parity check ---------
    cpu = parity = None
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        si = 2
        rate, dt, ref_vol = cpu_port_rate(img, psf, si, cores, want_result=True)
        cpu = {"value": rate, "unit": "voxel-iters/s", "cores": cores, "kind": "port",
               "sample": f"{si} RL iterations of the same {shape[2]}x{shape[1]}x{shape[0]} volume ({dt:.1f} s), numpy + scipy.fft "
                         "(pocketfft) float32 restatement of decon_singleview_OTF0; the reference's FFTW path cannot be built here"}
        d.run(si, stream=stream)
        got = d.result()
        got = got.cpu().numpy() if hasattr(got, "cpu") else np.asarray(got)
        err = float(np.linalg.norm(got.astype(np.float64) - ref_vol) / np.linalg.norm(ref_vol.astype(np.float64)))
        parity = {"rel_l2": err, "tolerance": 1e-4, "ok": bool(err <= 1e-4), "against": f"oracle/decon_oracle.py (CPU), same volume, {si} iterations",
                  "what": "final estimate of the CUDA loop vs the CPU restatement of the reference loop"}
        del ref_vol, got

    # ---- the reference's own GPU path on this GPU (oracle/_ref, checker + yardstick; after all timed regions) ----
    refgpu = None
    if world == 1 and not args.no_refgpu:
        try:
            from oracle import ref_gpu
            if ref_gpu.available() and h_out is not None:
                R = ref_gpu.api()
                ours = h_out.copy()
                R.decon_singleview(img, psf, 2, deviceNum=local_rank)           # warm-up (cuFFT plans, context)
                t0 = time.perf_counter()
                rout, rst, rrec = R.decon_singleview(img, psf, args.iters, deviceNum=local_rank)
                rdt = time.perf_counter() - t0
                rerr = float(np.linalg.norm(ours.astype(np.float64) - rout) / np.linalg.norm(rout.astype(np.float64)))
                refgpu = {"what": "the reference's own decon_singleview (cuFFT + its kernels, oracle/_ref/libapi_ref.so) on this GPU, same host "
                                  "buffers, same iteration count; pageable host memory as the reference allocates it",
                          "ms_per_call": rdt * 1e3, "voxel_iters_per_s_e2e": n_fft * args.iters / rdt,
                          "parity_rel_l2_ours_vs_reference": rerr, "iterations": args.iters, "ok": bool(rerr <= 1e-4),
                          "e2e_speedup_ours_vs_reference_gpu": (rdt * 1e3) / e2e["ms_per_step"] if e2e else None}
                if parity is not None:
                    parity["rel_l2_vs_reference_gpu"] = rerr
                del ours, rout
        except Exception as e:
            refgpu = {"error": str(e)[:300]}

    registration = None
    if world == 1:
        try:   # the other kernel of the hot path: fused warp + ZNCC cost at BASELINE config 4's size (8 N bytes per evaluation)
            rshape = (256, 512, 512)
            n_reg = float(np.prod(rshape))
            m = np.array([0.9994, 0.0349, 0, -5.1, -0.0349, 0.9994, 0, 6.3, 0, 0, 1, 1.75], np.float32)   # 2 deg about z + shift
            r = device.Reg(rshape)
            vol = d_img[:rshape[0]].contiguous()
            r.set_images(vol, vol)
            r.prepare()
            registration = {"what": "k_zncc: trilinear warp + ZNCC sums (double accumulation), K candidate matrices per launch, 512x512x256",
                            "algorithmic_bytes_per_evaluation": 8 * n_reg}
            for K in (1, 4, 8):
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
                registration[f"K{K}"] = {"ms_per_evaluation": ms_eval, "GBps": 8 * n_reg / (ms_eval * 1e-3) / 1e9,
                                         "frac_of_peak": 8 * n_reg / (ms_eval * 1e-3) / 1e9 / peak}
            r.close()
            del vol
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

    # ---- second record: BASELINE config 2 (device-resident loop only) ----------------------------
    config2 = None
    if world == 1 and args.shape2 and args.shape2 != args.shape:
        try:
            del d
            torch.cuda.empty_cache()
            shape2 = tuple(int(s) for s in args.shape2.split(","))
            img2, _ = make_inputs(shape2, args.psf, rank)
            d2 = device.Decon(shape2, 1)
            d2.set_psf(0, psf)
            d2.set_image(0, torch.from_numpy(img2).cuda(), stream)
            ms2, _, _ = device_loop(d2, args.iters, args.steps, args.warmup, stream, barrier, world, local_rank)
            n2 = float(np.prod(d2.fft_shape))
            it2 = ms2 / (args.steps * args.iters)
            a2 = ALG_BYTES_PER_VOXEL_ITER * n2 / (it2 * 1e-3) / 1e9
            config2 = {"config": workload_config(shape2, args.iters, args.psf), "value": n2 * args.iters * args.steps / (ms2 * 1e-3),
                       "unit": "voxel-iters/s", "ms_per_step": ms2 / args.steps, "ms_per_iteration": it2,
                       "roofline": {"bound": "hbm", "achieved": a2, "peak": peak, "unit": "GB/s", "frac": a2 / peak, "frac_of_nominal_8TBps": a2 / 8000.0}}
            del d2
        except Exception as e:
            config2 = {"error": str(e)[:300]}

    cfg = workload_config(shape, args.iters, args.psf)
    line = {
        "metric": "RL voxel-iters/sec", "value": value, "unit": "voxel-iters/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": cfg,
        "clocks": clk, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "parity": parity, "cpu_baseline": cpu,
        "reference_gpu_yardstick": refgpu, "yardstick": yard, "registration": registration, "config2": config2, "config3": config3,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
