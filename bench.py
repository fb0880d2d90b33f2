#!/usr/bin/env python
"""bench.py - video VAE encode+decode frames/sec at 17x576x1024 (BASELINE.json metric), one JSON line.

    python bench.py --gpus 1 --steps K --warmup W          # this framework (CUDA engine through the C ABI)
    python bench.py --impl reference ...                   # the reference algorithm on the host CPU cores
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

A "step" = encode(x).latent_dist.mode() -> decode(z).sample of one synthetic fp16 clip through the public
CVVAEModel API, wrapper tiling/chunking on (17x576x1024 = 1 chunk x 2 tiles of 576x576).
  value      frames/s, clip resident in HBM when the timed region starts
  e2e        same, host (pinned) clip -> H2D -> encode/decode -> D2H of the reconstruction, per step
  roofline   tensor bound of the dominant kernel (tcgen05 implicit-GEMM conv): algorithmic FLOPs of its
             launches / their CUDA-event time, against MEASURED_PEAKS.json (sustained bf16 cuBLAS TF/s)
  cpu_baseline  the oracle (CPU restatement of the reference algorithm, fp32) on the host cores, on a
             bounded sample, scaled to the workload by computed pixel count
N>1: the clip grows to 1+16N frames, sharded on the frame axis (one 17-frame chunk per rank, weak scaling),
one NCCL halo exchange per codec direction (cvvae_b200/parallel.py).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "video VAE encode+decode frames/sec at 17x576x1024"
UNIT = "frames/s"
SAMPLE_SHAPE = (1, 3, 17, 192, 192)  # bounded CPU sample (~14 s on the 16 usable cores of a 1-GPU box)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "torch-cuda"])
    ap.add_argument("--frames", type=int, default=17)
    ap.add_argument("--height", type=int, default=576)
    ap.add_argument("--width", type=int, default=1024)
    ap.add_argument("--dtype", default="fp16", choices=["fp16", "bf16"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d.get("bf16_tflops_sustained", 1400.0)), "MEASURED_PEAKS.json bf16_tflops_sustained (of measured)"
    return 1400.0, "fallback 1.4 PFLOP/s sustained (of fallback)"


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                       "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return None
        time.sleep(0.25)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.f.name)
        if not sm:
            return None
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------ CPU baseline
def usable_cores():
    """Host threads this process can really run at once: min(cpu_count, affinity mask, cgroup CPU quota).
    (A 1-GPU box of this pool shows 128 CPUs but a 16-CPU cgroup quota; 128 threads there run 8x SLOWER than 16.)"""
    n = os.cpu_count() or 1
    try:
        n = min(n, len(os.sched_getaffinity(0)))
    except AttributeError:
        pass
    for path in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
        try:
            with open(path) as f:
                parts = f.read().split()
            if path.endswith("cpu.max"):
                if parts[0] != "max":
                    n = min(n, max(1, -(-int(parts[0]) // int(parts[1]))))
            else:
                quota = int(parts[0])
                with open("/sys/fs/cgroup/cpu/cpu.cfs_period_us") as f:
                    period = int(f.read())
                if quota > 0:
                    n = min(n, max(1, -(-quota // period)))
            break
        except (OSError, ValueError, IndexError):
            continue
    return n


def sample_shape(steps):
    """Bounded CPU sample per step, shrunk when many steps are asked for so that the arm ends within a few minutes."""
    side = 192 if steps <= 6 else (128 if steps <= 24 else 96)
    return (1, 3, 17, side, side)


def cpu_sample(steps=1, shape=None):
    """Reference algorithm (oracle port, fp32) on the host cores on a bounded sample."""
    from oracle import cvvae_oracle as O  # the one place bench.py executes oracle/: the CPU baseline
    cores = usable_cores()
    torch.set_num_threads(cores)
    wrap = dict(tile_spatial_size=None, en_de_n_frames_a_time=None)
    cfg = O.VAEConfig(variant="sd21", **wrap)
    sd = O.make_state_dict(cfg, 1234)
    x = O.synthetic_video(shape or SAMPLE_SHAPE, 0)
    with torch.no_grad():
        O.decode(O.encode(O.synthetic_video((1, 3, 1, 32, 32), 0), sd, cfg).mode(), sd, cfg)  # page in oneDNN
        t0 = time.perf_counter()
        for _ in range(steps):
            O.decode(O.encode(x, sd, cfg).mode(), sd, cfg)
        dt = (time.perf_counter() - t0) / steps
    return dt, cores


def workload_pixels(frames, height, width):
    """Pixels the wrapper actually pushes through the networks (tile overlap included)."""
    def tiles(n, tile=576, stride=448):
        out, i = [], 0
        while True:
            out.append(min(tile, n - i))
            if i + tile >= n:
                break
            i += stride
        return out
    return frames * sum(tiles(height)) * sum(tiles(width))


def cpu_baseline_entry(args, steps=1):
    shp = sample_shape(steps)
    dt, cores = cpu_sample(steps, shp)
    ratio = workload_pixels(args.frames, args.height, args.width) / (shp[2] * shp[3] * shp[4])
    fps = args.frames / (dt * ratio)
    return {"value": fps, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"oracle fp32 encode+decode of one {shp[2]}x{shp[3]}x{shp[4]} clip "
                      f"({dt:.2f} s), scaled x{ratio:.1f} by network-input pixel count to "
                      f"{args.frames}x{args.height}x{args.width} (tile overlap of the 576/448 tiling included)"}, dt


def run_reference(args, rank):
    if rank != 0:
        return
    steps = max(1, args.steps)
    for _ in range(min(args.warmup, 1)):
        cpu_sample(1)
    entry, dt = cpu_baseline_entry(args, steps)
    line = {"impl": "reference", "metric": METRIC, "value": entry["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": steps, "warmup": min(args.warmup, 1), "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.frames}x3x{args.height}x{args.width} encode+decode, bounded CPU sample per step"},
            "cpu_baseline": entry,
            "e2e": {"value": entry["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------ GPU arms
def build_model(dtype):
    from cvvae_b200 import CVVAEModel
    torch.manual_seed(1234)
    m = CVVAEModel()  # reference defaults: 4-ch latent, tile 576, chunks of 16(+1) frames
    g = torch.Generator().manual_seed(4321)
    for k, p in m.named_parameters():  # non-trivial affine/bias so nothing is skipped or degenerate
        if p.dim() == 1:
            p.data.copy_(torch.rand(p.shape, generator=g) * (0.4 if k.endswith("bias") else 1.0) + (-0.2 if k.endswith("bias") else 0.5))
    return m.to(dtype).cuda()


class TorchCudaReference:
    """The reference algorithm on torch-CUDA library kernels (cuDNN etc.) - informational arm only
    (`--impl torch-cuda`): the north_star's '>= 4x the reference's own torch-cuda' denominator."""

    def __init__(self, dtype):
        from oracle import cvvae_oracle as O
        self.O = O
        self.cfg = O.VAEConfig(variant="sd21")
        m = build_model(dtype)
        self.sd = {k: v for k, v in m.state_dict().items()}

    def tiled_encode(self, x):
        return self.O.tiled_encode(x, self.sd, self.cfg)

    def tiled_decode(self, z):
        return self.O.tiled_decode(z, self.sd, self.cfg)

    encode_n_frames_a_time = 16
    decode_n_frames_a_time = 4


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    assert torch.cuda.is_available(), "bench.py needs a GPU (use --impl reference for the CPU arm)"
    torch.cuda.set_device(local)
    import torch.distributed as dist
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dtype = torch.float16 if args.dtype == "fp16" else torch.bfloat16
    zc = 4

    if args.impl == "ours":
        import __graft_entry__ as ge
        if rank == 0:
            ge.build()
        if world > 1:
            dist.barrier()
        model = build_model(dtype)
        ops = model._engine().ops
    else:
        model = TorchCudaReference(dtype)
        ops = None

    from cvvae_b200.parallel import FrameShardedVAE
    sharded = FrameShardedVAE(model) if world > 1 else None

    # this rank's shard of the clip: rank 0 holds frames 0..16, rank r frames 16r+1..16r+16
    n_local = args.frames if rank == 0 else args.frames - 1
    g = torch.Generator().manual_seed(rank)
    x_host = (torch.rand((1, 3, n_local, args.height, args.width), generator=g) * 2 - 1).to(dtype).pin_memory()
    x_dev = x_host.cuda()

    def step(x):
        with torch.no_grad():
            if sharded is not None:
                z = sharded.encode_local(x)[:, :zc].contiguous()
                return sharded.decode_local(z)
            z = model.tiled_encode(x)[:, :zc]
            return model.tiled_decode(z)

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps):
        sync()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(steps):
            fn()
        e.record()
        sync()
        ms = torch.tensor([s.elapsed_time(e)], device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    for _ in range(max(args.warmup, 3)):
        out = step(x_dev)
    sync()

    # ---- device-resident timing (+ per-conv CUDA events for the roofline)
    launches0 = ops.launch_count() if ops else 0
    sampler = ClockSampler(local) if rank == 0 else None
    if ops:
        ops.start_profile()
    ms_total = timed(lambda: step(x_dev), args.steps)
    prof = ops.stop_profile() if ops else None
    clocks = sampler.stop() if sampler else None
    launches = (ops.launch_count() - launches0) if ops else 0
    ms_step = ms_total / args.steps
    total_frames = args.frames + (world - 1) * (args.frames - 1)
    value = total_frames / (ms_step * 1e-3)

    # ---- end-to-end: pinned host clip -> H2D -> encode/decode -> D2H of the reconstruction, every step
    rec_host = torch.empty(out.shape, dtype=out.dtype).pin_memory()

    def e2e_step():
        xd = x_host.cuda(non_blocking=True)
        r = step(xd)
        rec_host.copy_(r, non_blocking=True)

    if args.impl == "ours":
        model.enable_cuda_graphs(True)  # public option of the model: every network call replays a captured graph
    e2e_step()
    e2e_step()
    ms_e2e = timed(e2e_step, args.steps) / args.steps
    e2e = {"value": total_frames / (ms_e2e * 1e-3), "unit": UNIT, "cuda_graphs": args.impl == "ours",
           "h2d_bytes_per_step": x_host.numel() * x_host.element_size(),
           "d2h_bytes_per_step": rec_host.numel() * rec_host.element_size()}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = peaks()
    roof = None
    if prof and prof["conv_tc"]["ms"] > 0:
        tc = prof["conv_tc"]
        achieved = tc["flops"] / (tc["ms"] * 1e-3) / 1e12
        traffic, traffic_src = None, None
        tp = os.path.join(ROOT, "profiles", "r01_traffic.json")
        if os.path.exists(tp):  # ncu dram bytes (read + write) per conv_tc launch, from the committed capture of this command
            with open(tp) as f:
                tj = json.load(f)
            traffic, traffic_src = tj["dram_bytes_per_launch"], tj["source"]
        roof = {"bound": "tensor", "kernel": "conv_tc_kernel (tcgen05 implicit-GEMM conv, all launches of the timed region)",
                "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic,
                "traffic_source": traffic_src,
                "algorithmic_bytes_per_launch": tc["bytes"] / max(tc["launches"], 1),
                "peak_source": peak_src, "algorithmic_tflop_per_step": tc["flops"] / args.steps / 1e12,
                "kernel_ms_per_step": tc["ms"] / args.steps, "launches_per_step": tc["launches"] / args.steps,
                "share_of_step": tc["ms"] / ms_total,
                # the up-sampling convs run as folded 2x2 phase kernels (2.25x fewer MACs than the reference issues);
                # `achieved` counts EXECUTED FLOPs, this is the same time against the reference's dense count
                "reference_dense_tflop_per_step": tc["ref_flops"] / args.steps / 1e12,
                "achieved_vs_reference_count": tc["ref_flops"] / (tc["ms"] * 1e-3) / 1e12,
                "conv_direct_ms_per_step": prof["conv_direct"]["ms"] / args.steps}
    def _n(n, tile=576, stride=448):
        k, i = 1, 0
        while i + tile < n:
            i += stride
            k += 1
        return k
    n_tiles = _n(args.height) * _n(args.width)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": f"{total_frames}x3x{args.height}x{args.width} clip, encode(x).mode() -> decode(z), "
                                   f"wrapper tiling 576/448 + 16-frame chunks; SD2.1-variant CVVAEModel, seeded random weights",
                       "per_gpu": f"one {args.frames}-frame chunk = {n_tiles} spatial tiles (<= 576x576, stride 448)", "parallelism": f"frame-shard x{world}",
                       "l2": "no explicit flush: every step streams >100 GB of activations (each up to 1.4 GB) through a 126 MB L2"},
            "impl": args.impl, "gpu_launches": launches // args.steps if launches else 0, "clocks": clocks, "e2e": e2e}
    line["peak_hbm_gb"] = round(torch.cuda.max_memory_allocated() / 1e9, 2)  # activations + weights + graph pools, this rank
    if roof:
        line["roofline"] = roof
    if args.impl == "ours" and not args.no_cpu_baseline and world == 1:
        line["cpu_baseline"] = cpu_baseline_entry(args, 1)[0]
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
