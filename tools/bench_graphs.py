"""Eager launches vs CUDA-graph replay of the network calls, at the sizes where the host is the bottleneck."""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
from oracle import cvvae_oracle as O  # noqa: E402  (weights only)
from cvvae_b200 import CVVAEModel  # noqa: E402


def timeit(fn, n=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    t0 = time.perf_counter()
    for a, b in ev:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / n * 1e3
    ms = sorted(a.elapsed_time(b) for a, b in ev)
    return ms[len(ms) // 2], wall


def main():
    cfg = O.VAEConfig(variant="sd21")
    m = CVVAEModel()
    m.load_state_dict(O.make_state_dict(cfg, 1234))
    m = m.half().cuda()
    out = {}
    shapes = {"clip_17x256x256": (1, 3, 17, 256, 256), "clip_b4_17x256x256": (4, 3, 17, 256, 256),
              "image_1x512x512": (1, 3, 1, 512, 512), "clip_17x576x1024": (1, 3, 17, 576, 1024)}
    for name, shp in shapes.items():
        x = O.synthetic_video(shp, 0).half().cuda()

        def run():
            z = m.encode(x).latent_dist.mode()
            return m.decode(z).sample

        row = {}
        for mode in ("eager", "graphs"):
            m.enable_cuda_graphs(mode == "graphs")
            n = 3 if shp[-1] >= 1024 else 8
            dev, wall = timeit(run, n=n)
            row[mode] = {"device_ms": round(dev, 3), "wall_ms": round(wall, 3), "fps": round(shp[0] * shp[2] / dev * 1e3, 2)}
        m.enable_cuda_graphs(False)
        out[name] = row
        print(name, json.dumps(row), flush=True)
    # SD-pipeline call: 4-D latents, decode(z / scaling_factor, num_frames=1)  (pipeline_stable_diffusion.py:1046)
    for name, shp in {"sd_pipeline_decode_1x64x64_latent": (1, 4, 64, 64), "sd_pipeline_decode_4x64x64_latent": (4, 4, 64, 64),
                      "sd_pipeline_decode_1x128x128_latent": (1, 4, 128, 128)}.items():
        z = (torch.randn(shp, generator=torch.Generator().manual_seed(3)) * 0.5).half().cuda()
        row = {}
        for mode in ("eager", "graphs"):
            m.enable_cuda_graphs(mode == "graphs")
            dev, wall = timeit(lambda: m.decode(z, num_frames=1).sample, n=10)
            row[mode] = {"device_ms": round(dev, 3), "wall_ms": round(wall, 3), "images_per_s": round(shp[0] / dev * 1e3, 1)}
        m.enable_cuda_graphs(False)
        out[name] = row
        print(name, json.dumps(row), flush=True)
    with open("gpurun_out/bench_graphs.json", "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
