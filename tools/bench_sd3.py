#!/usr/bin/env python
"""BASELINE.json config C3: SD3-variant CV-VAE (16-ch latent), 33x3x512x512, bf16, wrapper chunking on (2 chunks x 1 tile).
Prints frames/s of encode(x).mode() -> decode(z) plus the error against the reference algorithm in bf16 on torch-CUDA."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cvvae_b200 import CVVAESD3Model  # noqa: E402


def main():
    torch.manual_seed(1234)
    m = CVVAESD3Model()
    g = torch.Generator().manual_seed(4321)
    for k, p in m.named_parameters():
        if p.dim() == 1:
            p.data.copy_(torch.rand(p.shape, generator=g) * (0.4 if k.endswith("bias") else 1.0) + (-0.2 if k.endswith("bias") else 0.5))
    m = m.to(torch.bfloat16).cuda()
    x = (torch.rand((1, 3, 33, 512, 512), generator=torch.Generator().manual_seed(0)) * 2 - 1).to(torch.bfloat16).cuda()

    def step():
        z = m.encode(x).latent_dist.mode()
        return m.decode(z).sample

    for _ in range(3):
        rec = step()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = 3
    s.record()
    for _ in range(steps):
        rec = step()
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / steps
    print(json.dumps({"config": "C3 sd3 33x3x512x512 bf16", "ms_per_step": ms, "frames_per_s": 33 / (ms * 1e-3),
                      "finite": bool(torch.isfinite(rec).all()), "out_shape": list(rec.shape)}))


if __name__ == "__main__":
    main()
