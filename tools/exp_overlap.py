#!/usr/bin/env python
"""Experiment: the two 576x576 tiles of the 17x576x1024 clip as ONE batch of 2 (product path) vs as two independent
network calls on two CUDA streams, so that the HBM-bound GroupNorm pass of one tile can overlap the tensor-bound
convolutions of the other (needs CVVAE_CONV_SMEM_RESERVE / CVVAE_GN_CTAS_PER_SM for the CTAs to co-reside).

    [CVVAE_CONV_SMEM_RESERVE=8192 CVVAE_GN_CTAS_PER_SM=1] python tools/exp_overlap.py
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    cfg = dict(bench.CONFIGS["c2"])
    m = bench.build_model(cfg, torch.float16)
    g = torch.Generator().manual_seed(0)
    x = (torch.rand((1, 3, 17, 576, 1024), generator=g) * 2 - 1).half().cuda()
    t0, t1 = x[:, :, :, :, 0:576].contiguous(), x[:, :, :, :, 448:1024].contiguous()
    both = torch.cat([t0, t1], dim=0)
    sA, sB = torch.cuda.Stream(), torch.cuda.Stream()

    def batched():
        z = m.encoder(both)
        return m.decoder(z[:, :4].contiguous())

    def two_streams():
        cur = torch.cuda.current_stream()
        outs = []
        for s, t in ((sA, t0), (sB, t1)):
            s.wait_stream(cur)
            with torch.cuda.stream(s):
                z = m.encoder(t)
                outs.append(m.decoder(z[:, :4].contiguous()))
        cur.wait_stream(sA)
        cur.wait_stream(sB)
        return outs

    def timeit(fn, n=5):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(n):
            fn()
        e.record()
        torch.cuda.synchronize()
        return s.elapsed_time(e) / n

    with torch.no_grad():
        rb = batched()
        ra = two_streams()
        torch.cuda.synchronize()
        same = bool(torch.equal(rb[0:1], ra[0]) and torch.equal(rb[1:2], ra[1]))
        res = {"batched_ms": timeit(batched), "two_streams_ms": timeit(two_streams), "bit_identical": same,
               "smem_reserve": os.environ.get("CVVAE_CONV_SMEM_RESERVE"), "gn_ctas_per_sm": os.environ.get("CVVAE_GN_CTAS_PER_SM")}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
