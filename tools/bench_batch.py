"""BASELINE config 5: batch-32 17x3x256x256 latent encode (and encode+decode) throughput, batch-sharded across ranks
(32 / world clips per GPU, no communication on the data path).  Launch with python (1 GPU) or torchrun (N GPUs).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tools/bench_batch.py
"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, ".")
from oracle import cvvae_oracle as O  # noqa: E402  (seeded weights / inputs only)
from cvvae_b200 import CVVAEModel  # noqa: E402

ENC_TFLOP, DEC_TFLOP = 5.664, 17.107   # per 17x256x256 clip, dense reference count


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    total = 32
    per = total // world
    m = CVVAEModel()
    m.load_state_dict(O.make_state_dict(O.VAEConfig(variant="sd21"), 1234))
    m = m.half().cuda()
    x = O.synthetic_video((per, 3, 17, 256, 256), rank).half().cuda()
    out = {}
    for mode in ("encode", "encode_decode"):
        def step():
            z = m.encode(x).latent_dist.mode()
            return m.decode(z).sample if mode == "encode_decode" else z
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        steps = 5
        s.record()
        for _ in range(steps):
            step()
        e.record()
        torch.cuda.synchronize()
        ms = torch.tensor([s.elapsed_time(e) / steps], device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        ms = ms.item()
        tf = total * (ENC_TFLOP + (DEC_TFLOP if mode == "encode_decode" else 0.0))
        out[mode] = {"ms_per_batch32": round(ms, 2), "frames_per_s": round(total * 17 / ms * 1e3, 1),
                     "clips_per_s": round(total / ms * 1e3, 2), "reference_dense_TFLOP_per_s": round(tf / (ms * 1e-3), 1)}
    if rank == 0:
        print(json.dumps({"config": "batch-32 17x3x256x256 fp16, batch-sharded", "n_gpus": world, **out}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
