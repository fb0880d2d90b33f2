"""GroupNorm(+SiLU) apply / stats kernels vs HBM peak at the decoder's largest tensors (L2 flushed between reps)."""
import json
import sys

import torch

sys.path.insert(0, ".")
from cvvae_b200.ops import CudaOps  # noqa: E402


def main():
    ops = CudaOps()
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    out = {}
    for name, shape in {"128ch_17x576x576": (1, 17, 576, 576, 128), "256ch_9x288x288": (1, 9, 288, 288, 256),
                        "512ch_5x72x72": (1, 5, 72, 72, 512), "128ch_b2_17x576x576": (2, 17, 576, 576, 128)}.items():
        x = (torch.randn(shape, device="cuda") * 1.5).half()
        C = shape[-1]
        g = torch.rand(C, device="cuda") + 0.5
        b = torch.rand(C, device="cuda") - 0.5
        y = torch.empty_like(x)
        import ctypes as C
        from cvvae_b200 import _lib as L
        from cvvae_b200.ops import _t5, _stream, dtype_code
        stats = torch.empty((shape[0], 32, 2), dtype=torch.int64, device="cuda")
        xs = _t5(x)
        L.check(ops.lib.cvvae_groupnorm_stats(C.byref(xs), 32, 0, stats.data_ptr(), dtype_code(x.dtype), _stream(x)), "stats")
        for mode in ("apply_given_stats", "stats_plus_apply"):
            ts = []
            for _ in range(reps):
                flush.zero_()
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                ops.groupnorm(x, g, b, 32, 1e-5, silu=True, out=y, stats=stats if mode == "apply_given_stats" else None)
                e.record()
                torch.cuda.synchronize()
                ts.append(s.elapsed_time(e))
            ts.sort()
            ms = ts[len(ts) // 2]
            passes = 2 if mode == "apply_given_stats" else 3
            out[f"{name}/{mode}"] = {"ms": round(ms, 4), "GBps": round(x.numel() * 2 * passes / ms / 1e6, 1)}
            print(name, mode, out[f"{name}/{mode}"], flush=True)
    with open("gpurun_out/bench_gn.json", "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
