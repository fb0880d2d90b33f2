"""How fast does the CPU baseline sample run on this host for a given intra-op thread count?  (picks bench.py's setting)"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, ".")
from oracle import cvvae_oracle as O  # noqa: E402

n = int(sys.argv[1])
torch.set_num_threads(n)
cfg = O.VAEConfig(variant="sd21", tile_spatial_size=None, en_de_n_frames_a_time=None)
sd = O.make_state_dict(cfg, 1234)
x = O.synthetic_video((1, 3, 17, 128, 128), 0)
with torch.no_grad():
    O.decode(O.encode(O.synthetic_video((1, 3, 1, 32, 32), 0), sd, cfg).mode(), sd, cfg)
    t0 = time.perf_counter()
    O.decode(O.encode(x, sd, cfg).mode(), sd, cfg)
    dt = time.perf_counter() - t0
print(json.dumps({"threads": n, "omp_env": os.environ.get("OMP_NUM_THREADS"), "cpu_count": os.cpu_count(),
                  "affinity": len(os.sched_getaffinity(0)), "seconds": round(dt, 2)}), flush=True)
