"""2-rank NCCL check of the frame-axis sharding on real GPUs: sharded encode/decode (one halo send/recv per direction,
ragged gather) is bit-identical to the single-GPU run of the same engine.  Skipped on a 1-GPU box."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from cvvae_b200 import CVVAEModel
        from cvvae_b200.parallel import FrameShardedVAE, chunk_ranges, frame_range
        from oracle import cvvae_oracle as O
        wrap = dict(tile_spatial_size=72, en_de_n_frames_a_time=4)
        m = CVVAEModel(ch=32, **wrap)
        m.load_state_dict(O.make_state_dict(O.VAEConfig(variant="sd21", ch=32, **wrap), 1234))
        m = m.half().cuda()
        x = O.synthetic_video((1, 3, 13, 80, 96), 3).half().cuda()  # 3 chunks of 4 frames -> ranks get 2 + 1
        full_z = m.encode(x).latent_dist.parameters
        full_x = m.decode(full_z[:, :4].contiguous()).sample
        sh = FrameShardedVAE(m)
        ranges = chunk_ranges(3, world)
        c0, c1 = ranges[rank]
        f0, f1 = frame_range(c0, c1, 4)
        z_local = sh.encode_local(x[:, :, f0:f1].contiguous())
        l0, l1 = frame_range(c0, c1, 1)
        assert torch.equal(z_local, full_z[:, :, l0:l1]), "sharded encode differs from the single-GPU encode"
        x_local = sh.decode_local(z_local[:, :4].contiguous())
        assert torch.equal(x_local, full_x[:, :, f0:f1]), "sharded decode differs from the single-GPU decode"
        lens = [frame_range(a, b, 4)[1] - frame_range(a, b, 4)[0] for a, b in ranges]
        gathered = sh.gather_frames(x_local, lens)
        assert torch.equal(gathered, full_x)
        for _ in range(3):   # repeated exchanges reuse the side stream / pair communicators
            assert torch.equal(sh.encode_local(x[:, :, f0:f1].contiguous()), full_z[:, :, l0:l1])
        from cvvae_b200.parallel import UnitShardedVAE
        us = UnitShardedVAE(m)
        assert torch.equal(us.encode(x), full_z), "unit-sharded encode differs from the single-GPU encode"
        assert torch.equal(us.decode(full_z[:, :4].contiguous()), full_x), "unit-sharded decode differs"
        torch.cuda.synchronize()
        ret[rank] = "ok"
    except Exception:  # pragma: no cover
        import traceback
        ret[rank] = traceback.format_exc()
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_frame_sharding_two_gpus_nccl():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    port = 29600 + os.getpid() % 2000
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret.get(0) == "ok" and ret.get(1) == "ok", dict(ret)
