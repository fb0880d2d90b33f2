"""world_size-2 gloo test of the frame-axis sharding + halo exchange (host logic; CPU, fake operator set)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from cvvae_b200 import CVVAEModel
        from cvvae_b200.parallel import FrameShardedVAE, chunk_ranges, frame_range
        from fake_ops import FakeOps
        from oracle import cvvae_oracle as O
        torch.set_num_threads(2)
        wrap = dict(tile_spatial_size=72, en_de_n_frames_a_time=4)
        m = CVVAEModel(ch=32, **wrap)
        m.load_state_dict(O.make_state_dict(O.VAEConfig(variant="sd21", ch=32, **wrap), 1234))
        m._ops_factory = FakeOps
        x = O.synthetic_video((1, 3, 13, 80, 96), 3)  # 3 chunks of 4 -> ranks get 2 + 1
        full_z = m.encode(x).latent_dist.parameters
        full_x = m.decode(full_z[:, :4]).sample
        sh = FrameShardedVAE(m)
        ranges = chunk_ranges(3, world)
        c0, c1 = ranges[rank]
        f0, f1 = frame_range(c0, c1, 4)
        z_local = sh.encode_local(x[:, :, f0:f1].contiguous())
        l0, l1 = frame_range(c0, c1, 1)
        assert torch.equal(z_local, full_z[:, :, l0:l1]), "sharded encode differs from single-process encode"
        x_local = sh.decode_local(z_local[:, :4].contiguous())
        assert torch.equal(x_local, full_x[:, :, f0:f1]), "sharded decode differs from single-process decode"
        lens = [frame_range(a, b, 4)[1] - frame_range(a, b, 4)[0] for a, b in ranges]
        gathered = sh.gather_frames(x_local, lens)
        assert torch.equal(gathered, full_x)
        # more ranks than chunks: the tail rank owns no frames, takes no part in the halo and nobody waits for it
        x1 = x[:, :, :5].contiguous()                      # one chunk of 4(+1) frames
        full_z1 = m.encode(x1).latent_dist.parameters
        r1 = chunk_ranges(1, world)
        a, b = frame_range(*r1[rank], 4) if r1[rank][1] > r1[rank][0] else (0, 0)
        z1 = sh.encode_local(x1[:, :, a:b].contiguous(), total_chunks=1)
        if rank == 0:
            assert torch.equal(z1, full_z1)
        else:
            assert z1.shape[2] == 0 and z1.shape[1] == full_z1.shape[1] and z1.shape[3:] == full_z1.shape[3:]
        assert torch.equal(sh.gather_frames(z1, [full_z1.shape[2], 0]), full_z1)
        # work-unit (chunk x tile) sharding of a clip resident on every rank
        from cvvae_b200.parallel import UnitShardedVAE
        us = UnitShardedVAE(m)
        assert torch.equal(us.encode(x), full_z), "unit-sharded encode differs from single-process encode"
        assert torch.equal(us.decode(full_z[:, :4].contiguous()), full_x), "unit-sharded decode differs"
        x17 = x[:, :, :5]                                  # one chunk x 4 tiles: more units than ranks inside one chunk
        assert torch.equal(us.encode(x17), m.encode(x17).latent_dist.parameters)
        ret[rank] = "ok"
    except Exception as e:  # pragma: no cover
        import traceback
        ret[rank] = traceback.format_exc()
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_frame_sharding_two_ranks():
    from cvvae_b200.parallel import chunk_ranges, frame_range
    assert chunk_ranges(8, 8) == [(i, i + 1) for i in range(8)]
    assert chunk_ranges(3, 2) == [(0, 2), (2, 3)]
    assert frame_range(0, 1, 16) == (0, 17) and frame_range(1, 2, 16) == (17, 33)
    port = 29500 + os.getpid() % 2000
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret.get(0) == "ok" and ret.get(1) == "ok", dict(ret)
