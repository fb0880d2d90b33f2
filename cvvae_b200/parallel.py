"""Frame-axis sharding of a clip across the GPUs of one NVSwitch node (one process per GPU).

The reference has no multi-GPU inference path; its wrapper walks temporal chunks sequentially
(models/modeling_vae.py:193-210, 279-296).  Chunks are independent work units except for ONE boundary
frame: chunk n re-encodes frame 16n as its causal first frame (:204-206) and chunk n of the decoder
re-decodes latent frame 4n (:291-293).  So a clip sharded on the frame axis needs exactly one halo
exchange per direction of the codec - the last pixel frame (encode) / last latent frame (decode) of
rank r goes to rank r+1 over NVLink (NCCL send/recv; a few MB / a few hundred KB) - and no other
data-path collective: GroupNorm statistics never span a chunk.

Shard layout for world size W and a clip of 1 + 16*n_chunks frames, chunks split contiguously:
  rank r owns chunks [c0, c1):  pixel frames 16*c0+1 .. 16*c1 (rank 0 additionally frame 0)
                                latent frames 4*c0+1 .. 4*c1  (rank 0 additionally latent 0)
Results are bit-identical to the single-GPU engine on the whole clip.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def chunk_ranges(n_chunks: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous split of chunk indices over ranks (first ranks get the remainder)."""
    base, rem = divmod(n_chunks, world)
    out, c = [], 0
    for r in range(world):
        n = base + (1 if r < rem else 0)
        out.append((c, c + n))
        c += n
    return out


def frame_range(c0: int, c1: int, stride: int) -> Tuple[int, int]:
    """[first, last+1) frames owned by the rank holding chunks [c0, c1)."""
    return (0 if c0 == 0 else stride * c0 + 1), stride * c1 + 1


class FrameShardedVAE:
    """encode/decode on a frame-sharded clip with a single halo exchange each."""

    def __init__(self, model, group: Optional[dist.ProcessGroup] = None):
        self.model = model
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        if model.encode_n_frames_a_time is None:
            raise ValueError("frame sharding needs temporal chunking (en_de_n_frames_a_time)")

    # one frame from the left neighbour, one frame to the right neighbour
    def _halo(self, x_local: torch.Tensor, has_left: bool, has_right: bool) -> Optional[torch.Tensor]:
        if self.world == 1:
            return None
        ops = []
        halo = None
        if has_right:
            last = x_local[:, :, -1:].contiguous()
            ops.append(dist.P2POp(dist.isend, last, self._peer(self.rank + 1), self.group))
        if has_left:
            halo = torch.empty_like(x_local[:, :, :1]).contiguous()
            ops.append(dist.P2POp(dist.irecv, halo, self._peer(self.rank - 1), self.group))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        return halo

    def _peer(self, r: int) -> int:
        return dist.get_global_rank(self.group, r) if self.group is not None else r

    def _neighbours(self, n_local_frames: int, stride: int) -> Tuple[bool, bool]:
        # ranks with no chunk (more ranks than chunks) own no frames and take no part
        has_left = self.rank > 0 and n_local_frames > 0
        has_right = self.rank < self.world - 1 and n_local_frames > 0
        return has_left, has_right

    @torch.no_grad()
    def encode_local(self, x_local: torch.Tensor, right_has_frames: bool = True) -> torch.Tensor:
        """x_local: this rank's pixel frames [B,3,T_r,H,W].  Returns its moments [B,2z,T'_r,h,w]."""
        has_left, has_right = self._neighbours(x_local.shape[2], self.model.encode_n_frames_a_time)
        halo = self._halo(x_local, has_left, has_right and right_has_frames)
        x = torch.cat([halo, x_local], dim=2) if halo is not None else x_local
        z = self.model.tiled_encode(x)
        return z[:, :, 1:] if halo is not None else z

    @torch.no_grad()
    def decode_local(self, z_local: torch.Tensor, right_has_frames: bool = True) -> torch.Tensor:
        """z_local: this rank's latent frames.  Returns its pixel frames."""
        has_left, has_right = self._neighbours(z_local.shape[2], self.model.decode_n_frames_a_time)
        halo = self._halo(z_local, has_left, has_right and right_has_frames)
        z = torch.cat([halo, z_local], dim=2) if halo is not None else z_local
        x = self.model.tiled_decode(z)
        return x[:, :, 1:] if halo is not None else x

    def gather_frames(self, t_local: torch.Tensor, lengths: List[int]) -> torch.Tensor:
        """All-gather ragged time shards (dim 2) into the full tensor on every rank."""
        if self.world == 1:
            return t_local
        tmax = max(lengths)
        pad = torch.zeros(t_local.shape[:2] + (tmax,) + t_local.shape[3:], dtype=t_local.dtype, device=t_local.device)
        pad[:, :, : t_local.shape[2]] = t_local
        bufs = [torch.empty_like(pad) for _ in range(self.world)]
        dist.all_gather(bufs, pad, group=self.group)
        return torch.cat([b[:, :, :n] for b, n in zip(bufs, lengths)], dim=2)
