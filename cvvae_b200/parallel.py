"""Sharding of a clip across the GPUs of one NVSwitch node (one process per GPU).

The reference has no multi-GPU inference path; its wrapper walks temporal chunks and spatial tiles
sequentially (models/modeling_vae.py:144-210, 230-296).  Every (chunk, tile) is an independent
encoder / decoder call with its own GroupNorm statistics, coupled to the rest only by ONE boundary
frame per chunk - chunk n re-encodes frame 16n as its causal first frame (:204-206), chunk n of the
decoder re-decodes latent frame 4n (:291-293) - and by the cheap blend of adjacent tile outputs.

Two schemes, both bit-identical to the single-GPU engine on the whole clip:

``FrameShardedVAE``  the clip is SHARDED on the frame axis (rank r holds the frames of its contiguous
    chunks).  One halo exchange per direction of the codec: the last pixel frame (encode) / last latent
    frame (decode) of rank r goes to rank r+1 over NVLink (NCCL send/recv, a few MB / a few hundred KB);
    no other data-path collective - GroupNorm never spans a chunk.  The exchange is non-blocking: each
    neighbour pair has its own 2-rank communicator (so a rank's send never queues behind its receive),
    transfers are posted from a side stream the moment the boundary frame exists, and the compute stream
    only waits for the halo it is about to consume.  A fast rank is never stalled by a slow right
    neighbour; a rank waits for its left neighbour only if that one is late producing the frame.

``UnitShardedVAE``   the clip is RESIDENT on every rank (e.g. every rank decoded the same file); the
    (chunk x tile) work units are dealt round-robin, each rank runs the networks on its units and the
    unit results are broadcast from their owners, after which every rank blends/assembles locally.
    Scales clips with fewer chunks than GPUs: the 17-frame 576x1024 clip (1 chunk x 2 tiles) uses 2 GPUs,
    a 33-frame 720p clip (2 x 6 units) uses all 8.

Shard layout of FrameShardedVAE for world size W and a clip of 1 + 16*n_chunks frames, chunks split contiguously:
  rank r owns chunks [c0, c1):  pixel frames 16*c0+1 .. 16*c1 (rank 0 additionally frame 0)
                                latent frames 4*c0+1 .. 4*c1  (rank 0 additionally latent 0)
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

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
    """encode/decode on a frame-sharded clip with a single, non-blocking halo exchange each."""

    def __init__(self, model, group: Optional[dist.ProcessGroup] = None):
        self.model = model
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        if model.encode_n_frames_a_time is None:
            raise ValueError("frame sharding needs temporal chunking (en_de_n_frames_a_time)")
        # one 2-rank communicator per neighbour pair (created collectively, in the same order on every rank)
        self._pair: Dict[int, dist.ProcessGroup] = {}
        if self.world > 1:
            for r in range(self.world - 1):
                g = dist.new_group(ranks=[self._global(r), self._global(r + 1)])
                if self.rank in (r, r + 1):
                    self._pair[r] = g
        self._side: Optional[torch.cuda.Stream] = None
        self._inflight: list = []   # (work, tensor) of sends still possibly in flight: keeps the buffers alive

    def _global(self, r: int) -> int:
        return dist.get_global_rank(self.group, r) if self.group is not None else r

    # ---- halo: one frame from the left neighbour, one frame to the right neighbour ------------------------
    def _halo(self, x_local: torch.Tensor, has_left: bool, has_right: bool):
        """Post the exchange; returns (halo tensor or None, event the consumer stream must wait for or None)."""
        if self.world == 1 or not (has_left or has_right):
            return None, None
        cuda = x_local.is_cuda
        halo, ev = None, None
        if not cuda:   # gloo (CPU tests): plain non-blocking ops, completed before returning
            reqs = []
            if has_left:
                halo = torch.empty_like(x_local[:, :, :1]).contiguous()
                reqs.append(dist.irecv(halo, src=self._global(self.rank - 1), group=self._pair[self.rank - 1]))
            if has_right:
                reqs.append(dist.isend(x_local[:, :, -1:].contiguous(), dst=self._global(self.rank + 1), group=self._pair[self.rank]))
            for r in reqs:
                r.wait()
            return halo, None
        cur = torch.cuda.current_stream(x_local.device)
        if self._side is None:
            self._side = torch.cuda.Stream(device=x_local.device)
        side = self._side
        side.wait_stream(cur)          # the boundary frame is produced on the compute stream
        with torch.cuda.stream(side):
            for w, _ in self._inflight:  # sends of the previous exchange: long done; orders buffer reuse on the side stream
                w.wait()
            self._inflight = []
            recv_w = None
            if has_left:
                halo = torch.empty_like(x_local[:, :, :1]).contiguous()
                recv_w = dist.irecv(halo, src=self._global(self.rank - 1), group=self._pair[self.rank - 1])
            if has_right:
                last = x_local[:, :, -1:].contiguous()
                self._inflight.append((dist.isend(last, dst=self._global(self.rank + 1), group=self._pair[self.rank]), last))
            if recv_w is not None:
                recv_w.wait()           # side stream waits for the frame; the host does not block
                ev = torch.cuda.Event()
                ev.record(side)
        if halo is not None:
            halo.record_stream(cur)
        return halo, ev

    def _neighbours(self, n_local_frames: int, total_chunks: Optional[int]) -> Tuple[bool, bool]:
        """Participation derived from the GLOBAL chunk count: with more ranks than chunks the tail ranks own no frames
        and take no part, and the last owning rank has nobody to send to.  total_chunks=None: every rank owns >= 1 chunk."""
        if n_local_frames == 0:
            return False, False
        if total_chunks is None:
            return self.rank > 0, self.rank < self.world - 1
        ranges = chunk_ranges(total_chunks, self.world)
        owns = [c1 > c0 for c0, c1 in ranges]
        return (self.rank > 0 and owns[self.rank - 1]), (self.rank < self.world - 1 and owns[self.rank + 1])

    def _with_halo(self, t_local: torch.Tensor, total_chunks: Optional[int]):
        has_left, has_right = self._neighbours(t_local.shape[2], total_chunks)
        halo, ev = self._halo(t_local, has_left, has_right)
        if halo is None:
            return t_local, False
        # [halo | local frames] in one pre-sized buffer: the local frames are copied while the halo is still in flight
        full = torch.empty(t_local.shape[:2] + (t_local.shape[2] + 1,) + t_local.shape[3:], dtype=t_local.dtype, device=t_local.device)
        full[:, :, 1:].copy_(t_local)
        if ev is not None:
            torch.cuda.current_stream(t_local.device).wait_event(ev)
        full[:, :, :1].copy_(halo)
        return full, True

    @torch.no_grad()
    def encode_local(self, x_local: torch.Tensor, total_chunks: Optional[int] = None) -> torch.Tensor:
        """x_local: this rank's pixel frames [B,3,T_r,H,W].  Returns its moments [B,2z,T'_r,h,w]."""
        if x_local.shape[2] == 0:
            eng = self.model._engine()
            return x_local.new_empty((x_local.shape[0], eng.cfg.moments_channels, 0, eng.encoded_hw(x_local.shape[3]),
                                      eng.encoded_hw(x_local.shape[4])))
        x, had = self._with_halo(x_local, total_chunks)
        z = self.model.tiled_encode(x)
        return z[:, :, 1:] if had else z

    @torch.no_grad()
    def decode_local(self, z_local: torch.Tensor, total_chunks: Optional[int] = None) -> torch.Tensor:
        """z_local: this rank's latent frames.  Returns its pixel frames."""
        if z_local.shape[2] == 0:
            eng = self.model._engine()
            return z_local.new_empty((z_local.shape[0], eng.cfg.out_ch, 0, eng.decoded_hw(z_local.shape[3]),
                                      eng.decoded_hw(z_local.shape[4])))
        z, had = self._with_halo(z_local, total_chunks)
        x = self.model.tiled_decode(z)
        return x[:, :, 1:] if had else x

    def gather_frames(self, t_local: torch.Tensor, lengths: List[int]) -> torch.Tensor:
        """All-gather ragged time shards (dim 2) into the full tensor on every rank (ranks without frames pass the
        empty tensor their encode_local / decode_local returned)."""
        if self.world == 1:
            return t_local
        tmax = max(lengths)
        shp = t_local.shape
        pad = torch.zeros(shp[:2] + (tmax,) + shp[3:], dtype=t_local.dtype, device=t_local.device)
        if shp[2]:
            pad[:, :, : shp[2]] = t_local
        bufs = [torch.empty_like(pad) for _ in range(self.world)]
        dist.all_gather(bufs, pad, group=self.group)
        return torch.cat([b[:, :, :n] for b, n in zip(bufs, lengths)], dim=2)


class UnitShardedVAE:
    """(chunk x tile) work units of a clip resident on every rank, dealt round-robin; results broadcast by their owners.

    ``encode(x)`` / ``decode(z)`` take and return FULL tensors on every rank (same values as the single-GPU
    ``model.tiled_encode`` / ``tiled_decode``).  The only data-path communication is one broadcast per work unit of its
    (small: moments; or pixel-tile sized) result."""

    def __init__(self, model, group: Optional[dist.ProcessGroup] = None):
        self.model = model
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self._base = 0
        self._which = "encode"

    def _global(self, r: int) -> int:
        return dist.get_global_rank(self.group, r) if self.group is not None else r

    def _out_shape(self, in_shape) -> Tuple[int, ...]:
        eng = self.model._engine()
        B, _, T, H, W = in_shape
        if self._which == "encode":
            return (B, eng.cfg.moments_channels, eng.encoded_frames(T), eng.encoded_hw(H), eng.encoded_hw(W))
        return (B, eng.cfg.out_ch, eng.decoded_frames(T), eng.decoded_hw(H), eng.decoded_hw(W))

    def _runner(self, flat, fn):
        """Replacement of the wrapper's local tile loop for one temporal chunk: flat = [(row, col, tile view)]."""
        mine = [k for k in range(len(flat)) if (self._base + k) % self.world == self.rank]
        local = self.model._run_tiles_local([flat[k] for k in mine], fn) if mine else {}
        results = {}
        for k, (r, c, v) in enumerate(flat):
            owner = (self._base + k) % self.world
            if owner == self.rank:
                buf = local[(r, c)].contiguous()
            else:
                buf = torch.empty(self._out_shape(v.shape), dtype=v.dtype, device=v.device)
            if self.world > 1:
                dist.broadcast(buf, src=self._global(owner), group=self.group)
            results[(r, c)] = buf
        self._base += len(flat)
        return results

    def _run(self, which: str, t: torch.Tensor) -> torch.Tensor:
        m = self.model
        self._base, self._which = 0, which
        prev = m._tile_runner
        m._tile_runner = self._runner
        try:
            return m.tiled_encode(t) if which == "encode" else m.tiled_decode(t)
        finally:
            m._tile_runner = prev

    @torch.no_grad()
    def encode(self, x: torch.Tensor) -> torch.Tensor:
        return self._run("encode", x)

    @torch.no_grad()
    def decode(self, z: torch.Tensor) -> torch.Tensor:
        return self._run("decode", z)
