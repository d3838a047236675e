"""Runner layer (no counterpart in the reference, which only exposes ``model(x)``): how frames and
large images are fed through an engine model on one or several B200s.

* ``FramePipeline``  — host-to-host streaming of frames on one GPU: pinned H2D copy, forward and D2H
  copy of consecutive frames overlap on three CUDA streams.
* ``tiled_forward``  — exact halo tiling of one large image: every tile is extended by the model's
  receptive radius, so the stitched result is bit-identical to the untiled forward (zero padding only
  ever applies at true image borders).
* ``shard_indices`` / ``gather_to_rank`` — one process per GPU; frames (or tiles) are dealt round-robin
  to ranks with no collective on the compute path; a gather (NCCL over NVLink on GPUs, gloo in the CPU
  tests) is used only when one rank wants every output.

Everything takes a plain callable ``model(x) -> y`` plus its integer ``upscale``, so the logic is
testable without a GPU.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import torch

Model = Callable[[torch.Tensor], torch.Tensor]


# ---------------------------------------------------------------------------------------- tiling
def plan_tiles(h: int, w: int, tile_h: int, tile_w: int, halo: int, multiple: int = 1) -> List[Tuple[int, int, int, int, int, int, int, int]]:
    """Tile grid over an h x w image.  Each entry is
    (y0, y1, x0, x1, ey0, ey1, ex0, ex1): the core region [y0,y1) x [x0,x1) a tile is responsible for and
    the halo-extended region [ey0,ey1) x [ex0,ex1) it is computed from (clipped at the image border).
    ``multiple`` > 1 (models that re-grid their input, e.g. a pixel-unshuffle front end): tile sizes must be multiples of it
    and the halo is rounded up to one, so every extended region starts on the model's grid."""
    if tile_h <= 0 or tile_w <= 0 or halo < 0:
        raise ValueError('tile sizes must be positive and halo non-negative')
    if multiple > 1:
        if tile_h % multiple or tile_w % multiple:
            raise ValueError(f'tile sizes must be multiples of {multiple} for this model')
        halo = (halo + multiple - 1) // multiple * multiple
    tiles = []
    for y0 in range(0, h, tile_h):
        y1 = min(y0 + tile_h, h)
        for x0 in range(0, w, tile_w):
            x1 = min(x0 + tile_w, w)
            tiles.append((y0, y1, x0, x1, max(0, y0 - halo), min(h, y1 + halo), max(0, x0 - halo), min(w, x1 + halo)))
    return tiles


def tiled_forward(
    model: Model,
    x: torch.Tensor,
    upscale: int,
    tile: Tuple[int, int],
    halo: int,
    out: Optional[torch.Tensor] = None,
    only: Optional[Sequence[int]] = None,
    multiple: Optional[int] = None,
) -> torch.Tensor:
    """Run ``model`` tile by tile and stitch the centre crops on the device that holds ``x``.

    With ``halo`` >= the model's receptive radius the result equals ``model(x)`` bit for bit: inside a
    tile every layer sees exactly the values it would see in the full image for all pixels whose
    receptive field lies inside the extended region, and the extended region is only clipped where
    the image itself ends (where the untiled forward zero-pads too).  ``only`` restricts the work to a
    subset of tile indices (multi-GPU sharding); untouched output pixels are left as they are.
    ``multiple`` defaults to the model's ``tile_multiple`` (see plan_tiles).
    """
    n, _, h, w = x.shape
    if multiple is None:
        multiple = int(getattr(model, 'tile_multiple', 1))
    tiles = plan_tiles(h, w, tile[0], tile[1], halo, multiple)
    first = None
    for idx, (y0, y1, x0, x1, ey0, ey1, ex0, ex1) in enumerate(tiles):
        if only is not None and idx not in only:
            continue
        y = model(x[:, :, ey0:ey1, ex0:ex1].contiguous())
        if out is None:
            out = torch.empty((n, y.shape[1], h * upscale, w * upscale), dtype=y.dtype, device=y.device)
        if first is None:
            first = y
        cy0, cx0 = (y0 - ey0) * upscale, (x0 - ex0) * upscale
        out[:, :, y0 * upscale:y1 * upscale, x0 * upscale:x1 * upscale] = y[
            :, :, cy0:cy0 + (y1 - y0) * upscale, cx0:cx0 + (x1 - x0) * upscale
        ]
    if out is None:
        raise ValueError('no tile selected')
    return out


# ---------------------------------------------------------------------------------------- sharding
def shard_indices(count: int, rank: int, world_size: int) -> List[int]:
    """Indices of the units (frames, tiles) rank ``rank`` owns: round-robin, unit i -> rank i % world_size."""
    if not 0 <= rank < world_size:
        raise ValueError('rank out of range')
    return list(range(rank, count, world_size))


def gather_to_rank(local: Sequence[torch.Tensor], count: int, dst: int = 0, group=None) -> Optional[List[torch.Tensor]]:
    """Collect the outputs of round-robin sharded units on rank ``dst`` in unit order.

    ``local`` holds this rank's outputs in the order of ``shard_indices`` (possibly none, when there are fewer units than
    ranks); units may differ in shape (edge tiles of an image that the tile grid does not divide).  Returns the full list on
    ``dst`` and None elsewhere.  Shapes travel first (one small object gather), then every unit is one point-to-point message
    (NCCL send/recv over NVLink on GPUs, gloo in the CPU tests) — no padding, no copy through a stacked buffer.  This is the only
    communication of the engine; it is never on the compute path."""
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    local = [t.contiguous() for t in local]
    mine = shard_indices(count, rank, world)
    if len(local) != len(mine):
        raise ValueError(f'rank {rank} owns {len(mine)} of {count} units but passed {len(local)} tensors')
    if world == 1:
        return list(local)
    meta = [(tuple(t.shape), t.dtype) for t in local]
    all_meta = [None] * world if rank == dst else None
    dist.gather_object(meta, all_meta, dst=dst, group=group)
    peer = (lambda r: dist.get_global_rank(group, r)) if group is not None else (lambda r: r)
    if rank != dst:
        if local:
            for work in dist.batch_isend_irecv([dist.P2POp(dist.isend, t, peer(dst), group) for t in local]):
                work.wait()
        return None
    if local:
        device = local[0].device
    else:
        device = torch.device('cuda', torch.cuda.current_device()) if dist.get_backend(group) == 'nccl' else torch.device('cpu')
    out: List[Optional[torch.Tensor]] = [None] * count
    ops = []
    for r in range(world):
        for k, i in enumerate(shard_indices(count, r, world)):
            if r == dst:
                out[i] = local[k]
            else:
                shape, dtype = all_meta[r][k]
                out[i] = torch.empty(shape, dtype=dtype, device=device)
                ops.append(dist.P2POp(dist.irecv, out[i], peer(r), group))
    if ops:
        for work in dist.batch_isend_irecv(ops):
            work.wait()
    return out  # type: ignore[return-value]


# ---------------------------------------------------------------------------------------- frame streaming
class FramePipeline:
    """Host -> device -> host streaming of equally shaped frames through one model on one GPU.

    ``depth`` device slots are cycled; the H2D copy of frame i+1, the forward of frame i and the D2H copy
    of frame i-1 run concurrently on three streams.  Inputs should be pinned for the copies to be async.
    """

    def __init__(self, model: torch.nn.Module, upscale: int, device: torch.device, depth: int = 3):
        if device.type != 'cuda':
            raise RuntimeError('FramePipeline needs a CUDA device')
        self.model, self.upscale, self.device, self.depth = model, upscale, device, max(2, depth)
        self.s_in = torch.cuda.Stream(device)
        self.s_run = torch.cuda.Stream(device)
        self.s_out = torch.cuda.Stream(device)
        self._slots = None
        self._key = None

    def _ensure(self, frame: torch.Tensor, out_channels: int):
        key = (tuple(frame.shape), frame.dtype, out_channels)
        if key == self._key:
            return
        n, _, h, w = frame.shape
        self._slots = [
            (
                torch.empty(frame.shape, dtype=frame.dtype, device=self.device),
                torch.empty((n, out_channels, h * self.upscale, w * self.upscale), dtype=frame.dtype, device=self.device),
            )
            for _ in range(self.depth)
        ]
        self._key = key

    def run(self, frames: Sequence[torch.Tensor], outs: Optional[Sequence[torch.Tensor]] = None) -> List[torch.Tensor]:
        """Upscale host frames; returns host tensors (pinned when allocated here).  Blocks until all are done."""
        if len(frames) == 0:
            return []
        out_channels = getattr(self.model, 'out_channels', frames[0].shape[1])
        self._ensure(frames[0], out_channels)
        n, _, h, w = frames[0].shape
        if outs is None:
            outs = [
                torch.empty((n, out_channels, h * self.upscale, w * self.upscale), dtype=frames[0].dtype, pin_memory=True)
                for _ in frames
            ]
        ev_in = [torch.cuda.Event() for _ in frames]
        ev_run = [torch.cuda.Event() for _ in frames]
        ev_out = [torch.cuda.Event() for _ in frames]
        forward_into = getattr(self.model, 'forward_into', None)
        with torch.inference_mode():
            for i, frame in enumerate(frames):
                din, dout = self._slots[i % self.depth]
                with torch.cuda.stream(self.s_in):
                    if i >= self.depth:
                        self.s_in.wait_event(ev_run[i - self.depth])  # slot's previous forward has consumed din
                    din.copy_(frame, non_blocking=True)
                    ev_in[i].record(self.s_in)
                with torch.cuda.stream(self.s_run):
                    self.s_run.wait_event(ev_in[i])
                    if i >= self.depth:
                        self.s_run.wait_event(ev_out[i - self.depth])  # slot's previous result has left dout
                    if forward_into is not None:
                        forward_into(din, dout)
                    else:
                        dout.copy_(self.model(din))
                    ev_run[i].record(self.s_run)
                with torch.cuda.stream(self.s_out):
                    self.s_out.wait_event(ev_run[i])
                    outs[i].copy_(dout, non_blocking=True)
                    ev_out[i].record(self.s_out)
        self.s_out.synchronize()
        return list(outs)
