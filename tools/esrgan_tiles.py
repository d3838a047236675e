"""BASELINE.json config 4: one large image through ESRGAN 4x (RRDBNet nb=23) as exact-halo tiles sharded over the ranks of one
box (no collective on the compute path), then gathered to rank 0 over NCCL (send/recv per tile, the only communication).

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/esrgan_tiles.py [H] [W] [grid_y] [grid_x] [--check full|windows]

Default: the 8K input of config 4 (4320 x 7680) as a 2 x 4 grid — one tile per GPU at N = 8, halo 349 px.
Checks on rank 0:
  full     the untiled forward on one GPU, stitched result must be bit-identical (only for images whose untiled workspace fits
           180 GB: up to ~1080p x 2)
  windows  (default) size-independent: small windows straddling every interior tile corner / edge are recomputed on their own with
           the exact halo and must equal the stitched output bit for bit — the untiled 8K forward needs > 200 GB and cannot be run
Prints one JSON line (rank 0).
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from resselt_b200.archs import RRDBNet
from resselt_b200.runner import gather_to_rank, plan_tiles, shard_indices

argv = [a for a in sys.argv[1:] if not a.startswith('--')]
check = 'windows'
if '--check' in sys.argv:
    check = sys.argv[sys.argv.index('--check') + 1]
    argv = [a for a in argv if a != check]
H = int(argv[0]) if len(argv) > 0 else 4320
W = int(argv[1]) if len(argv) > 1 else 7680
GY = int(argv[2]) if len(argv) > 2 else 2
GX = int(argv[3]) if len(argv) > 3 else 4
rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
local = int(os.environ.get('LOCAL_RANK', 0))
dev = torch.device('cuda', local)
torch.cuda.set_device(dev)
if world > 1:
    dist.init_process_group('nccl', device_id=dev)
S = 4
model = RRDBNet(num_blocks=23, scale=S, seed=6).eval().to(dev).bfloat16()
x = torch.rand(1, 3, H, W, generator=torch.Generator().manual_seed(1)).to(dev, torch.bfloat16)  # same image on every rank
tile = ((H + GY - 1) // GY, (W + GX - 1) // GX)
halo = model.receptive_radius
tiles = plan_tiles(H, W, tile[0], tile[1], halo)
mine = shard_indices(len(tiles), rank, world)


def crop_of(y0, y1, x0, x1):
    """Upscaled core region [y0,y1) x [x0,x1), computed on its own from the halo-extended input region."""
    ey0, ey1, ex0, ex1 = max(0, y0 - halo), min(H, y1 + halo), max(0, x0 - halo), min(W, x1 + halo)
    y = model(x[:, :, ey0:ey1, ex0:ex1].contiguous())
    cy0, cx0 = (y0 - ey0) * S, (x0 - ex0) * S
    return y[:, :, cy0:cy0 + (y1 - y0) * S, cx0:cx0 + (x1 - x0) * S].contiguous()


def my_tiles():
    return [crop_of(*tiles[idx][:4]) for idx in mine]


with torch.inference_mode():
    t_wall = time.perf_counter()
    my_tiles()  # warm-up: plan, workspace, tensor maps
    torch.cuda.synchronize()
    if world > 1:
        # warm the communicator too (rendezvous, channel set-up), so that gather_ms below is the wire, not NCCL's bring-up
        gather_to_rank([torch.zeros(1, 3, 8, 8, device=dev, dtype=torch.bfloat16) for _ in mine], len(tiles), dst=0)
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    crops = my_tiles()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.barrier()
    torch.cuda.synchronize()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    every = gather_to_rank(crops, len(tiles), dst=0) if world > 1 else crops
    g1.record()
    torch.cuda.synchronize()
    if rank == 0:
        full = torch.empty(1, 3, H * S, W * S, device=dev, dtype=torch.bfloat16)
        for (y0, y1, x0, x1, *_), crop in zip(tiles, every):
            full[:, :, y0 * S:y1 * S, x0 * S:x1 * S] = crop
        gathered_bytes = sum(c.numel() * c.element_size() for i, c in enumerate(every) if i % world != 0)
        del every, crops
        torch.cuda.empty_cache()
        checked, ok = 0, True
        if check == 'full':
            ok = bool(torch.equal(full, model(x)))
            checked = 1
        else:
            # windows of 48 x 64 px centred on every interior grid crossing and on the middle of every interior tile edge
            ys = sorted({t[0] for t in tiles} - {0})
            xs = sorted({t[2] for t in tiles} - {0})
            centres = [(y, xx) for y in ys for xx in xs] + [(y, W // (2 * GX)) for y in ys] + [(H // (2 * GY), xx) for xx in xs]
            for cy, cx in centres:
                y0, x0 = max(0, cy - 24), max(0, cx - 32)
                y1, x1 = min(H, y0 + 48), min(W, x0 + 64)
                ok = ok and bool(torch.equal(crop_of(y0, y1, x0, x1), full[:, :, y0 * S:y1 * S, x0 * S:x1 * S]))
                checked += 1
        pixels = sum((t[1] - t[0]) * (t[3] - t[2]) for t in tiles)
        work = sum((t[5] - t[4]) * (t[7] - t[6]) for t in tiles)
        gather_ms = g0.elapsed_time(g1)
        flop_px = 35853696  # algorithmic FLOP per input pixel of RRDBNet nb23 4x
        print(json.dumps(dict(
            config='BASELINE config 4: ESRGAN RRDBNet 4x nb23, bf16, exact-halo tiles, one process per GPU', image=[H, W], grid=[GY, GX], halo=halo,
            n_gpus=world, tiles=len(tiles), redundancy=round(work / pixels, 3), compute_ms_max_over_ranks=round(float(ms), 2),
            gather_ms=round(gather_ms, 2), gathered_gb=round(gathered_bytes / 1e9, 3), gather_gbs_into_rank0=round(gathered_bytes / gather_ms / 1e6, 1),
            out_mp_per_s=round(H * W * S * S / 1e6 / float(ms) * 1e3, 1), out_mp_per_s_with_gather=round(H * W * S * S / 1e6 / (float(ms) + gather_ms) * 1e3, 1),
            algorithmic_tflops=round(flop_px * H * W / float(ms) / 1e9, 1), executed_tflops=round(flop_px * work / float(ms) / 1e9, 1),
            check=check, windows_checked=checked, bit_identical=ok, wall_s=round(time.perf_counter() - t_wall, 1))))
if world > 1:
    dist.destroy_process_group()
