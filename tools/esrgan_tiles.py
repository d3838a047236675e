"""BASELINE.json config 4 in miniature: one large image through ESRGAN 4x (RRDBNet nb=23) as exact-halo tiles sharded over the
ranks of one box, stitched on the owning devices and gathered to rank 0 over NCCL; rank 0 also runs the untiled forward and checks
that the stitched result is bit-identical.

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/esrgan_tiles.py [H] [W] [grid_y] [grid_x]
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from resselt_b200.archs import RRDBNet
from resselt_b200.runner import gather_to_rank, plan_tiles, shard_indices

H = int(sys.argv[1]) if len(sys.argv) > 1 else 1080
W = int(sys.argv[2]) if len(sys.argv) > 2 else 1920
GY = int(sys.argv[3]) if len(sys.argv) > 3 else 1
GX = int(sys.argv[4]) if len(sys.argv) > 4 else 2
rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
local = int(os.environ.get('LOCAL_RANK', 0))
dev = torch.device('cuda', local)
torch.cuda.set_device(dev)
if world > 1:
    dist.init_process_group('nccl', device_id=dev)
model = RRDBNet(num_blocks=23, scale=4, seed=6).eval().to(dev).bfloat16()
x = torch.rand(1, 3, H, W, generator=torch.Generator().manual_seed(1)).to(dev, torch.bfloat16)  # same image on every rank
tile = ((H + GY - 1) // GY, (W + GX - 1) // GX)
halo = model.receptive_radius
tiles = plan_tiles(H, W, tile[0], tile[1], halo)
mine = shard_indices(len(tiles), rank, world)
assert H % GY == 0 and W % GX == 0, 'pick a grid that divides the image (gather_to_rank moves equally shaped tiles)'


def my_tiles():
    """Core crops (upscaled) of this rank's tiles, in shard order; every tile is computed from its halo-extended region."""
    crops = []
    for idx in mine:
        y0, y1, x0, x1, ey0, ey1, ex0, ex1 = tiles[idx]
        y = model(x[:, :, ey0:ey1, ex0:ex1].contiguous())
        cy0, cx0 = (y0 - ey0) * 4, (x0 - ex0) * 4
        crops.append(y[:, :, cy0:cy0 + (y1 - y0) * 4, cx0:cx0 + (x1 - x0) * 4].contiguous())
    return crops


with torch.inference_mode():
    my_tiles()  # warm-up (plans, workspaces)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    crops = my_tiles()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    every = gather_to_rank(crops, len(tiles), dst=0) if world > 1 else crops
    g1.record()
    torch.cuda.synchronize()
    if rank == 0:
        full = torch.empty(1, 3, H * 4, W * 4, device=dev, dtype=torch.bfloat16)
        for (y0, y1, x0, x1, *_), crop in zip(tiles, every):
            full[:, :, y0 * 4:y1 * 4, x0 * 4:x1 * 4] = crop
        ref = model(x)  # untiled, one GPU
        pixels = sum((t[1] - t[0]) * (t[3] - t[2]) for t in tiles)
        work = sum((t[5] - t[4]) * (t[7] - t[6]) for t in tiles)
        print(json.dumps(dict(image=[H, W], grid=[GY, GX], halo=halo, n_gpus=world, tiles=len(tiles), redundancy=round(work / pixels, 3),
                              compute_ms_max_over_ranks=round(float(ms), 2), gather_ms=round(g0.elapsed_time(g1), 2),
                              out_mp_per_s=round(H * W * 16 / 1e6 / float(ms) * 1e3, 1), bit_identical_to_untiled=bool(torch.equal(full, ref)))))
if world > 1:
    dist.destroy_process_group()
