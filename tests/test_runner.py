"""Host-side runner logic on CPU: tile planning, exact-halo stitching, round-robin sharding and the gather
(gloo, world_size 2) — with a plain torch conv stack standing in for the engine model."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn.functional as F

from resselt_b200.runner import gather_to_rank, plan_tiles, shard_indices, tiled_forward


def _toy_model(depth=3, upscale=2, seed=0):
    g = torch.Generator().manual_seed(seed)
    ws = [torch.randn(4, 3 if i == 0 else 4, 3, 3, generator=g, dtype=torch.float64) * 0.3 for i in range(depth)]
    last = torch.randn(3 * upscale * upscale, 4, 3, 3, generator=g, dtype=torch.float64) * 0.3

    def model(x):
        for w in ws:
            x = torch.tanh(F.conv2d(x, w, padding=1))
        return F.pixel_shuffle(F.conv2d(x, last, padding=1), upscale)

    return model, depth + 1  # receptive radius: one pixel per 3x3 conv


def test_plan_tiles_covers_image_once():
    tiles = plan_tiles(37, 50, 16, 20, 5)
    cover = torch.zeros(37, 50, dtype=torch.int32)
    for (y0, y1, x0, x1, ey0, ey1, ex0, ex1) in tiles:
        cover[y0:y1, x0:x1] += 1
        assert 0 <= ey0 <= y0 and y1 <= ey1 <= 37 and 0 <= ex0 <= x0 and x1 <= ex1 <= 50
        assert y0 - ey0 in (0, 5) or ey0 == 0
    assert bool((cover == 1).all()) and len(tiles) == 3 * 3
    with pytest.raises(ValueError):
        plan_tiles(8, 8, 0, 4, 1)


@pytest.mark.parametrize('hw,tile', [((40, 52), (16, 20)), ((17, 9), (8, 8)), ((30, 30), (64, 64))])
def test_exact_halo_tiling_equals_untiled(hw, tile):
    model, radius = _toy_model()
    x = torch.rand(2, 3, *hw, dtype=torch.float64, generator=torch.Generator().manual_seed(1))
    full = model(x)
    assert torch.equal(tiled_forward(model, x, 2, tile, halo=radius), full)
    if hw[0] > tile[0]:
        assert not torch.equal(tiled_forward(model, x, 2, tile, halo=radius - 1), full)


def test_sharded_tiles_union_is_complete():
    model, radius = _toy_model()
    x = torch.rand(1, 3, 33, 47, dtype=torch.float64, generator=torch.Generator().manual_seed(2))
    n_tiles = len(plan_tiles(33, 47, 12, 16, radius))
    out = torch.full((1, 3, 66, 94), float('nan'), dtype=torch.float64)
    for rank in range(3):
        tiled_forward(model, x, 2, (12, 16), radius, out=out, only=shard_indices(n_tiles, rank, 3))
    assert torch.equal(out, model(x))


def test_shard_indices():
    assert shard_indices(10, 0, 4) == [0, 4, 8] and shard_indices(10, 3, 4) == [3, 7]
    assert sorted(sum((shard_indices(10, r, 4) for r in range(4)), [])) == list(range(10))
    with pytest.raises(ValueError):
        shard_indices(4, 4, 4)


def _gather_worker(rank, world, port, count, tmpdir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    model, _ = _toy_model()
    # ragged units (edge tiles of an image the grid does not divide): every frame has its own width
    frames = [torch.rand(1, 3, 8, 10 + i, dtype=torch.float64, generator=torch.Generator().manual_seed(100 + i)) for i in range(count)]
    mine = [model(frames[i]) for i in shard_indices(count, rank, world)]  # no collective on the compute path
    got = gather_to_rank(mine, count, dst=0)
    if rank == 0:
        ok = got is not None and len(got) == count and all(torch.equal(got[i], model(frames[i])) for i in range(count))
        open(os.path.join(tmpdir, 'ok'), 'w').write('1' if ok else '0')
    else:
        assert got is None
    dist.destroy_process_group()


@pytest.mark.parametrize('count', [4, 5, 1])  # 1: fewer units than ranks, rank 1 owns nothing
def test_frame_sharding_and_gather_world_size_2(tmp_path, count):
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    mp.spawn(_gather_worker, args=(2, port, count, str(tmp_path)), nprocs=2, join=True)
    assert open(tmp_path / 'ok').read() == '1'


def test_plan_tiles_multiple_aligns_origins_and_rounds_the_halo():
    tiles = plan_tiles(91, 118, 44, 60, halo=19, multiple=2)
    assert all(t[4] % 2 == 0 and t[6] % 2 == 0 for t in tiles)          # extended origins on the model's grid
    assert tiles[0][5] == 44 + 20 and tiles[0][7] == 60 + 20            # halo 19 -> 20
    with pytest.raises(ValueError):
        plan_tiles(64, 64, 33, 32, halo=4, multiple=2)


def _nccl_gather_worker(rank, world, port, tmpdir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    count = 5
    # ragged bf16 "tiles": unit i has its own width; rank r owns units r, r + world, ...
    units = [torch.full((1, 3, 16, 24 + 8 * i), float(i + 1), dtype=torch.bfloat16) for i in range(count)]
    mine = [units[i].to(dev) for i in shard_indices(count, rank, world)]
    got = gather_to_rank(mine, count, dst=0)
    torch.cuda.synchronize()
    if rank == 0:
        ok = got is not None and len(got) == count and all(g.is_cuda and torch.equal(g.cpu(), u) for g, u in zip(got, units))
        open(os.path.join(tmpdir, 'ok'), 'w').write('1' if ok else '0')
    else:
        assert got is None
    dist.destroy_process_group()


@pytest.mark.gpu
def test_nccl_gather_world_size_2(tmp_path):
    """The engine's only communication on real hardware: ragged tiles gathered to rank 0 with NCCL send/recv (needs two GPUs)."""
    if torch.cuda.device_count() < 2:
        pytest.skip('needs two CUDA devices')
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    mp.spawn(_nccl_gather_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert open(tmp_path / 'ok').read() == '1'
