"""N>1 path on CPU: two gloo ranks agree on a disjoint, complete frame/tile partition and on the
max-over-ranks timing reduction bench.py uses (no collective touches pixel data)."""
import os

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from distraytracer_b200 import shard
    frames = shard.frames_for_rank(rank, world, 11, start=5)
    gathered = [None] * world
    dist.all_gather_object(gathered, frames)
    t = shard.max_over_ranks(10.0 + 5.0 * rank, dist)
    dist.barrier()
    q.put((rank, gathered, t))
    dist.destroy_process_group()


def test_two_ranks_partition_and_timing():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, gathered, t in res:
        allf = sorted(f for fr in gathered for f in fr)
        assert allf == list(range(5, 16))                      # complete and disjoint
        assert set(gathered[0]).isdisjoint(gathered[1])
        assert t == 15.0                                       # max over ranks


def test_bands_and_tiles_cover_the_frame():
    from distraytracer_b200 import shard
    for n in (1, 2, 3, 8):
        bands = shard.bands_for_devices(1080, n)
        assert bands[0][0] == 0 and bands[-1][1] == 1080
        assert all(bands[i][1] == bands[i + 1][0] for i in range(n - 1))
        cover = np.zeros((1080, 1920), dtype=np.int32)
        for dev, tiles in enumerate(shard.interleaved_tiles(1920, 1080, n, 64)):
            for t in tiles:
                assert t.device == dev
                cover[t.y0:t.y0 + t.height, t.x0:t.x0 + t.width] += 1
        assert (cover == 1).all()
    frame = np.zeros((8, 4, 3), dtype=np.uint8)
    shard.place_band(frame, np.full((3, 4, 3), 7, dtype=np.uint8), 2, 5)    # loop rows 2..4 -> PPM rows 3..5
    assert (frame[3:6] == 7).all() and frame[:3].sum() == 0 and frame[6:].sum() == 0


def test_render_frame_blocks_places_every_block_once():
    """The dynamic single-frame partition (host threads claiming row blocks) with stand-in renderers: every row is
    rendered exactly once, lands at its PPM position whatever handle claimed it, and errors surface on the caller."""
    import threading
    from distraytracer_b200 import shard, abi

    class Fake:
        def __init__(self, device, fail_at=None):
            self.device, self.fail_at, self.calls = device, fail_at, []

        def render(self, settings, tile, out=None):
            assert tile.device == self.device and tile.x0 == 0 and tile.width == settings.xRes
            assert out.shape == (tile.height, settings.xRes, 3)
            if self.fail_at is not None and tile.y0 >= self.fail_at:
                raise RuntimeError("boom")
            self.calls.append((tile.y0, tile.height))
            for r in range(tile.height):                      # buffer row r of a tile is loop row y0 + height - 1 - r
                out[r] += np.uint8(1 + (tile.y0 + tile.height - 1 - r) % 200)
            return out

    st = abi.Settings(); st.xRes, st.yRes = 16, 101
    assert shard.row_blocks(101, 15)[-1] == (90, 101) and len(shard.row_blocks(101, 15)) == 7
    frame = np.zeros((101, 16, 3), dtype=np.uint8)
    handles = [Fake(d) for d in (0, 0, 1, 1, 2)]
    dt = shard.render_frame_blocks(handles, st, frame, rows=15)
    assert dt >= 0 and sum(len(h.calls) for h in handles) == 7
    want = np.array([1 + (101 - 1 - y) % 200 for y in range(101)], dtype=np.uint8)   # PPM row y is loop row yRes-1-y
    assert (frame == want[:, None, None]).all()
    import pytest
    with pytest.raises(RuntimeError):
        shard.render_frame_blocks([Fake(0, fail_at=30)], st, frame, rows=15)
    with pytest.raises(ValueError):
        shard.render_frame_blocks(handles, st, np.zeros((5, 5, 3), dtype=np.uint8))
    assert threading.active_count() >= 1
