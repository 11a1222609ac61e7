"""World-size-2 gloo tests (CPU) of the multi-GPU host logic: z-slab sharding + in-place field
all-gather reproduce the single-rank update.  The per-rank compute is the oracle here (test
infrastructure standing in for lumcu_update_slab, which needs a GPU); the sharding arithmetic and
the exchange are the product code under test (lum_b200/distributed.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lum_b200 import distributed as D
from lum_b200 import scenes


def test_slab_arithmetic():
    assert D.slab_bounds(16, 0, 8) == (0, 2) and D.slab_bounds(16, 7, 8) == (14, 16)
    assert D.slab_words((48, 48, 16), 1, 2) == (48 * 48 * 8, 48 * 48 * 16)
    with pytest.raises(ValueError):
        D.slab_bounds(16, 0, 3)
    cover = []
    for r in range(4):
        cover += list(range(*D.slab_bounds(64, r, 4)))
    assert cover == list(range(64))


@pytest.mark.parametrize("nwg", [0, 1, 3071, 3072, 3073, 6 * 3072 + 17, 100000])
@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_interleaved_shard_arithmetic(nwg, world):
    """The shards partition the dispatch, and the closed-form count (the one the launches are sized with)
    agrees with the enumeration."""
    seen = np.zeros(nwg, np.int32)
    for r in range(world):
        idx = D.shard_workgroups(nwg, r, world)
        assert len(idx) == D.shard_count(nwg, r, world)
        cont = D.shard_workgroups(nwg, r, world, 0)                    # LUMCU_SHARD_CONTIGUOUS
        assert len(cont) == D.shard_count(nwg, r, world, 0) and (len(cont) == 0 or (np.diff(cont) == 1).all())
        assert np.all(np.diff(idx) > 0)
        seen[idx] += 1
    assert np.all(seen == 1)
    for ch in (D.SHARD_CHUNK, 0, 48):
        own = D.shard_owner(nwg, world, ch)
        for r in range(world):
            assert np.array_equal(np.nonzero(own == r)[0], D.shard_workgroups(nwg, r, world, ch))
    ch = D.interleave_chunk(nwg, world)
    assert ch % 48 == 0 and (ch == 48 or ch * world * 64 <= nwg) and (ch == D.SHARD_CHUNK or 2 * ch * world * 64 > nwg)
    sizes = [D.shard_count(nwg, r, world) for r in range(world)]
    assert max(sizes) - min(sizes) <= D.SHARD_CHUNK


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, frames, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle as O
    d = scenes.demo_scene()
    sc = O.Scene(d["world"], d["bricks"], d["materials"])
    dims = (sc.wx, sc.wy, sc.wz)
    sp = np.array([(7, 9, 2, 0), (40, 40, 14, 0), (6, 8, 1, 0)], np.int16)      # unsorted special tail
    req = O.build_requests(d["world"], sp)
    z0, z1 = D.slab_bounds(sc.wz, rank, world)
    mine = np.nonzero((req[:, 2] >= z0) & (req[:, 2] < z1))[0]                   # this rank's workgroups
    field = torch.zeros(sc.cells, dtype=torch.int32)
    for f in range(frames):
        old = field.numpy().view(np.uint32).copy()
        new = old.copy()
        # contiguous runs of this rank's workgroup indices, computed with their GLOBAL indices
        runs = np.split(mine, np.nonzero(np.diff(mine) != 1)[0] + 1) if len(mine) else []
        for run in runs:
            r = O.update(sc, old, req, f, wg_range=(int(run[0]), int(run[-1]) + 1), nthreads=2)
            ch = r["field"] != old
            new[ch] = r["field"][ch]
        field = torch.from_numpy(new.view(np.int32).copy())
        D.all_gather_field(field, dims, rank, world)
    np.save(os.path.join(out_dir, f"field_{rank}.npy"), field.numpy().view(np.uint32))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_gloo_matches_single_rank(tmp_path):
    import oracle as O
    frames = 3
    port = _free_port()
    mp.spawn(_worker, args=(2, port, frames, str(tmp_path)), nprocs=2, join=True)
    d = scenes.demo_scene()
    sc = O.Scene(d["world"], d["bricks"], d["materials"])
    sp = np.array([(7, 9, 2, 0), (40, 40, 14, 0), (6, 8, 1, 0)], np.int16)
    req = O.build_requests(d["world"], sp)
    f = np.zeros(sc.cells, np.uint32)
    for k in range(frames):
        f = O.update(sc, f, req, k)["field"]
    a = np.load(tmp_path / "field_0.npy")
    b = np.load(tmp_path / "field_1.npy")
    assert np.array_equal(a, b)          # every rank ends with the identical field
    assert np.array_equal(a, f)          # and it is the single-rank result


def _shard_worker(rank, world, port, frames, out_dir, chunk):
    """The frame of lumcu_update_shard_exchange with gloo in the place of NVLink: every rank traces its shard
    (interleaved chunks, or contiguous ranges for chunk == 0) into its OWN staging array indexed by workgroup, the
    staging arrays are exchanged (the kernel reads them in place over peer memory; here an all_gather), and every rank
    commits the WHOLE dispatch locally, taking the texel of workgroup w from the staging array of w's owner
    (D.shard_owner, the arithmetic of gather_commit_kernel).  A small chunk makes the demo's 2,692 workgroups span
    several chunks."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle as O
    d = scenes.demo_scene()
    sc = O.Scene(d["world"], d["bricks"], d["materials"])
    sp = np.array([(7, 9, 2, 0), (40, 40, 14, 0), (6, 8, 1, 0)], np.int16)
    req = O.build_requests(d["world"], sp)
    nwg = len(req)
    mine = D.shard_workgroups(nwg, rank, world, chunk=chunk)
    owner = D.shard_owner(nwg, world, chunk)
    assert np.array_equal(np.nonzero(owner == rank)[0], mine)
    cell = req[:, 0].astype(np.int64) + sc.wx * (req[:, 1].astype(np.int64) + sc.wy * req[:, 2].astype(np.int64))
    field = np.zeros(sc.cells, np.uint32)
    for f in range(frames):
        staged = np.full(nwg, 0xDEADBEEF, np.uint32)                      # only this rank's workgroups get written
        runs = np.split(mine, np.nonzero(np.diff(mine) != 1)[0] + 1) if len(mine) else []
        for run in runs:
            r = O.update(sc, field, req, f, wg_range=(int(run[0]), int(run[-1]) + 1), nthreads=2)
            staged[run] = r["field"][cell[run]]
        all_staged = [None] * world
        dist.all_gather_object(all_staged, staged)
        stacked = np.stack(all_staged)
        field[cell] = stacked[owner, np.arange(nwg)]                      # gather from the owner + commit, whole dispatch
    np.save(os.path.join(out_dir, f"shard_{rank}.npy"), field)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
@pytest.mark.parametrize("chunk", [384, 0], ids=["interleaved", "contiguous"])
def test_two_rank_gloo_interleaved_shards_match_single_rank(tmp_path, chunk):
    import oracle as O
    frames = 3
    port = _free_port()
    mp.spawn(_shard_worker, args=(2, port, frames, str(tmp_path), chunk), nprocs=2, join=True)
    d = scenes.demo_scene()
    sc = O.Scene(d["world"], d["bricks"], d["materials"])
    req = O.build_requests(d["world"], np.array([(7, 9, 2, 0), (40, 40, 14, 0), (6, 8, 1, 0)], np.int16))
    f = np.zeros(sc.cells, np.uint32)
    for k in range(frames):
        f = O.update(sc, f, req, k)["field"]
    a, b = np.load(tmp_path / "shard_0.npy"), np.load(tmp_path / "shard_1.npy")
    assert np.array_equal(a, b) and np.array_equal(a, f)
