"""2-GPU test (skipped on a 1-GPU box): z-slab sharded update + NCCL all-gather through
torch.distributed equals the single-GPU update, bit for bit."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, os.environ["LUM_ROOT"])
import lum_b200
from lum_b200 import scenes, distributed as D
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{rank}"))
sc = scenes.synthetic_scene(64, 48, 16, palette=64, seed=5)
wz, wy, wx = sc["world"].shape
st = lum_b200.RadianceStage((wx, wy, wz), 64, device=rank)
st.upload_world(sc["world"]); st.upload_bricks(0, sc["bricks"]); st.upload_materials(sc["materials"])
sp = np.array([(7, 9, 2, 0), (40, 40, 14, 0), (6, 8, 1, 0)], np.int16)
st.build_requests(sp)
field = D.device_tensor_view(st.field_device_ptr(), st.cells, f"cuda:{rank}")
stream = torch.cuda.ExternalStream(st.stream(), device=rank)
z0, z1 = D.slab_bounds(wz, rank, world)
for f in range(4):
    st.update_slab(f, 1, 0, z0, z1)
    with torch.cuda.stream(stream):
        D.all_gather_field(field, (wx, wy, wz), rank, world)
out = st.download_field()
np.save(os.path.join(os.environ["OUT_DIR"], f"f{rank}.npy"), out)
# the same four frames with the exchange fused into the trace (peer stores into every rank's inbox, one flag barrier)
st.upload_field(np.zeros(st.cells, np.uint32)); st.wait_idle(); dist.barrier()
D.open_peer_exchange(st, rank, world, device=f"cuda:{rank}")
for f in range(4):
    st.update_slab_exchange(f, 1, 0, z0, z1)
out = st.download_field()
dist.barrier()
np.save(os.path.join(os.environ["OUT_DIR"], f"g{rank}.npy"), out)
# and with interleaved shards instead of z-slabs (same peer set)
st.upload_field(np.zeros(st.cells, np.uint32)); st.wait_idle(); dist.barrier()
for f in range(4):
    st.update_shard_exchange(f, 1, 0)
out = st.download_field()
dist.barrier()
np.save(os.path.join(os.environ["OUT_DIR"], f"h{rank}.npy"), out)
st.peers_close()
dist.barrier()
st.close()
# a dispatch of many chunks per rank (pooled kernel, peer stores and commits of neighbouring frames in flight together),
# random initial field, the request list rebuilt before every frame WITHOUT a host sync (the count stays on the device)
sc = scenes.synthetic_scene(96, 80, 32, palette=64, seed=9)
wz, wy, wx = sc["world"].shape
st = lum_b200.RadianceStage((wx, wy, wz), 64, device=rank)
st.upload_world(sc["world"]); st.upload_bricks(0, sc["bricks"]); st.upload_materials(sc["materials"])
f0 = np.random.default_rng(31).integers(0, 2 ** 32, st.cells, dtype=np.uint64).astype(np.uint32)
st.upload_field(f0); st.wait_idle()
D.open_peer_exchange(st, rank, world, device=f"cuda:{rank}")
for f in range(5):
    st.build_requests(sp, sync=False)
    st.update_shard_exchange(100 + f, 1, 0)
out = st.download_field()
dist.barrier()
np.save(os.path.join(os.environ["OUT_DIR"], f"k{rank}.npy"), out)
# magicSize dispatches (size 3, shift = frame % 3) through the same exchange
for f in range(3):
    st.update_shard_exchange(200 + f, 3, f % 3)
out = st.download_field()
dist.barrier()
np.save(os.path.join(os.environ["OUT_DIR"], f"m{rank}.npy"), out)
st.peers_close()
dist.barrier(); dist.destroy_process_group()
'''


def test_two_gpu_slabs_match_single_gpu(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import lum_b200
    from lum_b200 import scenes
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    w = tmp_path / "worker.py"
    w.write_text(WORKER)
    env = dict(os.environ, LUM_ROOT=ROOT, OUT_DIR=str(tmp_path))
    subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                    "--master-addr", "127.0.0.1", "--master-port", str(port), str(w)],
                   check=True, env=env, timeout=600)
    sc = scenes.synthetic_scene(64, 48, 16, palette=64, seed=5)
    st = lum_b200.RadianceStage((64, 48, 16), 64)
    st.upload_world(sc["world"]); st.upload_bricks(0, sc["bricks"]); st.upload_materials(sc["materials"])
    st.build_requests(np.array([(7, 9, 2, 0), (40, 40, 14, 0), (6, 8, 1, 0)], np.int16))
    for f in range(4):
        st.update(f)
    ref = st.download_field()
    a, b = np.load(tmp_path / "f0.npy"), np.load(tmp_path / "f1.npy")
    assert np.array_equal(a, b)
    assert np.array_equal(a, ref)
    ga, gb = np.load(tmp_path / "g0.npy"), np.load(tmp_path / "g1.npy")
    assert np.array_equal(ga, ref) and np.array_equal(gb, ref)        # fused exchange: same bits
    ha, hb = np.load(tmp_path / "h0.npy"), np.load(tmp_path / "h1.npy")
    assert np.array_equal(ha, ref) and np.array_equal(hb, ref)        # interleaved shards: same bits
    st.close()
    # the large dispatch: five frames from a random field, list rebuilt without a sync before each
    sc = scenes.synthetic_scene(96, 80, 32, palette=64, seed=9)
    st = lum_b200.RadianceStage((96, 80, 32), 64)
    st.upload_world(sc["world"]); st.upload_bricks(0, sc["bricks"]); st.upload_materials(sc["materials"])
    n = st.build_requests(np.array([(7, 9, 2, 0), (40, 40, 14, 0), (6, 8, 1, 0)], np.int16))
    assert n > 20 * 3072
    st.upload_field(np.random.default_rng(31).integers(0, 2 ** 32, st.cells, dtype=np.uint64).astype(np.uint32))
    for f in range(5):
        st.update(100 + f)
    ref = st.download_field()
    ka, kb = np.load(tmp_path / "k0.npy"), np.load(tmp_path / "k1.npy")
    assert np.array_equal(ka, kb)
    assert np.array_equal(ka, ref)
    for f in range(3):
        st.update(200 + f, 3, f % 3)
    ref = st.download_field()
    ma, mb = np.load(tmp_path / "m0.npy"), np.load(tmp_path / "m1.npy")
    assert np.array_equal(ma, ref) and np.array_equal(mb, ref)


def test_peer_exchange_single_rank_degenerates_to_update(oracle_mod, demo):
    """nranks = 1: the peer commit writes only the local replica and the barriers are trivial."""
    from conftest import make_stage
    from lum_b200 import distributed as D
    st = make_stage(demo)
    n = st.build_requests(np.array([(7, 9, 2, 0), (40, 40, 14, 0)], np.int16))
    D.open_peer_exchange(st, 0, 1)
    for f in range(3):
        st.update_slab_exchange(f, 1, 0, 0, 16)
    got = st.download_field()
    st.upload_field(np.zeros(st.cells, np.uint32))
    for f in range(3):
        st.update(f)
    ref = st.download_field()
    assert np.array_equal(got, ref)
    st.upload_field(np.zeros(st.cells, np.uint32))
    for f in range(3):
        st.update_shard_exchange(f, 1, 0)            # one shard = the whole dispatch
    assert np.array_equal(st.download_field(), ref)
    st.peers_close()
    import lum_b200
    with pytest.raises(lum_b200.LumcuError):           # no peers any more
        st.update_shard_exchange(0, 1, 0)
    with pytest.raises(lum_b200.LumcuError):
        st.update_shard(0, 1, 0, 2, 2)                 # rank out of range
