"""Multi-GPU plumbing of the radiance stage: probe sharding + per-frame field exchange.

Design: the read-only world / palette / materials are replicated on every
rank; the field is replicated too and updated Jacobi-style, so probes shard with one exchange per frame and N ranks
compute bit-identically to one (seeds use the global workgroup index).

* NCCL baseline: rank r of N owns the contiguous z-slab [r*Z/N, (r+1)*Z/N) of the field (the field is z-major, so a
  slab is one contiguous range of words), updates exactly the probes of that slab (lumcu_update_slab) and all ranks
  exchange their slabs with ONE in-place all-gather (all_gather_field).
* Peer exchange (open_peer_exchange wires it; the default of bench.py): lumcu_update_shard_exchange shards the
  dispatch by workgroup index -- INTERLEAVED chunks by default (every rank samples every part of the world, so the
  ranks finish together whatever the scene looks like; chunk size: stage.shard_chunk), or CONTIGUOUS equal ranges
  (lumcu_set_shard_mode) -- traces into the rank's own staging array, publishes its arrival in every rank's flag array
  and then reads every texel of the dispatch from its owner's staging array over NVLink into the local field: one
  flag barrier and one gather/commit kernel per frame, no collective call.  lumcu_update_slab_exchange does the same
  for z-slabs with the transfer as a push from the trace kernel.

torch.distributed is plumbing only (NCCL over NVLink on GPUs; gloo on CPU for the tests of this host-side logic).
"""
from __future__ import annotations


SHARD_CHUNK = 3072      # = IL_CHUNK of csrc/lumcu.cu: a multiple of every pooled kernel's probes per CTA


def shard_workgroups(nwg: int, rank: int, world_size: int, chunk: int = SHARD_CHUNK):
    """Workgroup indices of [0, nwg) that shard `rank` of `world_size` owns, in the order the kernels visit them
    (lumcu_update_shard).  chunk > 0 (LUMCU_SHARD_INTERLEAVED): chunk c of the dispatch belongs to rank
    c % world_size (the stage's default).  chunk == 0 (LUMCU_SHARD_CONTIGUOUS): [nwg*rank/n, nwg*(rank+1)/n)."""
    import numpy as np
    w = np.arange(nwg, dtype=np.int64)
    if chunk == 0:
        return w[nwg * rank // world_size: nwg * (rank + 1) // world_size]
    return w[(w // chunk) % world_size == rank]


def shard_count(nwg: int, rank: int, world_size: int, chunk: int = SHARD_CHUNK) -> int:
    """len(shard_workgroups(...)) in closed form (what sizes the shard's launches)."""
    if chunk == 0:
        return nwg * (rank + 1) // world_size - nwg * rank // world_size
    full, rem = divmod(nwg, chunk)
    return (full // world_size) * chunk + (chunk if full % world_size > rank else 0) + (rem if full % world_size == rank else 0)


def shard_owner(nwg: int, world_size: int, chunk: int = SHARD_CHUNK):
    """owner[w] = the rank whose shard holds workgroup w -- what gather_commit_kernel evaluates to know whose staging
    array to read the texel of w from.  Interleaved: (w // chunk) % n.  Contiguous (chunk == 0): the number of shard
    boundaries nwg*k//n (k = 1..n-1) that are <= w."""
    import numpy as np
    w = np.arange(nwg, dtype=np.int64)
    if chunk == 0:
        bounds = np.array([nwg * k // world_size for k in range(1, world_size)], np.int64)
        return np.searchsorted(bounds, w, side="right")
    return (w // chunk) % world_size


def interleave_chunk(total: int, world_size: int) -> int:
    """The chunk size lumcu_update_shard_exchange deals a dispatch of `total` workgroups out in (lumcu_shard_chunk):
    3072, halved until every rank gets at least 64 chunks, not below 48."""
    ch = SHARD_CHUNK
    while ch > 48 and ch * world_size * 64 > total:
        ch //= 2
    return ch


def slab_bounds(wz: int, rank: int, world_size: int):
    """z range [z0, z1) owned by `rank`.  Requires wz % world_size == 0 (16 and 64 both divide by 8)."""
    if wz % world_size:
        raise ValueError(f"world z extent {wz} is not divisible by {world_size} ranks")
    h = wz // world_size
    return rank * h, (rank + 1) * h


def slab_words(dims, rank: int, world_size: int):
    """[begin, end) of the rank's slab in field words (field index = x + X*(y + Y*z))."""
    wx, wy, wz = dims
    z0, z1 = slab_bounds(wz, rank, world_size)
    return wx * wy * z0, wx * wy * z1


def all_gather_field(field, dims, rank: int, world_size: int, group=None):
    """In-place all-gather of the ranks' z-slabs of `field` (a flat 1-D torch tensor of the whole
    field, on CPU for gloo or on the GPU for NCCL).  The send buffer is the rank's own slab inside
    the receive buffer, the layout ncclAllGather accepts in place."""
    import torch.distributed as dist
    if world_size == 1:
        return field
    b, e = slab_words(dims, rank, world_size)
    shard = field[b:e]
    if field.is_cuda:
        dist.all_gather_into_tensor(field, shard, group=group)
    else:  # gloo has no all_gather_into_tensor on every build: gather views of the same storage
        n = e - b
        outs = [field[i * n:(i + 1) * n] for i in range(world_size)]
        dist.all_gather(outs, shard.clone(), group=group)
    return field


def device_tensor_view(ptr: int, nwords: int, device):
    """A torch int32 tensor aliasing `nwords` 32-bit words of device memory at `ptr` (the field
    buffer owned by the stage), so NCCL can exchange it without a copy."""
    import torch

    class _Iface:
        __cuda_array_interface__ = {"shape": (nwords,), "typestr": "<i4", "data": (ptr, False), "version": 3}

    return torch.as_tensor(_Iface(), device=device)


def open_peer_exchange(stage, rank: int, world_size: int, device=None, group=None):
    """Wire the peer exchange (lumcu_update_shard_exchange / lumcu_update_slab_exchange): export this rank's IPC blob,
    all-gather the blobs through torch.distributed (NCCL: a small uint8 device tensor; gloo: CPU) and
    open the peers' staging / flag buffers.  Every rank must call it (collective)."""
    import numpy as np
    import torch
    import torch.distributed as dist
    blob = stage.ipc_export()
    if world_size == 1:
        stage.peers_open(0, blob.reshape(1, -1))
        return
    t = torch.from_numpy(blob.copy())
    if device is not None:
        t = t.to(device)
    out = torch.empty(world_size * t.numel(), dtype=torch.uint8, device=t.device)
    dist.all_gather_into_tensor(out, t, group=group)
    blobs = out.cpu().numpy().reshape(world_size, -1)
    stage.peers_open(rank, blobs)
    dist.barrier(group=group)        # nobody starts exchanging before everyone has mapped everyone
