"""Multi-GPU sharding of the visibility path; host-side logic only.

Every draw — with its meshlet instances, its dvb word and its mvb bit range — is independent, so draws are split into
contiguous ranges, one per rank.  Visibility-bit offsets restart at 0 inside a shard; ids emitted by a rank are
rank-local and are globalised by adding the shard's base.  The only exchange is the allgather of the per-rank
command slabs + counters (NCCL through the C ABI on GPUs; any torch.distributed backend for the host logic)."""
import numpy as np

from . import host, layout


def partition(draw_count, world):
    """Contiguous draw ranges [begin, end) per rank."""
    return [(draw_count * r // world, draw_count * (r + 1) // world) for r in range(world)]


def shard_draws(draws, meshes, rank, world):
    """This rank's draws with rank-local meshletVisibilityOffset; returns (draws, base_draw, base_bit, bit_count)."""
    begin, end = partition(len(draws), world)[rank]
    local = draws[begin:end].copy()
    base_bit = int(draws["meshletVisibilityOffset"][begin]) if begin < len(draws) else 0
    bits, _ = host.visibility_offsets(local, meshes)
    return local, begin, base_bit, bits


def slab_capacity(local_draw_count, meshes):
    """Fixed per-rank slab capacity (commands): every draw visible at its largest LOD, padded to x64."""
    max_groups = int(((meshes["lods"]["meshletCount"].max(axis=1) + 63) // 64).max()) if len(meshes) else 1
    return (local_draw_count * max(1, max_groups) + 63) // 64 * 64


def allgather_slabs(local_slab, local_count4, group=None):
    """torch.distributed allgather of fixed-capacity slabs + 4-word counters (host logic / CPU tests; the GPU path
    uses nvc_allgather_visible).  local_slab: uint8 tensor, local_count4: int32[4]."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    slabs = [torch.empty_like(local_slab) for _ in range(world)]
    counts = [torch.empty_like(local_count4) for _ in range(world)]
    dist.all_gather(slabs, local_slab, group=group)
    dist.all_gather(counts, local_count4, group=group)
    return slabs, counts


def globalise_task_commands(slabs, counts, bases, bit_bases, task_wglimit=layout.TASK_WGLIMIT):
    """Concatenates the valid prefix of every rank's MeshTaskCommand slab, rebasing drawId and
    meshletVisibilityOffset from rank-local to global."""
    out = []
    for slab, count4, base, bit_base in zip(slabs, counts, bases, bit_bases):
        n = min(int(np.uint32(count4[0])), task_wglimit)
        cmds = np.frombuffer(np.ascontiguousarray(slab).tobytes(), dtype=layout.MESHTASKCOMMAND_DTYPE, count=n).copy()
        cmds["drawId"] += np.uint32(base)
        cmds["meshletVisibilityOffset"] += np.uint32(bit_base)
        out.append(cmds)
    return np.concatenate(out) if out else np.zeros(0, dtype=layout.MESHTASKCOMMAND_DTYPE)
