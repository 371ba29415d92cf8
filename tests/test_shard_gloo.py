"""CPU-only, world_size 2 over gloo: the sharded path (contiguous draw ranges, rank-local visibility offsets, allgather
of the per-rank command slabs, globalisation) reproduces the single-rank result exactly.  The per-rank compute is
the oracle here (no GPU in this container); the GPU variant of the same check is tests/test_gpu_multi.py."""
import os
import sys
import tempfile

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, init_file, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    from niagara_b200 import layout, scenes, shard

    dist.init_process_group("gloo", init_method="file://" + init_file, rank=rank, world_size=world)
    s = scenes.instanced_scene(os.path.join(ROOT, "tests", "golden", "kitten_pirate.nvcg"), 6001, screen=(640, 480))
    local, base, bit_base, bits = shard.shard_draws(s.draws, s.meshes, rank, world)
    cap = shard.slab_capacity(len(local), s.meshes)
    cd = s.cull_data()
    cd.drawCount = len(local)
    o = oracle_lib.OraclePath(s.meshes, s.meshlets, local, *s.screen, cmd_capacity=cap)
    o.set_visibility_bits(bits)
    results = []
    for frame in range(2):
        for late in (False, True):
            if late:
                o.pyramid(s.depth)
            o.cull(cd, late)
            slab = torch.from_numpy(o.dcb[: cap * 20].copy())
            count4 = torch.from_numpy(o.dccb.astype(np.int32))
            slabs, counts = shard.allgather_slabs(slab, count4)
            bases = [b for b, _ in shard.partition(len(s.draws), world)]
            bit_bases = [int(s.draws["meshletVisibilityOffset"][b]) for b in bases]
            g = shard.globalise_task_commands([x.numpy() for x in slabs], [c.numpy() for c in counts], bases, bit_bases)
            results.append(g)
            o.render_clusters(cd, late, cluster_backface=True)
    if rank == 0:
        np.save(os.path.join(out_dir, "gathered.npy"), np.concatenate([r.view(np.uint32).reshape(-1, 5) for r in results]))
        np.save(os.path.join(out_dir, "sizes.npy"), np.array([len(r) for r in results]))
    # every rank must hold the identical gathered result
    digest = torch.tensor([int(np.concatenate([r.view(np.uint32).reshape(-1) for r in results]).astype(np.uint64).sum() % (1 << 31))])
    both = [torch.zeros_like(digest) for _ in range(world)]
    dist.all_gather(both, digest)
    assert all(int(b) == int(digest) for b in both)
    dist.destroy_process_group()


def test_two_rank_shard_equals_single_rank():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    from niagara_b200 import layout, scenes

    world = 2
    with tempfile.TemporaryDirectory() as d:
        init_file = os.path.join(d, "init")
        mp.spawn(_worker, args=(world, init_file, d), nprocs=world, join=True)
        gathered = np.load(os.path.join(d, "gathered.npy"))
        sizes = np.load(os.path.join(d, "sizes.npy"))

    s = scenes.instanced_scene(os.path.join(ROOT, "tests", "golden", "kitten_pirate.nvcg"), 6001, screen=(640, 480))
    cd = s.cull_data()
    o = oracle_lib.OraclePath(s.meshes, s.meshlets, s.draws, *s.screen)
    o.set_visibility_bits(s.visibility_bits)
    off = 0
    i = 0
    for frame in range(2):
        for late in (False, True):
            if late:
                o.pyramid(s.depth)
            o.cull(cd, late)
            n = int(o.dccb[0])
            want = oracle_lib.sorted_commands(o.read_task_commands(n))
            got = gathered[off : off + sizes[i]].copy().view(layout.MESHTASKCOMMAND_DTYPE).reshape(-1)
            assert sizes[i] == n
            assert np.array_equal(oracle_lib.sorted_commands(got), want), (frame, late)
            off += sizes[i]
            i += 1
            o.render_clusters(cd, late, cluster_backface=True)
    assert off == len(gathered) and sizes.sum() > 0


def test_partition_covers_everything():
    from niagara_b200 import shard

    for n in (0, 1, 7, 1000003):
        for w in (1, 2, 3, 8):
            parts = shard.partition(n, w)
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
            sizes = [e - b for b, e in parts]
            assert max(sizes) - min(sizes) <= 1
