"""Run under torchrun with >= 2 GPUs: sharded CUDA path + nvc_allgather_visible (NCCL over NVLink) must reproduce the
single-rank oracle result.  Used by tests/test_gpu_multi.py and directly:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py"""
import ctypes
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _device_bytes(torch, ptr, nbytes, dev):
    """uint8 tensor view of raw device memory owned by the library (via __cuda_array_interface__)"""

    class _Raw:
        pass

    raw = _Raw()
    raw.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}
    return torch.as_tensor(raw, device=dev)


def main():
    import oracle_lib
    from niagara_b200 import layout, scenes, shard
    from niagara_b200.lib import check
    from niagara_b200.path import VisibilityPath

    rank, world, local_rank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist.init_process_group("nccl", device_id=dev)

    s = scenes.instanced_scene(os.path.join(ROOT, "tests", "golden", "kitten_pirate.nvcg"), 40003, screen=(1280, 720))
    local, base, bit_base, bits = shard.shard_draws(s.draws, s.meshes, rank, world)
    # every rank uses the same (largest) capacity: ncclAllGather needs equal slab sizes
    cap = max(shard.slab_capacity(e - b, s.meshes) for b, e in shard.partition(len(s.draws), world))
    cd = s.cull_data()
    cd.drawCount = len(local)
    g = VisibilityPath(s.meshes, s.meshlets, local, *s.screen, device=dev)
    g.set_visibility_bits(bits)
    depth = torch.from_numpy(s.depth).to(dev)
    lib = g.lib

    uid = torch.zeros(128, dtype=torch.uint8, device=dev)
    if rank == 0:
        buf = (ctypes.c_ubyte * 128)()
        check(lib.nvc_nccl_unique_id(buf), g.ctx, "nvc_nccl_unique_id")
        uid = torch.tensor(list(buf), dtype=torch.uint8, device=dev)
    dist.broadcast(uid, 0)
    check(lib.nvc_nccl_init(g.ctx, (ctypes.c_ubyte * 128)(*uid.cpu().tolist()), rank, world), g.ctx, "nvc_nccl_init")

    slab_bytes = cap * 20

    # copy-engine gather (nvc_gather_*): IPC tickets exchanged with torch.distributed
    ticket = (ctypes.c_ubyte * 192)()
    check(lib.nvc_gather_create(g.ctx, slab_bytes, rank, world, ticket), g.ctx, "nvc_gather_create")
    mine = torch.tensor(list(ticket), dtype=torch.uint8, device=dev)
    everyone = torch.zeros(world * 192, dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(everyone, mine)
    check(lib.nvc_gather_connect(g.ctx, (ctypes.c_ubyte * (192 * world))(*everyone.cpu().tolist())), g.ctx, "nvc_gather_connect")
    ce_slabs, ce_counts = ctypes.c_void_p(), ctypes.c_void_p()
    ce_host = torch.zeros(world * slab_bytes, dtype=torch.uint8).pin_memory()
    ce_counts_host = torch.zeros(world * 4, dtype=torch.int32).pin_memory()

    gathered = torch.zeros(world * slab_bytes, dtype=torch.uint8, device=dev)
    gathered_counts = torch.zeros(world * 4, dtype=torch.int32, device=dev)
    bases = [b for b, _ in shard.partition(len(s.draws), world)]
    bit_bases = [int(s.draws["meshletVisibilityOffset"][b]) for b in bases]

    if rank == 0:
        o = oracle_lib.OraclePath(s.meshes, s.meshlets, s.draws, *s.screen, threads=8)
        o.set_visibility_bits(s.visibility_bits)
        cd_all = s.cull_data()

    ok = True
    for frame in range(4):
        check(lib.nvc_gather_set_mode(g.ctx, frame % 2), g.ctx, "nvc_gather_set_mode")  # alternate copy engines / SM push
        for late in (False, True):
            if late:
                g.pyramid(depth)
            g.cull(cd, late)
            stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
            check(lib.nvc_allgather_visible(g.ctx, stream, ctypes.c_void_p(g.dcb.data_ptr()), slab_bytes, ctypes.c_void_p(g.dccb.data_ptr()), ctypes.c_void_p(gathered.data_ptr()), ctypes.c_void_p(gathered_counts.data_ptr())), g.ctx, "nvc_allgather_visible")
            check(lib.nvc_gather_push(g.ctx, stream, ctypes.c_void_p(g.dcb.data_ptr()), ctypes.c_void_p(g.dccb.data_ptr())), g.ctx, "nvc_gather_push")
            g.render_clusters(cd, late, cluster_backface=True)
            check(lib.nvc_gather_wait(g.ctx, stream), g.ctx, "nvc_gather_wait")
            # the receive buffers are double-buffered by frame parity: ask for the latest frame's
            check(lib.nvc_gather_buffers(g.ctx, ctypes.byref(ce_slabs), ctypes.byref(ce_counts)), g.ctx, "nvc_gather_buffers")
            torch.cuda.synchronize()
            # the copy-engine gather must have delivered exactly what NCCL delivered
            ce_view = _device_bytes(torch, ce_slabs.value, world * slab_bytes, dev)
            ce_cnt = _device_bytes(torch, ce_counts.value, world * 16, dev).view(torch.int32)
            same_ce = torch.equal(ce_cnt, gathered_counts)
            for r in range(world):
                n = int(gathered_counts[4 * r].item()) * 20
                same_ce = same_ce and torch.equal(ce_view[r * slab_bytes : r * slab_bytes + n], gathered[r * slab_bytes : r * slab_bytes + n])
            if not same_ce:
                ok = False
                print("rank", rank, "CE gather differs from NCCL gather, frame", frame, "late", late)
            # (no barrier here: the gather's own flow control keeps a fast rank from overwriting what a slow one still reads)
            slabs = gathered.cpu().numpy().reshape(world, slab_bytes)
            counts = gathered_counts.cpu().numpy().reshape(world, 4)
            cmds = shard.globalise_task_commands(list(slabs), list(counts), bases, bit_bases)
            if rank == 0:
                if late:
                    o.pyramid(s.depth)
                o.cull(cd_all, late)
                want = oracle_lib.sorted_commands(o.read_task_commands(int(o.dccb[0])))
                o.render_clusters(cd_all, late, cluster_backface=True)
                same = len(cmds) == len(want) and np.array_equal(oracle_lib.sorted_commands(cmds), want)
                # rank-local cluster counts must add up to the single-rank count
                ok = ok and same
                if not same:
                    print("MISMATCH frame", frame, "late", late, len(cmds), len(want))
            total = torch.tensor([int(g.ccb[0].item())], dtype=torch.int64, device=dev)
            dist.all_reduce(total)
            if rank == 0 and int(total.item()) != int(o.ccb[0]):
                ok = False
                print("cluster count mismatch", int(total.item()), int(o.ccb[0]))
    # ---- flow-control stress: 24 pushes back to back, no collective / barrier in between, ranks deliberately out of step
    # (rank r stalls its stream before every r-th push); every frame's gathered slabs must carry that frame's pattern ----
    pattern = torch.empty(slab_bytes // 4, dtype=torch.int32, device=dev)
    cnt = torch.zeros(4, dtype=torch.int32, device=dev)
    bad = torch.zeros(1, dtype=torch.int32, device=dev)
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    for f in range(24):
        check(lib.nvc_gather_set_mode(g.ctx, (f // 3) % 2), g.ctx, "nvc_gather_set_mode")
        if (f + rank) % (rank + 2) == 0:
            torch.cuda._sleep(int(2e8))  # ~0.1 s stall on this rank only
        pattern.fill_(1000 * f + rank)
        cnt.fill_(slab_bytes // 20)
        check(lib.nvc_gather_push(g.ctx, stream, ctypes.c_void_p(pattern.data_ptr()), ctypes.c_void_p(cnt.data_ptr())), g.ctx, "nvc_gather_push")
        check(lib.nvc_gather_wait(g.ctx, stream), g.ctx, "nvc_gather_wait")
        check(lib.nvc_gather_buffers(g.ctx, ctypes.byref(ce_slabs), ctypes.byref(ce_counts)), g.ctx, "nvc_gather_buffers")
        view = _device_bytes(torch, ce_slabs.value, world * slab_bytes, dev).view(torch.int32).view(world, slab_bytes // 4)
        want = (1000 * f + torch.arange(world, device=dev, dtype=torch.int32)).view(world, 1)
        bad += (view != want).any().to(torch.int32)  # enqueued on the stream, evaluated after the wait kernel
    torch.cuda.synchronize()
    if int(bad.item()) != 0:
        ok = False
        print("rank", rank, "flow-control stress: %d frames carried stale or torn slabs" % int(bad.item()))

    # ---- NVSwitch multicast transports over a symmetric region (modes 2 and 3), where the system offers them ----
    mc_state = "skipped"
    err = None
    try:
        import torch.distributed._symmetric_memory as symm_mem

        region = int(lib.nvc_gather_region_bytes(slab_bytes, world))
        t = symm_mem.empty(region, dtype=torch.uint8, device=dev)
        hdl = symm_mem.rendezvous(t, dist.group.WORLD)
        mc_ptr = int(hdl.multicast_ptr)
        if not mc_ptr:
            raise RuntimeError("no multicast mapping")
        peers = (ctypes.c_void_p * world)(*[int(p) for p in hdl.buffer_ptrs])
        check(lib.nvc_gather_attach(g.ctx, slab_bytes, rank, world, peers, ctypes.c_void_p(mc_ptr)), g.ctx, "nvc_gather_attach")
    except Exception as e:
        err = e
    have = torch.tensor([0 if err else 1], dtype=torch.int32, device=dev)
    dist.all_reduce(have, op=dist.ReduceOp.MIN)
    dist.barrier()
    if int(have.item()):
        mc_state = "checked"
        stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        for f in range(6):
            mode = 2 + (f % 2)  # multicast push kernel / fused into the late drawcull
            check(lib.nvc_gather_set_mode(g.ctx, mode), g.ctx, "nvc_gather_set_mode")
            g.cull(cd, False)
            g.render_clusters(cd, False, cluster_backface=True)
            g.pyramid(depth)
            if mode == 3:
                check(lib.nvc_gather_fuse_next_drawcull(g.ctx, stream), g.ctx, "nvc_gather_fuse_next_drawcull")
            g.cull(cd, True)
            check(lib.nvc_allgather_visible(g.ctx, stream, ctypes.c_void_p(g.dcb.data_ptr()), slab_bytes, ctypes.c_void_p(g.dccb.data_ptr()), ctypes.c_void_p(gathered.data_ptr()), ctypes.c_void_p(gathered_counts.data_ptr())), g.ctx, "nvc_allgather_visible")
            check(lib.nvc_gather_push(g.ctx, stream, ctypes.c_void_p(g.dcb.data_ptr()), ctypes.c_void_p(g.dccb.data_ptr())), g.ctx, "nvc_gather_push")
            g.render_clusters(cd, True, cluster_backface=True)
            check(lib.nvc_gather_wait(g.ctx, stream), g.ctx, "nvc_gather_wait")
            check(lib.nvc_gather_buffers(g.ctx, ctypes.byref(ce_slabs), ctypes.byref(ce_counts)), g.ctx, "nvc_gather_buffers")
            torch.cuda.synchronize()
            view = _device_bytes(torch, ce_slabs.value, world * slab_bytes, dev)
            cnt = _device_bytes(torch, ce_counts.value, world * 16, dev).view(torch.int32)
            same = torch.equal(cnt, gathered_counts)
            for r in range(world):
                n = (int(gathered_counts[4 * r].item()) + 63) // 64 * 64 * 20  # incl. the zero padding the consumer dispatches over
                same = same and torch.equal(view[r * slab_bytes : r * slab_bytes + n], gathered[r * slab_bytes : r * slab_bytes + n])
            if not same:
                ok = False
                print("rank", rank, "multicast gather (mode %d) differs from the NCCL gather, frame" % mode, f)
    elif rank == 0:
        print("multicast transports not checked:", str(err)[:200] if err else "unavailable on another rank")

    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)  # every rank checked its own gathered buffers
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("MULTI_GPU_CHECK", "OK" if int(flag.item()) == 1 else "FAILED", "world", world, "multicast", mc_state)
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
