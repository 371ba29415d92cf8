"""CPU-only MODEL of the multi-GPU exchange protocol of csrc/nvc_peer.cu (nvc_gather_push / _wait / _graph_advance): the
stream programs the C ABI enqueues are restated as lists of operations over the same device-side state (data flags,
acknowledgements, receive buffers double-buffered by tag parity, the graph epoch) and executed under random interleavings of all
ranks' streams.  Checked for every schedule: no deadlock; a consumer that reads the gathered buffers between wait(T) and the next
wait sees exactly frame T from every rank; no receive buffer is written while it may still be read.  Covered: the eager frame loop
of bench.py (wait of frame T-1 in front of the late drawcull of frame T), CUDA-graph replay of 4-frame graphs with
nvc_gather_graph_advance (tags baked into the graph, shifted by the device-side epoch), 2-8 ranks, and — as a negative control — the
protocol WITHOUT acknowledgements, which the checker must catch overwriting a buffer that is being read.

This is a model of the protocol's logic, not of the CUDA code: the hardware run is tests/multi_gpu_check.py (2 and 8 GPUs,
profiles/r2_multi_gpu_check_n*.txt)."""
import random

import pytest


class Rank:
    def __init__(self, r, world):
        self.r, self.world = r, world
        self.flags = [0] * world  # flags[q]: latest frame tag whose data from rank q has landed here
        self.acks = [0] * world  # acks[q]: latest frame tag rank q has finished reading from here
        self.buf = [[None] * world for _ in range(2)]  # buf[parity][q]: effective tag of the slab stored
        self.reading = [False, False]  # a consumer is reading buf[parity]
        self.epoch = 0
        # host-side state of the C ABI
        self.tag = 0
        self.acked_tag = 0


def build_programs(world, frames, graph_frames=0, use_acks=True, capture_fails=False):
    """Per rank: a dict stream -> list of ops.  An op is (kind, args...); cross-stream order is expressed with events
    ("record", name) / ("wait_event", name) like cudaEventRecord / cudaStreamWaitEvent."""
    ranks = [Rank(r, world) for r in range(world)]
    programs = []
    for rk in ranks:
        main, lead = [], []

        def push(rk=rk, main=main, lead=lead):
            rk.tag += 1
            t = rk.tag
            ev = "fork%d" % t
            main.append(("record", ev))
            lead.append(("wait_event", ev))
            if use_acks and t >= 3:
                lead.append(("spin_acks", t - 2))
            for k in range(rk.world):
                lead.append(("copy", (rk.r + k) % rk.world, t))
            lead.append(("raise_flags", t))
            lead.append(("record", "join%d" % t))
            return t

        def wait(rk=rk, main=main):
            t = rk.tag
            if use_acks and t >= 2 and rk.acked_tag != t - 1:
                main.append(("raise_acks", t - 1))
                rk.acked_tag = t - 1
            main.append(("wait_event", "join%d" % t))
            main.append(("spin_flags", t))

        def eager_frame(pending):
            main.append(("work", "early passes"))
            if pending:
                main.append(("read_end",))  # the consumer of the previous frame's gathered data is done before the wait's ack
                wait()
                main.append(("read_begin", rk.tag))
            main.append(("work", "drawcull late"))
            push()
            main.append(("work", "cluster late"))

        # eager frames (bench.py's frame(): the wait of frame T-1 sits in front of the late drawcull of frame T)
        pending = False
        for _ in range(frames):
            eager_frame(pending)
            pending = True
        main.append(("read_end",)) if frames > 1 else None
        wait()
        main.append(("read_begin", rk.tag))
        main.append(("read_end",))

        # CUDA-graph replay: the ops of `graph_frames` frames are captured ONCE (host tags baked) and replayed `replays` times
        if graph_frames:
            g_main_start, g_lead_start = len(main), len(lead)
            pending = False
            for _ in range(graph_frames):
                eager_frame(pending)
                pending = True
            main.append(("read_end",))
            wait()
            main.append(("read_begin", rk.tag))
            main.append(("read_end",))
            main.append(("advance_epoch", graph_frames))
            g_main, g_lead = main[g_main_start:], lead[g_lead_start:]
            if capture_fails and rk.r == 0:
                # rank 0's capture failed: its captured calls advanced the host-side tags, but none of their operations runs, and it
                # falls back to eager frames while the other ranks replay their graphs
                del main[g_main_start:], lead[g_lead_start:]
                pending = False
                for _ in range(graph_frames):
                    eager_frame(pending)
                    pending = True
                main.append(("read_end",))
                wait()
                main.append(("read_begin", rk.tag))
                main.append(("read_end",))
                programs.append({"main": main, "lead": lead})
                continue
            for rep in range(1, 3):  # two more replays of the same captured ops; event names get a replay suffix
                def ren(op, rep=rep):
                    if op[0] in ("record", "wait_event"):
                        return (op[0], op[1] + "#%d" % rep)
                    return op

                main.extend(ren(o) for o in g_main)
                lead.extend(ren(o) for o in g_lead)
        programs.append({"main": main, "lead": lead})
    return ranks, programs


def run(world, frames, graph_frames, seed, use_acks=True, capture_fails=False):
    ranks, programs = build_programs(world, frames, graph_frames, use_acks, capture_fails)
    rng = random.Random(seed)
    pc = [{s: 0 for s in p} for p in programs]
    events = [set() for _ in range(world)]
    steps = 0
    while True:
        runnable = []
        done = True
        for r in range(world):
            for s, ops in programs[r].items():
                i = pc[r][s]
                if i >= len(ops):
                    continue
                done = False
                op = ops[i]
                rk = ranks[r]
                ok = True
                if op[0] == "wait_event":
                    ok = op[1] in events[r]
                elif op[0] == "spin_acks":
                    ok = all(a >= op[1] + rk.epoch for a in rk.acks)
                elif op[0] == "spin_flags":
                    ok = all(f >= op[1] + rk.epoch for f in rk.flags)
                if ok:
                    runnable.append((r, s))
        if done:
            return steps
        assert runnable, "deadlock: %s" % [(r, s, programs[r][s][pc[r][s]]) for r in range(world) for s in programs[r] if pc[r][s] < len(programs[r][s])]
        r, s = rng.choice(runnable)
        op = programs[r][s][pc[r][s]]
        rk = ranks[r]
        if op[0] == "record":
            events[r].add(op[1])
        elif op[0] == "copy":
            dst, t = ranks[op[1]], op[2] + rk.epoch
            assert not dst.reading[t & 1], "rank %d overwrites buffer parity %d of rank %d while it is being read (frame %d)" % (r, t & 1, op[1], t)
            dst.buf[t & 1][r] = t
        elif op[0] == "raise_flags":
            for q in ranks:
                q.flags[r] = op[1] + rk.epoch
        elif op[0] == "raise_acks":
            for q in ranks:
                q.acks[r] = op[1] + rk.epoch
        elif op[0] == "read_begin":
            t = op[1] + rk.epoch
            rk.reading[t & 1] = True
            assert rk.buf[t & 1] == [t] * world, "rank %d reads frame %d but holds %s" % (r, t, rk.buf[t & 1])
        elif op[0] == "read_end":
            rk.reading = [False, False]
        elif op[0] == "advance_epoch":
            rk.epoch += op[1]
        pc[r][s] += 1
        steps += 1


@pytest.mark.parametrize("world", [2, 3, 8])
def test_eager_frames(world):
    for seed in range(60):
        run(world, frames=6, graph_frames=0, seed=seed)


@pytest.mark.parametrize("world", [2, 3, 8])
def test_graph_replay_after_eager_frames(world):
    """4 frames per graph (an even number: buffer parities are baked), three executions of the same ops, epoch + 4 after each"""
    for seed in range(40):
        run(world, frames=3, graph_frames=4, seed=seed)


def test_checker_catches_a_protocol_without_acknowledgements():
    caught = 0
    for seed in range(200):
        try:
            run(3, frames=6, graph_frames=0, seed=seed, use_acks=False)
        except AssertionError:
            caught += 1
    assert caught > 0


def test_one_rank_falling_back_after_a_failed_capture_breaks_the_exchange():
    """why bench.py makes every rank agree on the outcome of a capture before anything else happens, and refuses to go on with the
    exchange when a capture of it failed: the captured calls have advanced that rank's frame tags on the host without the frames
    having run, so a rank that falls back to eager frames speaks of other frames (and other buffer parities are not guaranteed)
    than the ranks that replay their graphs — the checker sees a deadlock or a consumer reading the wrong frame"""
    broken = 0
    for seed in range(20):
        try:
            run(3, frames=3, graph_frames=4, seed=seed, capture_fails=True)
        except AssertionError:
            broken += 1
    assert broken == 20
