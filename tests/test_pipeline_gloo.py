"""N > 1 host-side logic on CPU: world_size-2 `gloo` process group: stage ranges partition the layers the
way the reference's split does, the NCCL id reaches every rank, timings reduce to the slowest rank."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from booster_b200 import pipeline as P


def test_split_matches_reference_rule():
    # equal proportions over 32 layers / 8 stages: contiguous blocks, every layer owned once (llama.cpp:5932-5968)
    dev = P.split_layers(32, [1] * 8)
    assert dev == sorted(dev) and set(dev) == set(range(8))
    # il / (n_layer + 1) against the normalised cumulative split: stage 0 gets ceil(33/8) = 5 layers, the last one 3
    assert [dev.count(d) for d in range(8)] == [5, 4, 4, 4, 4, 4, 4, 3]
    # gpu1..gpu4 style percentages (server.go:514-530)
    dev = P.split_layers(80, [50, 25, 25, 0])
    assert dev.count(0) == 41 and dev.count(1) == 20 and dev.count(2) == 19 and dev.count(3) == 0
    assert P.stage_range(80, 3, 4, [50, 25, 25, 0]) == (0, 0)
    with pytest.raises(ValueError):
        P.split_layers(4, [0, 0])


def test_equal_split_partitions():
    for L in (2, 7, 32, 80):
        for world in (1, 2, 4, 8):
            r = [P.stage_range(L, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == L
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        uid = P.share_unique_id(dist, lambda: bytes(range(128)))
        lb, le = P.stage_range(32, rank, world)
        slowest = P.max_over_ranks(dist, 1.0 + rank)
        launches = P.sum_over_ranks(dist, 100 * (rank + 1))
        # the residual-stream hand-off, stood in for by gloo send/recv of f32[n_embd] (NCCL does it on the GPU box)
        x = torch.full((4096,), float(rank), dtype=torch.float32)
        if rank == 0:
            dist.send(x, dst=1)
        else:
            dist.recv(x, src=0)
        # inbox handles of the direct NVLink hand-off: every rank learns its successor's and rank 0's
        nxt, first = P.exchange_inbox_handles(dist, bytes([rank]) * 64)

        # the peer hand-off is taken by every rank or by none: rank 1 cannot map its peer here -> both fall back to NCCL
        class FakeCtx:
            def __init__(self, fail):
                self.fail, self.connected, self.disabled = fail, False, False

            def p2p_handle(self):
                return bytes([rank]) * 64

            def p2p_connect(self, r, w, n, f):
                if self.fail:
                    raise RuntimeError("cudaIpcOpenMemHandle failed")
                self.connected = True

            def p2p_disable(self):
                self.disabled = True

        bad, good = FakeCtx(rank == 1), FakeCtx(False)
        decided = (P.connect_peer_handoff(dist, bad, rank, world), bad.disabled, P.connect_peer_handoff(dist, good, rank, world), good.disabled)
        dist.barrier()
        q.put((rank, uid, lb, le, slowest, launches, float(x[0]), nxt[0], first[0], decided))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_world2_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=100) for _ in range(world))
    for p in procs:
        p.join(30)
        assert p.exitcode == 0
    assert [r[1] for r in res] == [bytes(range(128))] * 2          # rank 0's id reached rank 1
    assert [(r[2], r[3]) for r in res] == [(0, 16), (16, 32)]      # stages partition the 32 layers
    assert all(r[4] == 2.0 for r in res)                           # slowest rank's time everywhere
    assert all(r[5] == 300 for r in res)
    assert res[1][6] == 0.0                                        # stage 1 received stage 0's stream
    assert [(r[7], r[8]) for r in res] == [(1, 0), (1, 0)]         # next rank's handle (the last rank: its own), rank 0's handle
    assert all(r[9] == (False, True, True, False) for r in res)    # one rank's failure disables the peer hand-off on both
