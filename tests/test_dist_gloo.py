"""world_size-2 gloo test of the multi-GPU plumbing (hash-partition all-to-all of
unique-sequence records, table all-reduce) on CPU tensors."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

RB = 40   # record bytes used by the fake records below


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from mirge_b200 import dist as mdist
    rng = np.random.default_rng(100 + rank)
    # records: first byte = destination rank, byte 1 = source rank, rest = payload; grouped by destination
    counts = [int(rng.integers(0, 50)) for _ in range(world)]
    recs = []
    for d, c in enumerate(counts):
        for _ in range(c):
            r = rng.integers(0, 255, RB, dtype=np.uint8)
            r[0], r[1] = d, rank
            recs.append(r)
    send = torch.from_numpy(np.concatenate(recs) if recs else np.zeros(0, np.uint8))
    recv, n = mdist.exchange_records(send, counts, RB)
    got = recv[:n * RB].numpy().reshape(n, RB)
    ok = bool((got[:, 0] == rank).all())
    tabs = mdist.reduce_tables({'t': np.full((3, 2), rank + 1, np.int64)}, torch.device('cpu'))
    q.put((rank, ok, n, counts, int(got[:, 1].astype(np.int64).sum()), tabs['t'].tolist()))
    dist.destroy_process_group()


def test_exchange_and_reduce_world2():
    world, port = 2, 29517
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(60)
    sent = {r[0]: r[3] for r in res}
    for rank, ok, n, counts, src_sum, tab in res:
        assert ok
        assert n == sum(sent[s][rank] for s in range(world))          # everything addressed to me arrived
        assert src_sum == sum(s * sent[s][rank] for s in range(world))
        assert tab == [[3, 3]] * 3                                       # 1 + 2 summed over ranks
