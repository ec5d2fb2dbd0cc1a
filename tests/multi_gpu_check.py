"""Multi-GPU equality check, run under torchrun on N GPUs (tests/test_gpu_multi.py launches it when >= 2 devices are
visible):

  A. one sample per rank (what bench.py --gpus N does): every rank trims+collapses its own sample, unique sequences are
     hash-partitioned with an NCCL all-to-all (X1), each rank annotates and tallies what it owns, the integer tables are
     summed on the device with one all-reduce (X2) -- and must equal the single-process CPU oracle over all samples;
  B. ONE sample split into record-aligned file chunks over the ranks (north_star: "reads shard by file chunk"): same
     check, plus the owned unique sets of the ranks are disjoint and their union is the oracle's set with its counts
     and per-sequence round / reference / offset.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/multi_gpu_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from synth.world import World                         # noqa: E402
from mirge_b200.engine import Engine, LIB_IDS         # noqa: E402
from mirge_b200 import dist as mdist                  # noqa: E402

ADAPTER = 'TGGAATTCTCGGGTGCCAAGGAACTCCAG'


def run_job(w, world, rank, dev, n_samples, my_chunks):
    """my_chunks: list of (sample index, FASTQ bytes) this rank trims."""
    e = Engine(n_samples, False, 1 << 20, 64, device=dev.index)
    e.set_adapter(ADAPTER)
    for t, i in LIB_IDS.items():
        e.load_library(i, *w.libs[t])
    stream = torch.cuda.Stream(device=dev)
    e.set_stream(stream.cuda_stream)
    with torch.cuda.stream(stream):
        for s, fq in my_chunks:
            e.feed_fastq(s, fq, 33)
        mdist.hash_partition(e)
        e.annotate()
        e.tally()
        mdist.reduce_engine_tables(e)
        t = e.export_tables()
        p, k, n_owned = e.stats()
        u = e.export_unique()
    return e, t, p, k, n_owned, u


def oracle_job(w, n_samples, chunks):
    from oracle import oracle as O
    o = O.Oracle(n_samples, False, ADAPTER)
    for tname, i in O.LIB_IDS.items():
        o.load_library(i, *w.libs[tname])
    for s, fq in chunks:
        o.feed_fastq(s, fq, 33)
    o.annotate()
    return O, o


def check_tables(O, o, t, p, k, n_total, tag):
    q, isc, tot = o.tally()
    op, ok = o.stats()
    assert np.array_equal(t['mir_quant'], q), tag + ': mir_quant differs'
    assert np.array_equal(t['mir_iscan'], isc), tag + ': mir_iscan differs'
    assert np.array_equal(t['lib_totals'], tot), tag + ': lib_totals differs'
    assert np.array_equal(t['len_hist'].astype(np.int64), o.len_hist(64).astype(np.int64)), tag + ': len_hist differs'
    assert [int(x) for x in p] == [int(x) for x in op] and [int(x) for x in k] == [int(x) for x in ok], tag + ': processed/kept differ'
    assert n_total == int(O.lib().orc_n_unique(o.h)), tag + ': owned uniques do not add up to the global unique count'


def main():
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    local = int(os.environ.get('LOCAL_RANK', rank))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    w = World(seed=20180626, scale=0.05)

    # ---- A: one sample per rank
    n_reads = 100000
    fqs = [w.fastq(n_reads, seed=1000 + r).tobytes() for r in range(world)]
    e, t, p, k, n_owned, u = run_job(w, world, rank, dev, world, [(rank, fqs[rank])])
    tot = torch.tensor([n_owned], dtype=torch.int64, device=dev)
    dist.all_reduce(tot)
    if rank == 0:
        O, o = oracle_job(w, world, list(enumerate(fqs)))
        check_tables(O, o, t, p, k, int(tot.item()), 'A')
        print('multi-GPU check A ok: %d ranks x 1 sample, %d unique sequences, tables identical to the oracle' % (world, int(tot.item())))
    e.close()

    # ---- B: one sample, record-aligned file chunks over the ranks
    fq = w.fastq(150000, seed=77).tobytes()
    rec = w.record_bytes(50)
    n_rec = len(fq) // rec
    bounds = [(n_rec * r // world) * rec for r in range(world + 1)]
    e, t, p, k, n_owned, u = run_job(w, world, rank, dev, 1, [(0, fq[bounds[rank]:bounds[rank + 1]])])
    mine = {s.decode(): (int(c[0]), int(ro), int(rf) if ro else 0, int(ps) if ro else 0)
            for s, c, ro, rf, ps in zip(u['seq'], u['counts'], u['round_of'], u['ref_of'], u['pos_of'])}
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    if rank == 0:
        O, o = oracle_job(w, 1, [(0, fq)])
        check_tables(O, o, t, p, k, sum(len(g) for g in gathered), 'B')
        union = {}
        for g in gathered:
            assert not (set(g) & set(union)), 'B: a sequence is owned by two ranks'
            union.update(g)
        oe = o.export()
        want = {s: (int(c[0]), int(ro), int(rf) if ro else 0, int(ps) if ro else 0)
                for s, c, ro, rf, ps in zip(oe['seqs'], oe['counts'], oe['round_of'], oe['ref_of'], oe['pos_of'])}
        assert union == want, 'B: per-sequence counts / assignments differ from the oracle'
        print('multi-GPU check B ok: 1 sample in %d file chunks, %d unique sequences, per-sequence results identical to the oracle'
              % (world, len(union)))
    e.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
