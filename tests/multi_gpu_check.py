"""Multi-GPU equality check (run under torchrun on N GPUs, not collected by pytest):
every rank trims+collapses its own sample, unique sequences are hash-partitioned with
an NCCL all-to-all, each rank annotates and tallies what it owns, tables are
all-reduced -- and must equal the single-process CPU oracle over all samples.

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


def main():
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    local = int(os.environ.get('LOCAL_RANK', rank))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    w = World(seed=20180626, scale=0.05)
    n_reads = 100000
    fq = w.fastq(n_reads, seed=1000 + rank)
    e = Engine(world, False, 1 << 20, 64, device=local)
    e.set_adapter(ADAPTER)
    for t, i in LIB_IDS.items():
        e.load_library(i, *w.libs[t])
    stream = torch.cuda.Stream(device=dev)
    e.set_stream(stream.cuda_stream)
    with torch.cuda.stream(stream):
        e.feed_fastq(rank, fq, 33)
        p, k, u_local = e.stats()
        n_owned = mdist.hash_partition(e)
        e.annotate()
        e.tally()
        t = e.export_tables()
        lh = t['len_hist'].astype(np.int64)
        red = mdist.reduce_tables({'mir_quant': t['mir_quant'], 'mir_iscan': t['mir_iscan'], 'lib_totals': t['lib_totals'],
                                   'len_hist': lh, 'processed': p.astype(np.int64), 'kept': k.astype(np.int64),
                                   'n_unique': np.array([n_owned], np.int64)}, dev)
    if rank == 0:
        from oracle import oracle as O
        o = O.Oracle(world, False, ADAPTER)
        for tname, i in O.LIB_IDS.items():
            o.load_library(i, *w.libs[tname])
        for r in range(world):
            o.feed_fastq(r, w.fastq(n_reads, seed=1000 + r).tobytes(), 33)
        o.annotate()
        q, isc, tot = o.tally()
        op, ok = o.stats()
        assert np.array_equal(red['mir_quant'], q), 'mir_quant differs'
        assert np.array_equal(red['mir_iscan'], isc), 'mir_iscan differs'
        assert np.array_equal(red['lib_totals'], tot), 'lib_totals differs'
        assert np.array_equal(red['len_hist'], o.len_hist(64).astype(np.int64)), 'len_hist differs'
        assert red['processed'].tolist() == [int(x) for x in op] and red['kept'].tolist() == [int(x) for x in ok]
        assert int(red['n_unique'][0]) == int(O.lib().orc_n_unique(o.h)), 'owned uniques do not add up to the global unique count'
        print('multi-GPU check ok: %d ranks, %d unique sequences, tables identical to the oracle' % (world, int(red['n_unique'][0])))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
