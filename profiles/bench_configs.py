"""BASELINE.json configs other than the bench.py headline (C2), on ONE B200, through the Engine API.  One JSON object per
config on stdout.  Run under gpurun:  python profiles/bench_configs.py [c1,c3,c4,c5]

  C1  1 M reads, 1 sample                          resident + end-to-end (pinned host text)
  C3  8 samples x 25 M reads batched on one GPU    end-to-end, S = 8 count columns (the N-GPU runs are bench.py --gpus N)
  C4  2 x 25 M reads with -di -gff -ai             device part, then the host tables (isomiR, GFF rows, A-to-I report)
  C5  C3's table + the un-annotated-read genome pass (`-n 0 -m 3 -l 25 -a --best`, __main__.py:538) vs a 10^8-nt genome
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from synth.world import World
from mirge_b200.engine import Engine, LIB_IDS

ADAPTER = 'TGGAATTCTCGGGTGCCAAGGAACTCCAG'
which = (sys.argv[1] if len(sys.argv) > 1 else 'c1,c3,c4,c5').split(',')
world = World(seed=20180626, scale=1.0)
rec = world.record_bytes(50)


def engine(S, max_unique):
    e = Engine(S, False, max_unique, 64, device=0)
    e.set_adapter(ADAPTER)
    for t, i in LIB_IDS.items():
        e.load_library(i, *world.libs[t])
    return e


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


def host_text(n, seed):
    h = torch.empty(n * rec, dtype=torch.uint8, pin_memory=True)
    world.fastq(n, seed=seed, L=50, out=h.numpy())
    return h


if 'c1' in which:
    n = 1_000_000
    h = host_text(n, 1000)
    d = h.cuda()
    e = engine(1, 1 << 20)

    def run(resident):
        e.reset()
        if resident:
            e.feed_fastq_device(0, d.data_ptr(), n * rec, 33)
        else:
            e.feed_fastq(0, h.numpy(), 33)
        e.annotate(); e.tally()
        return e.export_tables()
    tr, te = timed(lambda: run(True), 10), timed(lambda: run(False), 10)
    p, k, U = e.stats()
    print(json.dumps({'config': 'C1', 'reads': n, 'unique': U, 'resident_reads_per_s': n / tr, 'e2e_reads_per_s': n / te,
                      'resident_ms': tr * 1e3, 'e2e_ms': te * 1e3}), flush=True)
    e.close()

if 'c3' in which or 'c5' in which or 'c4' in which:
    n = 25_000_000
    S = 8 if ('c3' in which or 'c5' in which) else 2
    texts = [host_text(n, 1000 + s) for s in range(S)]

if 'c3' in which or 'c5' in which:
    e = engine(8, 48_000_000)

    def run8():
        e.reset()
        for s in range(8):
            e.feed_fastq(s, texts[s].numpy(), 33)
        e.annotate(); e.tally()
        return e.export_tables()
    t8 = timed(run8, 2)
    p, k, U = e.stats()
    tabs = e.export_tables()
    assert int(tabs['lib_totals'][:, 1:11].sum()) == int(k.sum())
    print(json.dumps({'config': 'C3 on one GPU', 'reads': 8 * n, 'samples': 8, 'unique': U, 'e2e_reads_per_s': 8 * n / t8, 'e2e_ms': t8 * 1e3,
                      'h2d_gb': 8 * n * rec / 1e9}), flush=True)
    if 'c5' in which:
        t0 = time.perf_counter()
        gn, gs = world.genome(total=100_000_000, n_chrom=8)
        t_gen = time.perf_counter() - t0
        t0 = time.perf_counter()
        e.load_genome(gn, gs)
        e.sync()
        t_idx = time.perf_counter() - t0
        u = e.export_unique(with_seq=False)
        tot = u['counts'].sum(axis=1)
        ids = np.nonzero((u['round_of'] == 0) & (u['len'] >= 16) & (u['len'] <= 25) & (tot >= 2))[0].astype(np.uint32)   # convert2Fasta.py:124,179
        t0 = time.perf_counter()
        c = e.genome_pass_ids(ids, 0, 0, 25, 0)
        t_pass = time.perf_counter() - t0
        hits = c.sum(axis=1)
        print(json.dumps({'config': 'C5 genome pass', 'genome_nt': sum(len(x) for x in gs), 'index_build_s': t_idx, 'queries': int(len(ids)),
                          'pass_ms': t_pass * 1e3, 'queries_per_s': len(ids) / t_pass, 'aligned': int((hits > 0).sum()),
                          'suppressed_by_m3': int((hits > 3).sum()), 'genome_synth_s': t_gen,
                          'c5_total_reads_per_s': 8 * n / (t8 + t_pass)}), flush=True)
    e.close()

if 'c4' in which:
    import tempfile
    from mirge_b200 import tables, gff as gff_mod, a2i
    e = engine(2, 24_000_000)
    t0 = time.perf_counter()
    for s in range(2):
        e.feed_fastq(s, texts[s].numpy(), 33)
    e.annotate(); e.tally()
    tabs = e.export_tables()
    e.sync()
    t_dev = time.perf_counter() - t0
    t0 = time.perf_counter()
    uniq = e.export_unique()
    t_exp = time.perf_counter() - t0
    print(json.dumps({'config': 'C4 device part', 'reads': 2 * n, 'unique': int(len(uniq['len'])), 'device_e2e_ms': t_dev * 1e3,
                      'reads_per_s': 2 * n / t_dev, 'export_unique_s': t_exp}), flush=True)
    # the -ai genome passes on every miRNA-annotated unique sequence (writeDataToCSV.py:1260: -n 1 -a -3 2) and the
    # -n 0 variant (:1488), against a 10^8-nt genome
    gn, gs = world.genome(total=100_000_000, n_chrom=8)
    t0 = time.perf_counter()
    e.load_genome(gn, gs)
    e.sync()
    t_idx = time.perf_counter() - t0
    ids = np.nonzero((uniq['round_of'] == 1) | (uniq['round_of'] == 9))[0].astype(np.uint32)
    out = {}
    for label, nmm in (('n1', 1), ('n0', 0)):
        e.genome_pass_ids(ids[:1000], 0, nmm, 28, 2)
        t0 = time.perf_counter()
        c = e.genome_pass_ids(ids, 0, nmm, 28, 2)
        dt = time.perf_counter() - t0
        occ = c.sum(axis=1)
        out[label] = {'pass_ms': dt * 1e3, 'queries_per_s': len(ids) / dt, 'aligned': int((occ > 0).sum()), 'alignments': int(c.sum())}
    best_unique = int(((c[:, 0] == 1) | ((c[:, 0] == 0) & (c[:, 1] == 1))).sum())
    print(json.dumps({'config': 'C4 genome passes (-ai)', 'genome_nt': sum(len(x) for x in gs), 'index_build_s': t_idx, 'queries': int(len(ids)),
                      **{k + '_' + kk: vv for k, v in out.items() for kk, vv in v.items()}}), flush=True)
    e.close()
