"""The CLI path end to end: `python -m mirge_b200 annotate -s <file.fastq>` on an uncompressed FASTQ in
tmpfs, timed from process start of main() to the last table written -- file reading (pinned double buffers filled by a
thread pool), H2D, kernels, export of every unique sequence, the Python table writers (mapped.csv / unmapped.csv dominate).
    python profiles/bench_cli.py [reads]        (default 25 M reads = 2.95 GB of text)
"""
import json
import os
import shutil
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from synth.world import World
from mirge_b200.__main__ import main

n = int(sys.argv[1]) if len(sys.argv) > 1 else 25_000_000
w = World(seed=20180626, scale=1.0)
root = tempfile.mkdtemp(dir='/dev/shm' if os.path.isdir('/dev/shm') else None)
try:
    w.write(os.path.join(root, 'libs'))
    fq = os.path.join(root, 'sample.fastq')
    with open(fq, 'wb') as f:
        step = 5_000_000
        for a in range(0, n, step):
            f.write(w.fastq(min(step, n - a), seed=1000, first_read=a).tobytes())
    size = os.path.getsize(fq)
    stages = {}
    import mirge_b200.pipeline as P
    orig_run, orig_write = P.AnnotateRun.run, P.AnnotateRun.write_outputs

    def run(self):
        t = time.perf_counter()
        r = orig_run(self)
        stages['trim_collapse_annotate_tally_s'] = time.perf_counter() - t
        return r

    def write_outputs(self, *a, **k):
        t = time.perf_counter()
        r = orig_write(self, *a, **k)
        stages['export_and_tables_s'] = time.perf_counter() - t
        return r
    orig_new, orig_trim = P.AnnotateRun._new_engine, P.AnnotateRun.trim_file

    def new_engine(self):
        t = time.perf_counter()
        r = orig_new(self)
        self.engine.sync()
        stages['create_engine_load_libraries_s'] = stages.get('create_engine_load_libraries_s', 0.0) + time.perf_counter() - t
        return r

    def trim_file(self, *a, **k):
        t = time.perf_counter()
        r = orig_trim(self, *a, **k)
        stages['read_file_feed_s'] = stages.get('read_file_feed_s', 0.0) + time.perf_counter() - t
        return r
    P.AnnotateRun._new_engine, P.AnnotateRun.trim_file = new_engine, trim_file
    P.AnnotateRun.run, P.AnnotateRun.write_outputs = run, write_outputs
    t0 = time.perf_counter()
    out = main(['annotate', '-s', fq, '-o', root, '-lib', os.path.join(root, 'libs'), '-sp', 'synth', '-ad', 'illumina'])
    total = time.perf_counter() - t0
    print(json.dumps({'config': 'CLI end to end', 'reads': n, 'fastq_bytes': size, 'total_s': total, 'reads_per_s': n / total,
                      'device_stage_reads_per_s': n / stages['trim_collapse_annotate_tally_s'], **stages,
                      'outputs': sorted(os.listdir(out))}))
finally:
    shutil.rmtree(root, ignore_errors=True)
