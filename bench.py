#!/usr/bin/env python
"""bench.py -- raw reads/s trimmed + collapsed + annotated (BASELINE.json metric).

One "step" = one complete annotate job over one synthetic sample per GPU:
FASTQ text -> K0 parse/pack -> K1 trim+collapse -> (N>1: hash-partition
all-to-all) -> K2 mapping rounds -> K3 tallies -> (N>1: table all-reduce).

    python bench.py --gpus 1 --steps 3 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference          # the CPU arm (oracle port, all host cores)

`value` is timed with the FASTQ text already resident in HBM; `e2e` feeds the
same text from pinned host memory through the public API (H2D inside the timed
region) and reads the count tables back every step.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

ADAPTER = 'TGGAATTCTCGGGTGCCAAGGAACTCCAG'     # -ad illumina, src/mirge/__main__.py:124-125
METRIC = 'raw reads/sec trimmed+collapsed+annotated'
READ_LEN = 50


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--reads', type=int, default=25_000_000, help='raw reads per GPU per step (BASELINE configs[1]: 25M)')
    ap.add_argument('--scale', type=float, default=1.0, help='library scale (1.0 = human-like sizes, SURVEY 8d)')
    ap.add_argument('--cpu-sample', type=int, default=2_000_000, help='reads in the CPU-baseline sample')
    ap.add_argument('--max-unique', type=int, default=0)
    return ap.parse_args()


def workload_name(args):
    return ('single sample %dM reads/GPU (L=%d, illumina adapter), mature+hairpin+tRNA/rRNA/snoRNA/ncRNA/mRNA+isomiR rounds, '
            'synthetic libs scale %.2f' % (args.reads // 1_000_000, READ_LEN, args.scale))


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled while the timed regions run (B200_PROFILING.md)."""
    Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu):
        super().__init__(daemon=True)
        self.gpu, self.rows, self.stop_flag = gpu, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(['nvidia-smi', '-i', str(self.gpu), '--query-gpu=' + self.Q, '--format=csv,noheader,nounits'],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(',')])
            except Exception:
                pass
            time.sleep(0.01)

    def stop(self):
        self.stop_flag = True
        self.join(timeout=6)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r and r[0].replace('.', '').isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace('.', '').isdigit()]
        reasons = set()
        for r in self.rows:
            for name, v in zip(['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'], r[2:6]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(self.rows)}


def ncu_traffic(kernel, n_reads):
    """dram bytes per launch of `kernel` from the committed ncu --set full capture of this workload, else None."""
    path = os.path.join(ROOT, 'profiles', 'r1_k1_full_metrics.txt')
    if kernel != 'k1_trim_collapse' or n_reads != 25_000_000 or not os.path.exists(path):
        return None
    tot = 0.0
    for line in open(path):
        f = line.split()
        if len(f) >= 3 and f[0] in ('dram__bytes_read.sum', 'dram__bytes_write.sum'):
            tot += float(f[1]) * {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0}[f[2]]
    return tot or None


def build_world(args):
    from synth.world import World
    return World(seed=20180626, scale=args.scale)


def load_libs(target, world, ids):
    for t, i in ids.items():
        target.load_library(i, *world.libs[t])


# ---------------------------------------------------------------------------
# CPU arm: the oracle port of the reference pipeline on all host cores
# ---------------------------------------------------------------------------
def cpu_run(world, n_reads, steps, warmup, seed):
    from oracle import oracle as O
    cores = os.cpu_count() or 1
    o = O.Oracle(1, False, ADAPTER, threads=cores)
    load_libs(o, world, O.LIB_IDS)        # index build = bowtie-build, not timed (the reference ships prebuilt indexes)
    fq = world.fastq(n_reads, seed=seed, L=READ_LEN).tobytes()
    times = []
    for it in range(warmup + steps):
        o.reset()
        t0 = time.perf_counter()
        o.feed_fastq(0, fq, 33)
        o.annotate()
        o.tally()
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    sec = float(np.mean(times))
    return n_reads / sec, sec, cores


def probe_real_reference():
    """is the real miRge2.0 stack (python2 + cutadapt + bowtie 1.x) on this box?
    It never is in this image; when it is, the oracle should be diffed against it first."""
    found = {}
    for name, cmd in (('python2+cutadapt', ['python2', '-c', 'import cutadapt; print(cutadapt.__version__)']),
                      ('bowtie', ['bowtie', '--version']), ('bowtie-build', ['bowtie-build', '--version'])):
        try:
            r = subprocess.run(cmd, capture_output=True, text=True, timeout=20)
            found[name] = (r.stdout.strip().splitlines() or [''])[0][:60] if r.returncode == 0 else None
        except Exception:
            found[name] = None
    return found


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    probe = probe_real_reference()
    if all(probe.values()):
        sys.stderr.write('bench: python2+cutadapt+bowtie found (%s) -- the real miRge2.0 could run here; this arm still '
                         'times the oracle port (diff the oracle against the real tools first, SURVEY.md 8c)\n' % probe)
    world = build_world(args)
    sample = min(args.reads, args.cpu_sample)
    value, sec, cores = cpu_run(world, sample, args.steps, args.warmup, 1000)
    absent = ', '.join(k for k, v in probe.items() if not v) or 'none'
    desc = 'first %d reads of the workload per step; oracle C port (real miRge2.0 needs python2+cutadapt+bowtie; probed, absent: %s)' % (sample, absent)
    print(json.dumps({
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': 'reads/s', 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': sec * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'u8/u64 (2-bit packed integer)', 'data': 'synthetic', 'config': {'workload': workload_name(args), 'sample_reads': sample},
        'cpu_baseline': {'value': value, 'unit': 'reads/s', 'cores': cores, 'kind': 'port', 'sample': desc},
        'e2e': {'value': value, 'unit': 'reads/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }))


# ---------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------
def algorithmic_bytes(kernel, n_reads, kept, n_unique, text_bytes, rounds_processed):
    """Algorithmic HBM bytes per step for one kernel family."""
    if kernel == 'k1_trim_collapse':
        return 28 * n_reads + 40 * kept + 60 * n_unique
    if kernel == 'k0_pack':
        return text_bytes + 28 * n_reads + 20 * n_reads
    if kernel in ('k0_count_newlines', 'k0_line_starts'):
        return text_bytes + (16 * n_reads if kernel == 'k0_line_starts' else 0)
    if kernel.startswith('k2_round'):
        return 130 * rounds_processed
    return None


def run_ours(args):
    import torch
    import torch.distributed as dist
    from mirge_b200.engine import Engine, LIB_IDS
    from mirge_b200 import dist as mdist

    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world_size = int(os.environ.get('WORLD_SIZE', '1'))
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world_size > 1:
        if os.environ.get('NCCL_DEBUG', 'VERSION').upper() == 'VERSION':
            os.environ['NCCL_DEBUG'] = 'WARN'      # keep stdout to the one JSON line
        dist.init_process_group('nccl', device_id=dev)
    N, S = world_size, world_size            # one sample per rank (weak scaling)
    world = build_world(args)
    n_reads = args.reads
    rec = world.record_bytes(READ_LEN)
    nbytes = n_reads * rec
    host = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    world.fastq(n_reads, seed=1000 + rank, L=READ_LEN, out=host.numpy())
    dev_text = host.to(dev)
    # a rank holds its own sample's uniques before the exchange (0.32 per read on this workload) and at most
    # total/N of them after it, so the single-GPU bound serves every N (a larger table only costs memset time)
    max_unique = args.max_unique or max(1 << 20, int(n_reads * 0.6))
    eng = Engine(S, False, max_unique, 64, device=local_rank)
    eng.set_adapter(ADAPTER)
    load_libs(eng, world, LIB_IDS)
    stream = torch.cuda.Stream(device=dev)
    eng.set_stream(stream.cuda_stream)
    sample = rank

    def step(resident):
        eng.reset()
        if resident:
            eng.feed_fastq_device(sample, dev_text.data_ptr(), nbytes, 33)
        else:
            eng.feed_fastq(sample, host.numpy(), 33)
        if N > 1:
            mdist.hash_partition(eng)
        eng.annotate()
        eng.tally()
        tables = None
        if not resident:
            tables = eng.export_tables()     # D2H of the step's result
            if N > 1:
                tables = mdist.reduce_tables({k: tables[k] for k in ('mir_quant', 'mir_iscan', 'lib_totals')}, dev)
        return tables

    def timed(resident, steps, warmup, profile=False):
        with torch.cuda.stream(stream):
            for _ in range(warmup):
                step(resident)
            eng.sync()
            if profile:
                eng.profile_enable(True)
            l0 = eng.launch_count()
            if N > 1:
                dist.barrier()
            torch.cuda.synchronize()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record(stream)
            for _ in range(steps):
                step(resident)
            ev1.record(stream)
            eng.sync()
            torch.cuda.synchronize()
            ms = ev0.elapsed_time(ev1)
            if N > 1:
                dist.barrier()
                t = torch.tensor([ms], dtype=torch.float64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = float(t.item())
            prof = eng.profile_read() if profile else None
            if profile:
                eng.profile_enable(False)
            return ms, eng.launch_count() - l0, prof

    sampler = ClockSampler(local_rank)
    sampler.start()
    ms_res, launches, prof = timed(True, args.steps, args.warmup, profile=True)
    # size-independent correctness properties at full size (tests/ hold the bit-exact oracle comparisons)
    p, k, U = eng.stats()
    slow_frac = eng.slow_reads() / float(n_reads)
    cls = eng.trim_classes()
    tabs = eng.export_tables()
    if N > 1:
        tabs = mdist.reduce_tables({kk: tabs[kk] for kk in ('lib_totals',)}, dev)
    tot = tabs['lib_totals']
    assert int(p[sample]) == n_reads, 'processed != reads fed'
    if N == 1:
        assert int(tot[sample, 1:11].sum()) == int(k[sample]), 'library totals do not add up to kept reads'
    ms_e2e, _, _ = timed(False, args.steps, max(3, args.warmup))
    sampler.stop()
    d2h = sum(int(np.asarray(v).nbytes) for v in eng.export_tables().values())
    total_reads = n_reads * N
    value = total_reads * args.steps / (ms_res / 1e3)
    e2e = total_reads * args.steps / (ms_e2e / 1e3)

    roof = None
    cpu = None
    if rank == 0:
        peaks = {}
        pk_path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
        if os.path.exists(pk_path):
            peaks = json.load(open(pk_path))
        peak = float(peaks.get('hbm_gbs', 6650.0))
        fam = {n: v for n, v in prof.items() if n.startswith('k') and v[1] > 0 and not n.startswith(('k1a', 'k1b', 'k1c', 'k1d'))}   # k1a..k1d are the four launches of the k1_trim_collapse family
        dom = max(fam, key=lambda n: fam[n][0])
        dom_ms, dom_n = fam[dom]
        rounds_proc = int(eng.export_tables()['round_processed'][int(dom[len('k2_round')])]) if dom.startswith('k2_round') else 0
        ab = algorithmic_bytes(dom, n_reads, int(k[sample]), U, nbytes, rounds_proc)
        per_launch_bytes = ab * args.steps / dom_n if ab else None
        per_launch_ms = dom_ms / dom_n
        ach = per_launch_bytes / (per_launch_ms * 1e-3) / 1e9 if ab else None
        roof = {'bound': 'hbm', 'kernel': dom, 'achieved': ach, 'peak': peak, 'unit': 'GB/s',
                'frac': (ach / peak) if ach else None, 'traffic': ncu_traffic(dom, n_reads),
                'algorithmic_bytes_per_launch': per_launch_bytes,
                'peak_source': 'measured (MEASURED_PEAKS.json)' if peaks else 'fallback (B200_PROFILING.md)',
                'kernel_ms_per_step': {n: round(v[0] / args.steps, 4) for n, v in prof.items() if v[1] > 0},
                'share_of_step': dom_ms / ms_res}
        sample_reads = min(n_reads, args.cpu_sample)
        cv, csec, cores = cpu_run(world, sample_reads, 1, 0, 1000)
        cpu = {'value': cv, 'unit': 'reads/s', 'cores': cores, 'kind': 'port',
               'sample': 'first %d reads of the workload, one pass (trim+collapse+9 rounds+tally), oracle C port with OpenMP' % sample_reads}
        print(json.dumps({
            'metric': METRIC, 'value': value, 'unit': 'reads/s', 'n_gpus': N, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': ms_res / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'u8/u64 (2-bit packed integer)', 'data': 'synthetic',
            'config': {'workload': workload_name(args), 'reads_per_gpu': n_reads, 'unique_per_gpu': U, 'kept_frac': float(k[sample]) / n_reads, 'k1_slow_path_frac': slow_frac, 'k1_class1_frac': cls[1] / float(n_reads), 'k1_class2_frac': cls[2] / float(n_reads),
                       'l2_policy': 'inputs (%.2f GB FASTQ text per GPU) larger than the 126 MB L2' % (nbytes / 1e9),
                       'parallelism': 'reads sharded by sample across %d GPU(s); hash-partition all-to-all + table all-reduce' % N},
            'e2e': {'value': e2e, 'unit': 'reads/s', 'h2d_bytes_per_step': nbytes, 'd2h_bytes_per_step': d2h, 'ms_per_step': ms_e2e / args.steps},
            'gpu_launches': launches, 'clocks': sampler.summary(), 'roofline': roof, 'cpu_baseline': cpu,
        }))
    if world_size > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
