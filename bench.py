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
    """dram bytes per step of kernel family `kernel` from the committed ncu --set full capture of THIS workload
    (profiles/r2_traffic.json, written by profiles/summarize_ncu.py from the capture named inside), else None."""
    path = os.path.join(ROOT, 'profiles', 'r2_traffic.json')
    if n_reads != 25_000_000 or not os.path.exists(path):
        return None
    try:
        return json.load(open(path)).get('dram_bytes_per_step', {}).get(kernel)
    except Exception:
        return None


def bind_to_gpu_numa_node(gpu):
    """Pin this process to the CPUs next to its GPU (NVML's ideal affinity) BEFORE pinned host memory is allocated, so
    first-touch places the staging buffers on the GPU's NUMA node: at 8 ranks the H2D copies otherwise cross the
    socket interconnect (round 1: 49 GB/s per GPU at N=1, 21 GB/s at N=8)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = [64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1]
        cpus = [c for c in cpus if c in os.sched_getaffinity(0)]
        if cpus:
            os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:
        return 0


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
    if kernel == 'k0_parse_pack':
        return text_bytes + 28 * n_reads          # the text once in, one packed record out
    if kernel == 'k2_rounds':
        return 130 * rounds_processed             # ~130 B per (unique sequence x round)
    return None


def tables_checksum(tabs):
    import hashlib
    h = hashlib.sha256()
    for k in ('mir_quant', 'mir_iscan', 'lib_totals', 'len_hist'):
        h.update(np.ascontiguousarray(tabs[k]).astype(np.int64).tobytes())
    return h.hexdigest()[:16]


EQ_CHUNKS, EQ_READS = 8, 50_000      # the fixed equality job: 8 samples of 50k reads, dealt round-robin to the ranks


def equality_job(world_obj, make_engine, rank, N, dev, reduce_fn, partition_fn):
    """A fixed small job whose reduced tables must not depend on N: 8 samples x 50k reads, sample s trimmed by rank
    s % N, hash partition, annotate, tally, device all-reduce.  Returns the checksum of the job-wide tables."""
    eng = make_engine(EQ_CHUNKS)
    for s in range(EQ_CHUNKS):
        if s % N == rank:
            eng.feed_fastq(s, world_obj.fastq(EQ_READS, seed=2000 + s, L=READ_LEN), 33)
    if N > 1:
        partition_fn(eng)
    eng.annotate()
    eng.tally()
    if N > 1:
        reduce_fn(eng)
    tabs = eng.export_tables()
    p, k, _ = eng.stats()
    assert int(p.sum()) == EQ_CHUNKS * EQ_READS, 'equality job: processed != reads fed'
    assert int(tabs['lib_totals'][:, 1:11].sum()) == int(k.sum()), 'equality job: library totals do not add up to kept reads'
    eng.sync()
    cs = tables_checksum(tabs)
    eng.close()
    return cs


def run_ours(args):
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world_size = int(os.environ.get('WORLD_SIZE', '1'))
    n_cpus_bound = bind_to_gpu_numa_node(local_rank)      # before torch allocates pinned memory
    import torch
    import torch.distributed as dist
    from mirge_b200.engine import Engine, LIB_IDS
    from mirge_b200 import dist as mdist

    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world_size > 1:
        if os.environ.get('NCCL_DEBUG', '').upper() in ('VERSION', 'WARN'):
            os.environ.pop('NCCL_DEBUG')           # NCCL prints its version banner on stdout: keep stdout to the one JSON line
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
    # total/N of them after it, so the single-GPU bound serves every N
    max_unique = args.max_unique or max(1 << 20, int(n_reads * 0.6))
    stream = torch.cuda.Stream(device=dev)

    def make_engine(n_samples):
        e = Engine(n_samples, False, max_unique if n_samples == S else (1 << 20), 64, device=local_rank)
        e.set_adapter(ADAPTER)
        load_libs(e, world, LIB_IDS)
        e.set_stream(stream.cuda_stream)
        return e

    eng = make_engine(S)
    if N == 1:
        eng.annotate_while_feeding(True)     # end-to-end mode: the rounds of region r run under the copy of region r+1
    sample = rank

    def step(resident):
        eng.reset()
        if resident:
            eng.feed_fastq_device(sample, dev_text.data_ptr(), nbytes, 33)
        else:
            eng.feed_fastq(sample, host.numpy(), 33)
        if N > 1:
            mdist.hash_partition(eng)                # X1
        eng.annotate()
        eng.tally()
        if N > 1:
            mdist.reduce_engine_tables(eng)          # X2, on the device, inside the timed region of both modes
        return eng.export_tables() if not resident else None     # D2H of the step's result

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
    # size-independent correctness properties at full size (tests/ hold the bit-exact oracle comparisons): after the
    # reduction every rank holds the job-wide tables
    p, k, U = eng.stats()
    slow_frac = eng.slow_reads() / float(n_reads)
    cls = eng.trim_classes()
    tabs = eng.export_tables()
    tot = tabs['lib_totals']
    assert int(p[sample]) == n_reads and int(p.sum()) == n_reads * N, 'processed != reads fed'
    assert int(tot[:, 1:11].sum()) == int(k.sum()), 'library totals do not add up to kept reads'
    assert int(tabs['len_hist'].sum()) == int(k.sum()), 'read-length histogram does not add up to kept reads'
    rounds_proc_all = tabs['round_processed']
    ms_e2e, _, _ = timed(False, args.steps, max(3, args.warmup))
    sampler.stop()
    d2h = sum(int(np.asarray(v).nbytes) for v in eng.export_tables().values())
    total_reads = n_reads * N
    value = total_reads * args.steps / (ms_res / 1e3)
    e2e = total_reads * args.steps / (ms_e2e / 1e3)
    # the fixed equality job: its checksum must be the same at every N (and equal the oracle's, checked at N=1 below)
    with torch.cuda.stream(stream):
        eq_sum = equality_job(world, make_engine, rank, N, dev, mdist.reduce_engine_tables, mdist.hash_partition)

    if rank == 0:
        peaks = {}
        pk_path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
        if os.path.exists(pk_path):
            peaks = json.load(open(pk_path))
        peak = float(peaks.get('hbm_gbs', 6650.0))
        per_step = {n: v[0] / args.steps for n, v in prof.items() if v[1] > 0}
        # kernel families: k1a..k1d are the launches of k1_trim_collapse; the mapping rounds are one family
        fam_ms = {'k0_parse_pack': per_step.get('k0_parse_pack', 0.0), 'k1_trim_collapse': per_step.get('k1_trim_collapse', 0.0),
                  'k2_rounds': sum(v for n, v in per_step.items() if n.startswith('k2_round')),
                  'k3_tally': sum(v for n, v in per_step.items() if n.startswith('k3_'))}
        rounds_proc = int(np.asarray(rounds_proc_all).sum())
        fam = {}
        for name, ms in fam_ms.items():
            ab = algorithmic_bytes(name, n_reads, int(k[sample]), U, nbytes, rounds_proc)
            fam[name] = {'ms_per_step': round(ms, 4), 'algorithmic_bytes_per_step': ab,
                         'achieved_gbs': (ab / (ms * 1e-3) / 1e9) if (ab and ms > 0) else None,
                         'frac': (ab / (ms * 1e-3) / 1e9 / peak) if (ab and ms > 0) else None,
                         'traffic': ncu_traffic(name, n_reads)}
        dom = max(fam_ms, key=lambda n: fam_ms[n])
        roof = {'bound': 'hbm', 'kernel': dom, 'achieved': fam[dom]['achieved_gbs'], 'peak': peak, 'unit': 'GB/s',
                'frac': fam[dom]['frac'], 'traffic': fam[dom]['traffic'],
                'algorithmic_bytes_per_launch': fam[dom]['algorithmic_bytes_per_step'],
                'launch_unit': 'one kernel family per step (k1_trim_collapse = its 3-4 launches per feed region; k2_rounds = the 9 round kernels)',
                'peak_source': 'measured (MEASURED_PEAKS.json)' if peaks else 'fallback (B200_PROFILING.md)',
                'families': fam, 'kernel_ms_per_step': {n: round(v, 4) for n, v in per_step.items()},
                'share_of_step': fam_ms[dom] / (ms_res / args.steps)}
        cpu = None
        if N == 1:
            sample_reads = min(n_reads, args.cpu_sample)
            cv, csec, cores = cpu_run(world, sample_reads, 1, 0, 1000)
            cpu = {'value': cv, 'unit': 'reads/s', 'cores': cores, 'kind': 'port',
                   'sample': 'first %d reads of the workload, one pass (trim+collapse+9 rounds+tally), oracle C port with OpenMP' % sample_reads,
                   'equality_checksum_matches_oracle': oracle_equality_checksum(world) == eq_sum}
        print(json.dumps({
            'metric': METRIC, 'value': value, 'unit': 'reads/s', 'n_gpus': N, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': ms_res / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'u8/u64 (2-bit packed integer)', 'data': 'synthetic',
            'config': {'workload': workload_name(args), 'reads_per_gpu': n_reads, 'unique_owned_rank0': U, 'kept_frac': float(k[sample]) / n_reads, 'k1_slow_path_frac': slow_frac, 'k1_class1_frac': cls[1] / float(n_reads), 'k1_class2_frac': cls[2] / float(n_reads), 'k0_regions_reparsed_exactly': cls[3],
                       'l2_policy': 'inputs (%.2f GB FASTQ text per GPU) larger than the 126 MB L2' % (nbytes / 1e9),
                       'parallelism': 'reads sharded by sample across %d GPU(s); hash-partition all-to-all (X1) + device table all-reduce (X2), both inside the timed region' % N,
                       'cpus_bound_to_gpu_numa_node': n_cpus_bound,
                       'equality_job': '%d samples x %d reads dealt round-robin to the ranks; checksum of the reduced tables must not depend on N' % (EQ_CHUNKS, EQ_READS),
                       'equality_checksum': eq_sum},
            'e2e': {'value': e2e, 'unit': 'reads/s', 'h2d_bytes_per_step': nbytes, 'd2h_bytes_per_step': d2h, 'ms_per_step': ms_e2e / args.steps},
            'gpu_launches': launches, 'clocks': sampler.summary(), 'roofline': roof, 'cpu_baseline': cpu,
        }))
    if world_size > 1:
        dist.destroy_process_group()


def oracle_equality_checksum(world_obj):
    """The equality job on the CPU oracle (cpu_baseline leg, N=1 only): the checker for the GPU checksum."""
    from oracle import oracle as O
    o = O.Oracle(EQ_CHUNKS, False, ADAPTER, threads=os.cpu_count() or 1)
    load_libs(o, world_obj, O.LIB_IDS)
    for s in range(EQ_CHUNKS):
        o.feed_fastq(s, world_obj.fastq(EQ_READS, seed=2000 + s, L=READ_LEN).tobytes(), 33)
    o.annotate()
    q, i, t = o.tally()
    return tables_checksum({'mir_quant': q, 'mir_iscan': i, 'lib_totals': t, 'len_hist': o.len_hist(64)})


def main():
    args = parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
