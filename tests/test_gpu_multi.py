"""Multi-GPU equality (collected by pytest): runs tests/multi_gpu_check.py under torch.distributed.run on every visible
device when there are at least two; its output is kept in gpurun_out/ as evidence."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_n_rank_tables_equal_the_oracle():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip('needs at least two GPUs (gpurun --gpus 2)')
    n = min(n, 8)
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(n), '--master-addr', '127.0.0.1',
           '--master-port', '29533', os.path.join(ROOT, 'tests', 'multi_gpu_check.py')]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    with open(os.path.join(ROOT, 'gpurun_out', 'multi_gpu_check_%dgpu.log' % n), 'w') as f:
        f.write(r.stdout + '\n---- stderr ----\n' + r.stderr[-4000:])
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert 'multi-GPU check A ok' in r.stdout and 'multi-GPU check B ok' in r.stdout


def _cli(args, n, port, tmp):
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(n), '--master-addr', '127.0.0.1',
           '--master-port', str(port), '-m', 'mirge_b200'] if n > 1 else [sys.executable, '-m', 'mirge_b200']
    r = subprocess.run(cmd + args, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    outs = [d for d in os.listdir(tmp) if d.startswith('miRge.')]
    assert len(outs) == 1
    return os.path.join(tmp, outs[0])


def test_cli_on_n_ranks_equals_one_rank(tmp_path):
    """`torchrun -m mirge_b200 annotate` shards the FILE CHUNKS of every sample over the ranks (plain and gzipped input);
    its output directory must equal the single-process one: count tables byte for byte, mapped/unmapped as row sets."""
    import gzip
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip('needs at least two GPUs (gpurun --gpus 2)')
    n = min(n, 4)
    from synth.world import World
    w = World(seed=20180626, scale=0.05, snpc_names=True)      # the naming -gff expects next to .SNP entries
    libdir = tmp_path / 'libs'
    w.write(str(libdir))
    s1, s2 = tmp_path / 'a.fastq', tmp_path / 'b.fastq.gz'
    s1.write_bytes(w.fastq(120000, seed=31).tobytes())
    with gzip.open(s2, 'wb') as f:
        f.write(w.fastq(60000, seed=32).tobytes())
    common = ['annotate', '-s', str(s1), str(s2), '-lib', str(libdir), '-sp', 'synth', '-ad', 'illumina', '-di', '-tcf', '-gff']
    one, many = tmp_path / 'one', tmp_path / 'many'
    one.mkdir()
    many.mkdir()
    d1 = _cli(common + ['-o', str(one)], 1, 0, str(one))
    dn = _cli(common + ['-o', str(many)], n, 29541, str(many))
    files = sorted(os.listdir(d1))
    assert files == sorted(os.listdir(dn)) and 'miR.Counts.csv' in files
    for f in files:
        a, b = open(os.path.join(d1, f)).read(), open(os.path.join(dn, f)).read()
        if f in ('mapped.csv', 'unmapped.csv'):
            assert a.split('\n')[0] == b.split('\n')[0] and sorted(a.split('\n')) == sorted(b.split('\n')), f
        elif f.endswith('.gff') or f.startswith('isomirs'):
            assert sorted(a.split('\n')) == sorted(b.split('\n')), f       # row order follows the unique-sequence order
        else:
            assert a == b, f
