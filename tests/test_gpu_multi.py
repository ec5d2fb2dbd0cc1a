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
