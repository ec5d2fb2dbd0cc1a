import ctypes
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

ILLUMINA = 'TGGAATTCTCGGGTGCCAAGGAACTCCAG'   # src/mirge/__main__.py:124-125


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


@pytest.fixture(scope='session')
def small_world():
    from synth.world import World
    return World(seed=20180626, scale=0.05)


@pytest.fixture(scope='session')
def oracle_mod():
    from oracle import oracle as O
    O.lib()
    return O


@pytest.fixture(scope='session')
def host_kernel():
    """Device functions of the trim kernels compiled for the HOST (tests/helpers/prefilter_host.cu): k1_prefilter,
    k1_match_full, match_back / locate_back are __host__ __device__, so their exact code runs on the CPU."""
    import shutil
    nvcc = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(nvcc):
        pytest.skip('nvcc not available')
    so = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'helpers', 'libprefilter_host.so')
    srcs = [os.path.join(os.path.dirname(os.path.abspath(__file__)), 'helpers', 'prefilter_host.cu'),
            os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'mirge_b200', 'csrc', 'k1_prefilter.cuh'), os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'mirge_b200', 'csrc', 'common.cuh'),
            os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'mirge_b200', 'csrc', 'k1_kernel.cuh'),
            os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'mirge_b200', 'csrc', 'k1_trim.cuh')]
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(s) for s in srcs):
        subprocess.check_call([nvcc, '-O2', '-std=c++17', '-arch=sm_100a', '-shared', '-Xcompiler', '-fPIC', '-o', so, srcs[0]])
    return ctypes.CDLL(so)
