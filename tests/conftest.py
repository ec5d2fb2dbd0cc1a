import os
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
