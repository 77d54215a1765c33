import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


def load_golden(name):
    """-> (npz dict, numpy state_dict regenerated from the committed key list)"""
    from oracle import fixtures as fx
    g = dict(np.load(os.path.join(GOLDEN, name + '.npz'), allow_pickle=False))
    shapes = [(k, tuple(s), d) for k, s, d in json.loads(str(g['keys']))]
    return g, fx.make_state(shapes)


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))


@pytest.fixture(scope='session')
def golden():
    return load_golden
