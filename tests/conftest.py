import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'oracle')):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a B200 (run with -m gpu on the GPU box)')
    # the CUDA library is built in-tree (nvcc cross-compiles without a GPU); tests never fall back to anything else
    import __graft_entry__ as ge
    ge.build()


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason='no CUDA device in this container (GPU tests run under gpurun)')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope='session')
def golden_dir():
    return GOLDEN


def load_golden(name):
    import numpy as np
    return dict(np.load(os.path.join(GOLDEN, name + '.npz'), allow_pickle=False))


def golden_names(prefixes=('cfg2', 'rec_', 'seg_', 'misc_')):
    return sorted(f[:-4] for f in os.listdir(GOLDEN) if f.endswith('.npz') and f.startswith(tuple(prefixes)))


def dec_from_golden(g):
    out = []
    for i, c in enumerate(g['dec_count'].tolist()):
        out.append([(int(g['dec_label'][i, j]), int(g['dec_start'][i, j]), int(g['dec_end'][i, j]), float(g['dec_conf'][i, j])) for j in range(c)])
    return out
