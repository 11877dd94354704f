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
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (B200)')


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason='no CUDA device in this container')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


@pytest.fixture(scope='session')
def mlgk_golden():
    return load_golden('mlgk_reference.json')


@pytest.fixture(scope='session')
def microkernel_golden():
    return load_golden('microkernel_reference.json')


def golden_kernels(name):
    """The node/edge microkernels of each golden case, rebuilt with this
    package's microkernels (same definitions as reference
    test/kernel/marginalized/test_kernel.py:129-170)."""
    from graphdot_b200.microkernel import (Additive, Constant, Convolution,
                                           KroneckerDelta, SquareExponential,
                                           TensorProduct)
    if name == 'unlabeled':
        return Constant(1.0), Constant(1.0)
    if name == 'labeled':
        return (TensorProduct(hybridization=KroneckerDelta(0.3),
                              charge=SquareExponential(1.) + 0.01).normalized,
                Additive(order=KroneckerDelta(0.3),
                         length=SquareExponential(0.05)).normalized)
    if name == 'weighted':
        return (Additive(hybridization=KroneckerDelta(0.3),
                         charge=SquareExponential(1.0)).normalized,
                TensorProduct(order=KroneckerDelta(0.3),
                              length=SquareExponential(0.05)))
    if name == 'vario-features':
        return (TensorProduct(rings=Convolution(KroneckerDelta(0.3))),
                TensorProduct(spectrum=Convolution(SquareExponential(1.0))))
    if name.startswith('molecular'):
        return (TensorProduct(element=KroneckerDelta(0.5),
                              x=SquareExponential(1.0)),
                TensorProduct(length=SquareExponential(0.1)))
    raise KeyError(name)


def golden_graphs(case):
    from graphdot_b200 import Graph
    return [Graph.from_columns(g['nodes'], g['edges'], g['title'])
            for g in case['graphs']]
