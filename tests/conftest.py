import json
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA sm_100 (B200) device; run with -m gpu on the GPU box')


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason='no CUDA device in this container')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


def golden_index():
    with open(os.path.join(GOLDEN_DIR, 'index.json')) as f:
        return json.load(f)['cases']


def golden_case(name):
    """-> (kind, state_dict, x, y_reference, meta) for a committed fixture; weights regenerated from the seed."""
    from resselt_b200 import archs

    info = golden_index()[name]
    cls = {'SPAN': archs.SPAN, 'SPANPlus': archs.SpanPlus, 'Compact': archs.SRVGGNetCompact,
           'ESRGAN': getattr(archs, 'RRDBNet', None), 'RealPLKSR': getattr(archs, 'RealPLKSR', None), 'DAT': getattr(archs, 'DAT', None), 'SwinIR': getattr(archs, 'SwinIR', None), 'PLKSR': getattr(archs, 'PLKSR', None), 'SpanPP': getattr(archs, 'SpanPP', None), 'RTMoSR': getattr(archs, 'RTMoSR', None), 'GateRV3': getattr(archs, 'GateRV3', None)}[info['kind']]
    model = cls(seed=info['weight_seed'], **info['kwargs'])
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    data = np.load(os.path.join(GOLDEN_DIR, name + '.npz'))
    return info['kind'], sd, torch.from_numpy(data['x']), torch.from_numpy(data['y']), info['meta']


def psnr(y, ref):
    """PSNR with peak = the reference output's own range (random-init outputs are not in [0, 1])."""
    span = float(ref.max() - ref.min())
    mse = float(((y.double() - ref.double()) ** 2).mean())
    return 10.0 * np.log10(span * span / max(mse, 1e-30))


def norm_err(y, ref):
    span = max(1.0, float(ref.max() - ref.min()))
    return float((y.double() - ref.double()).abs().max()) / span
