import os
import sys

import numpy as np
import pytest
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

GOLDEN = os.path.join(REPO, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason='no CUDA device')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name + '.npz')) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture(scope='session')
def bundles():
    """The seeded parameter bundles every fixture in tests/golden was generated with
    (oracle/make_golden.py main())."""
    from brushstroke_engine_b200 import params as P
    cfg, ecfg = P.GeneratorConfig(), P.EncoderConfig()
    gp = P.init_generator_params(cfg, seed=0, perturb=0.1)
    ep = P.init_encoder_params(ecfg, seed=1, perturb_bn=0.1)
    g = load_golden('generator')
    assert bytes(g['gen_digest']).decode() == P.bundle_digest(gp), 'generator weights differ from the golden run'
    assert bytes(g['enc_digest']).decode() == P.bundle_digest(ep), 'encoder weights differ from the golden run'
    return cfg, ecfg, gp, ep


def t(a):
    return torch.from_numpy(np.ascontiguousarray(a))
