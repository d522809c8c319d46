"""CPU, build container only (skipped where /root/reference is absent, e.g. on the GPU box): the drop-in hooks rebind
the reference's operator surface, signatures are identical, and configs are recovered from reference modules."""
import inspect
import os
import sys

import pytest

REF = '/root/reference'
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason='reference tree not present')


@pytest.fixture(scope='module')
def ref():
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
    from oracle import make_golden
    make_golden.bootstrap_reference()
    import torch_utils.ops.bias_act as rba
    import torch_utils.ops.upfirdn2d as rup
    import torch_utils.ops.conv2d_resample as rcr
    import training.networks as rnet
    import thirdparty.stylegan2_ada_pytorch.training.networks as rnet2
    return rba, rup, rcr, rnet, rnet2


def _params(fn):
    # misc.profiled_function (SG2/torch_utils/misc.py:98-103) wraps without functools.wraps: unwrap through the closure
    if getattr(fn, '__closure__', None) and list(inspect.signature(fn).parameters) == ['args', 'kwargs']:
        fn = fn.__closure__[0].cell_contents
    return [(p.name, p.default) for p in inspect.signature(fn).parameters.values()]


def test_signatures_equal_reference(ref):
    rba, rup, rcr, rnet, _ = ref
    from brushstroke_engine_b200 import bias_act, upfirdn2d, conv2d_resample, modconv
    assert _params(bias_act.bias_act) == _params(rba.bias_act)
    for name in ('upfirdn2d', 'filter2d', 'upsample2d', 'downsample2d'):
        assert _params(getattr(upfirdn2d, name)) == _params(getattr(rup, name)), name
    assert [n for n, _ in _params(upfirdn2d.setup_filter)] == [n for n, _ in _params(rup.setup_filter)]
    assert _params(conv2d_resample.conv2d_resample) == _params(rcr.conv2d_resample)
    assert _params(modconv.modulated_conv2d) == _params(rnet.modulated_conv2d)
    assert {k: v[2] for k, v in bias_act.activation_funcs.items()} == {k: v.cuda_idx for k, v in rba.activation_funcs.items()}
    for k, v in rba.activation_funcs.items():
        assert abs(bias_act.activation_funcs[k][0] - v.def_alpha) < 1e-12 and abs(bias_act.activation_funcs[k][1] - v.def_gain) < 1e-12


def test_install_rebinds_both_aliases_and_restores(ref):
    rba, rup, rcr, rnet, rnet2 = ref
    from brushstroke_engine_b200 import install, bias_act, modconv
    orig = rba.bias_act
    patched = install.install_ops()
    try:
        assert rba.bias_act is bias_act.bias_act
        assert rnet.modulated_conv2d is modconv.modulated_conv2d and rnet2.modulated_conv2d is modconv.modulated_conv2d
        assert 'torch_utils.ops.upfirdn2d.upsample2d' in patched and 'torch_utils.ops.conv2d_resample.conv2d_resample' in patched
    finally:
        install.uninstall_ops()
    assert rba.bias_act is orig


def test_configs_recovered_from_reference_modules(ref, bundles):
    from oracle import make_golden
    from brushstroke_engine_b200 import install, params as P
    cfg, ecfg, gp, ep = bundles
    G, enc, _ = make_golden.build_reference_modules(gp, ep, cfg, ecfg)
    assert install.generator_config_from_reference(G) == cfg
    e2 = install.encoder_config_from_reference(enc)
    assert (e2.pre_filters, tuple(e2.down_filters), tuple(e2.post_filters), tuple(e2.encode_resolutions)) == \
        (ecfg.pre_filters, tuple(ecfg.down_filters), tuple(ecfg.post_filters), tuple(ecfg.encode_resolutions))
    b = P.bundle_from_module(G)
    for k, v in gp.items():
        assert k in b and (b[k] - v).abs().max() == 0, k
