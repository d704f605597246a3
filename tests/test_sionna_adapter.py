"""Row f4b: the (a, tau) layout of the reference's Sionna adapter (deepmimo/integrations/sionna_adapter.py:174-200) from the
time-domain kernel.  tests/golden/sionna.npz was produced by the live adapter class on top of the live v4 time-domain channels
(tests/golden/make_golden_sionna.py)."""
import os

import numpy as np
import pytest

from util import TOL_REL_FRO, per_user_rel_fro

HERE = os.path.dirname(os.path.abspath(__file__))


def _case():
    from make_golden_sionna import sionna_case
    return sionna_case() + (np.load(os.path.join(HERE, "golden", "sionna.npz")),)


def test_adapter_shapes_and_index_handling_without_gpu():
    import deepmimo_b200 as dmb
    from deepmimo_b200.sionna_adapter import DeepMIMOSionnaAdapter
    data, params, bs_idx, ue_idx, g = _case()
    dss = [dmb.Dataset(dict(d)) for d in data]
    ad = DeepMIMOSionnaAdapter(dss, dmb.ChannelGenParameters(params), bs_idx=bs_idx, ue_idx=ue_idx)
    assert len(ad) == int(g["multi_len"]) == 8
    assert ad.ch_shape == g["multi_a"].shape[1:] == (2, 2, 2, 8, 10, 1) and ad.t_shape == g["multi_tau"].shape[1:]
    ad0 = DeepMIMOSionnaAdapter(dss, dmb.ChannelGenParameters(params))
    assert len(ad0) == int(g["default_len"]) == 40 and ad0.ch_shape == g["default_a"].shape[1:]
    assert DeepMIMOSionnaAdapter(dss, bs_idx=1, ue_idx=[3, 4, 5]).ue_idx.shape == (3, 1)
    with pytest.raises(TypeError):
        DeepMIMOSionnaAdapter(dss, ue_idx="all")
    with pytest.raises(ValueError):
        DeepMIMOSionnaAdapter(dss, ue_idx=np.zeros((2, 2, 2), dtype=int))


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["multi", "default"])
def test_gpu_sionna_arrays_match_the_live_adapter(name):
    import deepmimo_b200 as dmb
    from deepmimo_b200.sionna_adapter import DeepMIMOSionnaAdapter
    data, params, bs_idx, ue_idx, g = _case()
    dss = dmb.MacroDataset([dmb.Dataset(dict(d)) for d in data])
    kw = dict(bs_idx=bs_idx, ue_idx=ue_idx) if name == "multi" else {}
    ad = DeepMIMOSionnaAdapter(dss, dmb.ChannelGenParameters(params), **kw)
    a, tau = ad.arrays()
    ga, gt = g[f"{name}_a"], g[f"{name}_tau"]
    assert a.shape == ga.shape and a.dtype == np.complex64 and tau.shape == gt.shape and tau.dtype == np.float32
    assert np.array_equal(tau, gt)                                   # delays are copied, bit for bit, zeros in the empty slots
    err = per_user_rel_fro(a, ga)
    assert err.max() <= TOL_REL_FRO, err.max()
    assert np.array_equal(a == 0, ga == 0)                           # empty slots / users without paths are exact zeros
    samples = list(ad())                                             # generator protocol of the reference
    assert len(samples) == len(ad) and np.array_equal(samples[3][0], a[3]) and np.array_equal(samples[3][1], tau[3])
    at, tt = ad.arrays(out="torch")
    assert at.is_cuda and np.array_equal(at.cpu().numpy(), a)
