"""GPU parity against the committed reference-generated golden vectors (tests/golden/*.npz).

The fixtures were produced by the LIVE reference (tests/golden/make_golden.py); inputs are regenerated
from the case seeds.  Channel coefficients: per-user relative Frobenius error <= 1e-5 (north_star);
FoV mask and path validity: bit-exact.  Every call goes through the C ABI (libdmk.so).
"""
import os

import numpy as np
import pytest

from cases import case_data, case_list, params_dict
from util import assert_channels_close, make_dataset

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = case_list()


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_matches_reference_golden(case):
    import deepmimo_b200 as dmb
    g = np.load(os.path.join(GOLD, case["name"] + ".npz"))
    ds = make_dataset(dmb, case_data(case), case["bs_fov"], case["ue_fov"])
    H, info = ds.compute_channels(dmb.ChannelGenParameters(params_dict(case)), return_info=True, warn=False)
    assert_channels_close(H, g["H"], what=case["name"])
    assert np.array_equal(info.valid, g["valid"]), "path validity mask must be bit-exact"
    if bool(g["has_fov"]):
        assert info.fov_mask is not None and np.array_equal(info.fov_mask, g["fov_mask"]), "FoV mask must be bit-exact"
    else:
        assert info.fov_mask is None
    assert ds["channel"] is H          # cached like dataset.py:266
    assert info.launches > 0 and "kernel" in info.kernel
