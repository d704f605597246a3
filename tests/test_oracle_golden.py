"""The oracle (oracle/channel_oracle.py) against the committed reference-generated golden vectors.

tests/golden/*.npz were written by tests/golden/make_golden.py from the LIVE reference
(jmoraispk/DeepMIMO v4.0.0a3, `Dataset.compute_channels`); this replays every case on CPU.  In the
build container the restatement is bit-identical to the reference (manifest.json records it); the
assertion allows 1e-12 so that another host's NumPy SIMD dispatch cannot make it flaky.
"""
import json
import os

import numpy as np
import pytest

from cases import case_data, case_list, oracle_kwargs
from oracle import channel_oracle as orc
from util import per_user_rel_fro

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = case_list()


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_oracle_matches_reference_golden(case):
    g = np.load(os.path.join(GOLD, case["name"] + ".npz"))
    o = orc.compute_channels(case_data(case), **oracle_kwargs(case))
    assert o["H"].shape == g["H"].shape and o["H"].dtype == np.complex64
    err = per_user_rel_fro(o["H"], g["H"])
    assert err.size == 0 or err.max() <= 1e-12
    assert np.array_equal(o["valid"], g["valid"])
    if bool(g["has_fov"]):
        assert np.array_equal(o["fov_mask"], g["fov_mask"])
    else:
        assert o["fov_mask"] is None


def test_manifest_records_the_pin():
    with open(os.path.join(GOLD, "manifest.json")) as f:
        m = json.load(f)
    assert m["reference"].startswith("4.0.0")
    assert set(m["cases"]) == {c["name"] for c in CASES}
    for name, rec in m["cases"].items():
        assert rec["oracle_max_rel_fro"] <= 1e-12 and rec["fov_mask_equal"] and rec["valid_equal"], name
    assert sum(rec["nonzero_users"] for rec in m["cases"].values()) > 300


def test_time_domain_layout_and_slots():
    """channel.py:285-287: valid paths are compacted to the leading slots; FoV-masked paths keep a zero slot."""
    case = next(c for c in CASES if c["name"] == "holes_td")
    d = case_data(case)
    o = orc.compute_channels(d, **oracle_kwargs(case))
    n_valid = o["valid"].sum(1)
    for i in range(d["power"].shape[0]):
        assert np.all(o["H"][i, ..., n_valid[i]:] == 0)
        assert np.array_equal(np.sort(o["path_slot"][i][o["valid"][i]]), np.arange(n_valid[i]))
    assert (o["path_slot"][~o["valid"]] == -1).all()


def test_doppler_extension_reduces_to_reference_at_t0():
    """Row a11 (parity unpinned): at t = 0 the Doppler/time extension must equal the reference path."""
    from deepmimo_b200.synth import doppler_from_velocity
    for name in ("cfg1_shape", "td_basic"):
        case = next(c for c in CASES if c["name"] == name)
        d = case_data(case)
        fd = doppler_from_velocity(d, 9, 3.5e9)
        base = orc.compute_channels(d, **oracle_kwargs(case))["H"]
        ext = orc.compute_channels(d, **oracle_kwargs(case), doppler_hz=fd, times=np.array([0.0, 1e-3]))["H"]
        assert ext.shape == base.shape + (2,)
        assert np.array_equal(ext[..., 0], base)
        assert not np.array_equal(ext[..., 1], base)
        if name == "td_basic":      # one path per slot: Doppler only rotates the phase
            np.testing.assert_allclose(np.abs(ext[..., 1]), np.abs(base), rtol=1e-5, atol=1e-15)
