"""Row f1: scenario loader.  Writes a synthetic scenario in the reference's v4 on-disk format, loads it back, and -- in
the build container only, where /root/reference exists -- checks the result against the reference's own
`_load_tx_rx_raydata` (deepmimo/generator/core.py:186-258) on the same files."""
import os
import sys
from unittest.mock import MagicMock

import numpy as np
import pytest

import deepmimo_b200 as dmb
from deepmimo_b200.synth import make_paths


def _write(tmp_path):
    pairs = {}
    for tx_idx in range(2):
        d = make_paths(30, 70 + tx_idx, n_cols=25)
        d["inter_pos"] = np.zeros((30, 25, 3, 3), np.float32)
        pairs[(0, tx_idx, 1)] = d
    return dmb.save_scenario(str(tmp_path / "synth_scen"), pairs), pairs


def test_round_trip_and_selection(tmp_path):
    folder, pairs = _write(tmp_path)
    macro = dmb.load_scenario(folder)
    assert isinstance(macro, dmb.MacroDataset) and len(macro) == 2
    for ti in range(2):
        ds = macro[ti]
        assert ds["txrx"] == {"tx_set_id": 0, "rx_set_id": 1, "tx_idx": ti}
        for k in ("power", "phase", "delay", "aoa_az", "aoa_el", "aod_az", "aod_el"):
            assert ds[k].dtype == np.float32 and ds[k].flags.c_contiguous
            assert np.array_equal(ds[k], pairs[(0, ti, 1)][k], equal_nan=True)
        assert ds.n_ue == 30
    one = dmb.load_scenario(folder, max_paths=7, tx_sets={0: [1]}, rx_sets={1: np.arange(5, 20)}, matrices=["power", "delay", "rx_pos"])
    assert isinstance(one, dmb.Dataset) and one["power"].shape == (15, 7) and one["rx_pos"].shape == (15, 3)
    assert np.array_equal(one["power"], pairs[(0, 1, 1)]["power"][5:20, :7], equal_nan=True)
    assert "phase" not in one.keys()
    with pytest.raises(ValueError):
        dmb.load_scenario(folder, tx_sets=[5])
    with pytest.raises(ValueError):
        dmb.load_scenario(folder, matrices=["nope"])
    with pytest.raises(ValueError):
        dmb.load_scenario(str(tmp_path / "missing"))


@pytest.mark.skipif(not os.path.isdir("/root/reference/deepmimo"), reason="reference tree only exists in the build container")
def test_matches_reference_loader(tmp_path):
    for m in ("matplotlib", "matplotlib.pyplot", "matplotlib.figure", "matplotlib.axes", "matplotlib.colorbar",
              "matplotlib.colors", "mpl_toolkits", "mpl_toolkits.mplot3d", "mpl_toolkits.mplot3d.art3d"):
        sys.modules.setdefault(m, MagicMock())
    sys.path.insert(0, "/root/reference")
    from deepmimo.generator.core import _load_tx_rx_raydata
    folder, _ = _write(tmp_path)
    rx = np.array([0, 3, 4, 29])
    ref = _load_tx_rx_raydata(folder, 0, 1, 1, rx, 10, "all")
    got = dmb.load_tx_rx_raydata(folder, 0, 1, 1, rx, 10, "all")
    for k, v in ref.items():
        if v is None:
            continue
        assert np.array_equal(got[k], v, equal_nan=True), k
        assert got[k].shape == v.shape


@pytest.mark.gpu
def test_gpu_scenario_folder_to_channels(tmp_path):
    """Row f1 as a device feed: scenario folder -> parse / H2D pipeline -> CUDA path matrices -> compute_channels, against the oracle."""
    import torch
    from oracle import channel_oracle as orc
    from util import assert_channels_close, oracle_kwargs_from_params
    from deepmimo_b200.synth import scenario
    s = scenario(5, 700)
    pairs = {(0, 0, 1): dict(s.data), (0, 1, 1): dict(scenario(5, 700, bs_index=1).data)}
    folder = dmb.save_scenario(str(tmp_path / "city"), pairs)
    macro = dmb.load_scenario(folder, device="cuda")
    assert len(macro) == 2
    p = dmb.ChannelGenParameters(s.params)
    for ti in range(2):
        ds = macro[ti]
        for k in ("power", "phase", "delay", "aoa_az", "aoa_el", "aod_az", "aod_el"):
            assert isinstance(ds[k], torch.Tensor) and ds[k].is_cuda and ds[k].dtype == torch.float32
            assert np.array_equal(ds[k].cpu().numpy(), pairs[(0, ti, 1)][k], equal_nan=True)
        assert isinstance(ds["rx_pos"], np.ndarray) and ds.n_ue == 700
        H = ds.compute_channels(p, warn=False)
        o = orc.compute_channels(pairs[(0, ti, 1)], **oracle_kwargs_from_params(s.params))
        assert_channels_close(H, o["H"], what=f"loaded pair {ti}")
    host = dmb.load_scenario(folder, pin=True)                      # host route: same numbers
    assert np.array_equal(host[0].compute_channels(p, warn=False), macro[0].compute_channels(p, warn=False))
