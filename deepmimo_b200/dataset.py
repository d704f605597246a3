"""`Dataset` / `MacroDataset`: host-side mirror of the reference containers, channel path only.

The reference `Dataset` (deepmimo/generator/dataset.py:61-869) is a lazy dict of arrays; the part on
the channel path is `compute_channels` (:224-268), `set_channel_params` (:197-222), `apply_fov`
(:423-448) with its cache invalidation (:515-535, :358-378), the lazy `channel` / `n_ue` keys
(:831-838, :657-659) and the aliases (deepmimo/consts.py:261-322).  `MacroDataset` (:888-998) fans a
call out to one child per base station.  Only that surface is mirrored here; everything numeric is
delegated to `channels.compute_channels` (CUDA).  Loading, plotting, sampling are out of scope.
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional

import numpy as np

from . import channels as _ch
from .params import ChannelGenParameters, DotDict

ALIASES = {            # subset of deepmimo/consts.py:261-322 that touches this path
    "ch": "channel", "chs": "channel", "channels": "channel", "channel_params": "ch_params",
    "pwr": "power", "powers": "power", "toa": "delay", "time_of_arrival": "delay",
    "aoa_phi": "aoa_az", "aoa_theta": "aoa_el", "aod_phi": "aod_az", "aod_theta": "aod_el",
    "ue_pos": "rx_pos", "rx_loc": "rx_pos", "bs_pos": "tx_pos", "tx_loc": "tx_pos",
    "pwr_ant_gain": "_power_linear_ant_gain", "lin_pwr": "power_linear", "linear_power": "power_linear", "pwr_lin": "power_linear",
    "pl": "pathloss", "path_loss": "pathloss", "n_paths": "num_paths", "los_status": "los",
    "bounce_type": "inter", "interactions": "inter",
}

_HOST_BYPRODUCTS = {"power_linear": "_compute_power_linear", "pathloss": "compute_pathloss",
                    "num_paths": "_compute_num_paths", "los": "_compute_los"}

ROT_KEYS = ("_aod_el_rot", "_aod_az_rot", "_aoa_el_rot", "_aoa_az_rot")                 # consts.py:212-215
FOV_KEYS = ("_fov_mask", "num_paths", "los", "channel", "_power_linear_ant_gain",        # dataset.py:527-532
            "_aod_el_rot_fov", "_aod_az_rot_fov", "_aoa_el_rot_fov", "_aoa_az_rot_fov")


class Dataset(DotDict):
    """Path matrices of one (TX, RX-set) pair + GPU channel generation."""

    def __init__(self, data: Optional[Dict[str, Any]] = None):
        object.__setattr__(self, "_data", {})
        for k, v in (data or {}).items():
            self._data[k] = v          # arrays are stored as is (no dict wrapping of array payloads)

    # -- lazy / aliased access (dataset.py:130-182)
    def _resolve(self, key: str):
        key = ALIASES.get(key, key)
        if key in self._data:
            return self._data[key]
        if key == "n_ue":
            self._data[key] = int(self._data["rx_pos"].shape[0] if "rx_pos" in self._data
                                  else self._data["power"].shape[0])
            return self._data[key]
        if key == "channel":
            return self.compute_channels()
        if key == "ch_params":
            self.set_channel_params()
            return self._data[key]
        if key in ROT_KEYS + FOV_KEYS[4:] + ("_fov_mask",):
            self._compute_path_byproducts()
            return self._data[key]
        if key in _HOST_BYPRODUCTS:
            self._data[key] = getattr(self, _HOST_BYPRODUCTS[key])()
            return self._data[key]
        raise KeyError(key)

    def __getitem__(self, key):
        try:
            return self._data[key]
        except KeyError:
            return self._resolve(key)

    def __getattr__(self, key):
        if key.startswith("__"):
            raise AttributeError(key)
        try:
            return self._data[key]
        except KeyError:
            try:
                return self._resolve(key)
            except KeyError:
                raise AttributeError(key) from None

    def __setitem__(self, key, value):
        self._data[ALIASES.get(key, key)] = value

    # -- channel parameters (dataset.py:197-222)
    def set_channel_params(self, params: Optional[ChannelGenParameters] = None) -> ChannelGenParameters:
        if params is None:
            params = ChannelGenParameters()
        params.validate(self.n_ue)
        old = self._data.get("ch_params")
        self._data["ch_params"] = params.deepcopy()
        if old is not None:
            same = (np.array_equal(old["bs_antenna"]["rotation"], params["bs_antenna"]["rotation"]) and
                    np.array_equal(old["ue_antenna"]["rotation"], params["ue_antenna"]["rotation"]))
            if not same:
                self._clear_cache_rotated_angles()
        return params

    # -- FoV (dataset.py:423-448)
    def apply_fov(self, bs_fov=np.array([360, 180]), ue_fov=np.array([360, 180])) -> None:
        self._clear_cache_fov()
        self._data["bs_fov"] = np.asarray(bs_fov)
        self._data["ue_fov"] = np.asarray(ue_fov)

    def _clear_cache_fov(self) -> None:
        for k in FOV_KEYS:
            self._data.pop(k, None)

    def _clear_cache_rotated_angles(self) -> None:
        for k in ROT_KEYS:
            self._data.pop(k, None)
        self._clear_cache_fov()

    # -- the hot path (dataset.py:224-268)
    def compute_channels(self, params: Optional[ChannelGenParameters] = None, **kwargs):
        """GPU replacement of the reference method; see channels.compute_channels for the extras."""
        if params is None:
            params = self._data.get("ch_params") or ChannelGenParameters()
        return _ch.compute_channels(self, params, **kwargs)

    # -- first consumer of H
    def beam_amplitude(self, beams, params: Optional[ChannelGenParameters] = None, **kwargs):
        """Beam amplitude map `np.abs(beams @ channel).mean(axis=1).mean(axis=-1)` without materialising the channel
        (docs/manual.ipynb cell 105; deepmimo_b200/beams.py)."""
        from .beams import beam_amplitude
        if params is None:
            params = self._data.get("ch_params") or ChannelGenParameters()
        return beam_amplitude(self, beams, params, **kwargs)

    # -- per-user by-products: cheap host reductions over the path matrices and the GPU FoV mask
    def _compute_power_linear(self) -> np.ndarray:
        """dataset.py:694-696, generator_utils.py:35."""
        return 10 ** (self["power"] / 10)

    def compute_pathloss(self, coherent: bool = True) -> np.ndarray:
        """Path loss in dB per user (dataset.py:541-566): -10 log10 |sum_p sqrt(p_lin) e^{j phase}|^2, NaN where no power.
        Computed by the per-user by-product kernel (one warp per user)."""
        from .byproducts import user_byproducts
        r = user_byproducts(self, self._data.get("ch_params"), want=("pathloss",))
        pl = r["pathloss"] if coherent else r["pathloss_noncoherent"]
        self._data["pathloss"] = pl
        return pl

    def _compute_num_paths(self) -> np.ndarray:
        """Valid paths per user after FoV filtering (dataset.py:613-619), device reduction over the prologue's FoV mask."""
        from .byproducts import user_byproducts
        return user_byproducts(self, self._data.get("ch_params"), want=("num_paths",))["num_paths"]

    def _compute_los(self) -> np.ndarray:
        """1 LoS / 0 NLoS / -1 no path: interaction code of the first in-FoV path (dataset.py:569-611), on the device."""
        from .byproducts import user_byproducts
        return user_byproducts(self, self._data.get("ch_params"), want=("los",))["los"]

    def _compute_path_byproducts(self) -> None:
        """Rotated angles, FoV mask/angles and power with antenna gain (dataset.py:310-356, :461-512,
        :665-691) from the prologue kernel; fills the same cache keys as the reference."""
        from .byproducts import path_byproducts
        self._data.update(path_byproducts(self, self._data.get("ch_params")))


class MacroDataset:
    """List of per-BS datasets; method calls fan out to every child (dataset.py:888-998)."""

    def __init__(self, datasets: Optional[List[Dataset]] = None):
        self.datasets = list(datasets) if datasets is not None else []

    def __len__(self) -> int:
        return len(self.datasets)

    def __getitem__(self, idx):
        if isinstance(idx, (int, slice)):
            return self.datasets[idx]
        res = [d[idx] for d in self.datasets]
        return res[0] if len(res) == 1 else res

    def __setitem__(self, key, value) -> None:
        for d in self.datasets:
            d[key] = value

    def append(self, dataset: Dataset) -> None:
        self.datasets.append(dataset)

    def compute_channels(self, params: Optional[ChannelGenParameters] = None, *, devices=None, **kwargs):
        """Per-child `compute_channels` (dataset.py:939-950: a list of arrays, or the single result).

        `devices` (e.g. ["cuda:0", "cuda:1"]) spreads the children round-robin over several GPUs of this process, one
        host thread per device: base stations are independent, so there is no exchange between them.  Without
        it the children run one after the other on the current device, like the reference's loop."""
        if not devices or len(self.datasets) <= 1:
            if devices:
                kwargs.setdefault("device", devices[0])
            res = [d.compute_channels(params, **kwargs) for d in self.datasets]
            return res[0] if len(res) == 1 else res
        from concurrent.futures import ThreadPoolExecutor
        devices = list(devices)

        def worker(slot):
            return [(i, self.datasets[i].compute_channels(params, device=devices[slot], **kwargs))
                    for i in range(slot, len(self.datasets), len(devices))]

        out = [None] * len(self.datasets)
        with ThreadPoolExecutor(len(devices)) as pool:
            for part in pool.map(worker, range(len(devices))):
                for i, h in part:
                    out[i] = h
        return out

    def __getattr__(self, name):
        if name.startswith("__") or name == "datasets":
            raise AttributeError(name)
        attr = getattr(Dataset, name, None)
        if callable(attr):
            def fan_out(*args, **kwargs):
                res = [getattr(d, name)(*args, **kwargs) for d in self.datasets]
                return res[0] if len(res) == 1 else res
            return fan_out
        res = [getattr(d, name) for d in self.datasets]
        return res[0] if len(res) == 1 else res
