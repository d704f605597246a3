"""Channel-generation parameters: host-side mirror of the reference's `ChannelGenParameters`.

Mirrors deepmimo/generator/channel.py:20-139 (class, defaults, `validate`) and the `DotDict`
semantics it inherits (deepmimo/general_utils.py:124-261): dot and item access at every level,
nested dicts become DotDicts, shallow `update`, deep `deepcopy`.  Key names are the reference's
(deepmimo/consts.py:235-254).  The north_star's spelling `ChannelParameters` is an alias, and
`fov` inside an antenna dict is accepted as an alternative to `Dataset.apply_fov`.

A reference `dm.ChannelGenParameters` object can be passed wherever this class is accepted: the
driver only uses item access.
"""
from __future__ import annotations

from copy import deepcopy
from pprint import pformat
from typing import Any, Dict, Iterator, Mapping, Optional

import numpy as np

MAX_PATHS = 25                                         # deepmimo/consts.py:180
RADIATION_PATTERNS = ("isotropic", "halfwave-dipole")  # deepmimo/consts.py:254


class DotDict(Mapping):
    """dict with attribute access; nested dicts are wrapped recursively."""

    def __init__(self, data: Optional[Mapping] = None):
        object.__setattr__(self, "_data", {})
        for k, v in (data or {}).items():
            self[k] = v

    # -- access
    def __getattr__(self, key: str) -> Any:
        try:
            return self._data[key]
        except KeyError:
            raise AttributeError(key) from None

    def __setattr__(self, key: str, value: Any) -> None:
        self[key] = value

    def __getitem__(self, key: str) -> Any:
        return self._data[key]

    def __setitem__(self, key: str, value: Any) -> None:
        if isinstance(value, dict):
            value = DotDict(value)
        self._data[key] = value

    def __delitem__(self, key: str) -> None:
        del self._data[key]

    def __iter__(self) -> Iterator:
        return iter(self._data)

    def __len__(self) -> int:
        return len(self._data)

    def __contains__(self, key) -> bool:
        return key in self._data

    def __dir__(self):
        return sorted(set(list(super().__dir__()) + list(self._data)))

    def __repr__(self) -> str:
        return pformat(self._data)

    def keys(self):
        return self._data.keys()

    def values(self):
        return self._data.values()

    def items(self):
        return self._data.items()

    def get(self, key, default=None):
        return self._data.get(key, default)

    def update(self, other: Mapping) -> None:
        """Shallow update, like the reference: a nested dict replaces the whole sub-dict."""
        for k, v in other.items():
            self[k] = v

    def to_dict(self) -> Dict:
        return {k: (v.to_dict() if isinstance(v, DotDict) else v) for k, v in self._data.items()}

    def deepcopy(self):
        out = {}
        for k, v in self._data.items():
            if isinstance(v, DotDict):
                out[k] = v.deepcopy().to_dict()
            elif isinstance(v, np.ndarray):
                out[k] = v.copy()
            else:
                out[k] = deepcopy(v)
        return type(self)(out)


def _default_params() -> Dict:
    """Defaults of deepmimo/generator/channel.py:33-63."""
    return {
        "bs_antenna": {"shape": np.array([8, 1]), "spacing": 0.5, "rotation": np.array([0, 0, 0]),
                       "radiation_pattern": RADIATION_PATTERNS[0]},
        "ue_antenna": {"shape": np.array([1, 1]), "spacing": 0.5, "rotation": np.array([0, 0, 0]),
                       "radiation_pattern": RADIATION_PATTERNS[0]},
        "enable_doppler": 0,
        "enable_dual_polar": 0,
        "num_paths": MAX_PATHS,
        "freq_domain": 1,
        "ofdm": {"subcarriers": 512, "selected_subcarriers": np.arange(1), "bandwidth": 10e6, "rx_filter": 0},
    }


def _extra_keys(d: Mapping, ref: Mapping, prefix: str = "") -> list:
    extra = []
    for k, v in d.items():
        if k not in ref:
            extra.append(prefix + str(k))
        elif isinstance(v, Mapping) and isinstance(ref[k], Mapping):
            extra += _extra_keys(v, ref[k], prefix + str(k) + ".")
    return extra


class ChannelGenParameters(DotDict):
    """Channel generation parameters (bs/ue antenna, OFDM, domain, num_paths).

    `ChannelGenParameters()` gives the reference defaults; `ChannelGenParameters(dict)` applies a
    shallow update on top of them (channel.py:65-76).
    """

    def __init__(self, data: Optional[Mapping] = None):
        super().__init__(_default_params())
        if data is not None:
            self.update(data)

    def validate(self, n_ues: int) -> "ChannelGenParameters":
        """Same checks as channel.py:78-139 (assertions on rotation shapes and pattern names)."""
        known = _default_params()
        known["bs_antenna"]["fov"] = known["ue_antenna"]["fov"] = None   # accepted alternative spelling
        extra = _extra_keys(self, known)
        if extra:
            print("The following parameters seem unnecessary:")
            print(extra)

        bs, ue = self["bs_antenna"], self["ue_antenna"]
        if "rotation" in bs.keys():
            shp = np.asarray(bs["rotation"]).shape
            assert len(shp) == 1 and shp[0] == 3, "The BS antenna rotation must be a 3D vector"
        else:
            bs["rotation"] = None

        if "rotation" in ue.keys() and ue["rotation"] is not None:
            shp = np.asarray(ue["rotation"]).shape
            ok = (len(shp) == 1 and shp[0] == 3) or (len(shp) == 2 and shp == (3, 2)) or (shp[0] == n_ues)
            assert ok, ("The UE antenna rotation must either be a 3D vector for constant values "
                        "or 3 x 2 matrix for random values")
        else:
            ue["rotation"] = np.array([0, 0, 0])

        for side, name in ((bs, "BS"), (ue, "UE")):
            if "radiation_pattern" in side.keys() and side["rotation"] is not None:
                assert side["radiation_pattern"] in RADIATION_PATTERNS, (
                    f"The {name} antenna radiation pattern must have one of the following values: "
                    f"{list(RADIATION_PATTERNS)}")
            else:
                side["radiation_pattern"] = RADIATION_PATTERNS[0]
        return self


ChannelParameters = ChannelGenParameters   # north_star / later-release spelling
