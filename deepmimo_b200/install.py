"""Opt-in monkey-patch: route the reference's own `Dataset.compute_channels` through the CUDA path.

    import deepmimo as dm, deepmimo_b200 as dmb
    dmb.install(dm)                      # or dmb.install(dm, only_if_config=True) to honour dm.config('use_gpu')
    H = dataset.compute_channels(params) # unchanged user code (deepmimo/generator/dataset.py:224)

`dm.config` declares `use_gpu` / `gpu_device_id` (deepmimo/config.py:58-59) but nothing in the reference
reads them; `only_if_config=True` makes them the switch.
"""
from __future__ import annotations

from . import channels as _ch

_ORIG = {}


def install(dm, only_if_config: bool = False):
    """Rebind `dm.Dataset.compute_channels` (and so `dataset.channel`, MacroDataset fan-out) to the GPU path."""
    cls = dm.Dataset
    if cls in _ORIG:
        return cls
    orig = cls.compute_channels
    _ORIG[cls] = orig

    def compute_channels(self, params=None, **kwargs):
        if only_if_config:
            try:
                use_gpu = bool(dm.config.get("use_gpu"))
                dev = dm.config.get("gpu_device_id")
            except Exception:  # noqa: BLE001
                use_gpu, dev = False, None
            if not use_gpu:
                return orig(self, params)
            kwargs.setdefault("device", None if dev is None else f"cuda:{int(dev)}")
        if params is None:
            params = dm.ChannelGenParameters() if self.ch_params is None else self.ch_params
        return _ch.compute_channels(self, params, **kwargs)

    compute_channels.__doc__ = orig.__doc__
    cls.compute_channels = compute_channels
    return cls


def uninstall(dm):
    cls = dm.Dataset
    if cls in _ORIG:
        cls.compute_channels = _ORIG.pop(cls)
