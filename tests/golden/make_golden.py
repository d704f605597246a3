"""Generate golden vectors by running the LIVE reference (jmoraispk/DeepMIMO v4.0.0a3).

Run in the build container only (needs /root/reference; the GPU box has no copy):

    python tests/golden/make_golden.py

For every case in cases.py this imports the unmodified reference (matplotlib
stubbed -- plotting only: deepmimo/scene.py:40, generator/visualization.py:20,
summary.py:39), builds `dm.Dataset(dict)`, calls `apply_fov` when the case has a
FoV and `compute_channels(params)` (deepmimo/generator/dataset.py:224), and
stores inputs-free outputs (inputs are regenerated from the seed by cases.py):
H (complex64), `_fov_mask`, `~isnan(_power_linear_ant_gain[:, :num_paths])`.
It also runs oracle/channel_oracle.py on the same inputs and records how far the
restatement is from the reference (manifest.json) -- that is the oracle's pin.
"""
import json
import os
import sys
from unittest.mock import MagicMock

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

for m in ("matplotlib", "matplotlib.pyplot", "matplotlib.figure", "matplotlib.axes", "matplotlib.colorbar",
          "matplotlib.colors", "mpl_toolkits", "mpl_toolkits.mplot3d", "mpl_toolkits.mplot3d.art3d"):
    sys.modules.setdefault(m, MagicMock())
sys.path.insert(0, "/root/reference")
os.environ.setdefault("TQDM_DISABLE", "1")

import deepmimo as dm  # noqa: E402  (the reference)

from cases import case_list, case_data, oracle_kwargs, params_dict  # noqa: E402
from oracle import channel_oracle as orc  # noqa: E402


def run_reference(c, data):
    ds = dm.Dataset({k: v.copy() for k, v in data.items()})
    if c["bs_fov"] is not None or c["ue_fov"] is not None:
        kw = {}
        if c["bs_fov"] is not None:
            kw["bs_fov"] = c["bs_fov"]
        if c["ue_fov"] is not None:
            kw["ue_fov"] = c["ue_fov"]
        ds.apply_fov(**kw)
    p = dm.ChannelGenParameters(params_dict(c))
    H = ds.compute_channels(p)
    fov = ds["_fov_mask"]
    valid = ~np.isnan(ds["_power_linear_ant_gain"][:, :c["num_paths"]])
    return H, fov, valid


def rel_fro(a, b):
    a = a.reshape(a.shape[0], -1).astype(np.complex128)
    b = b.reshape(b.shape[0], -1).astype(np.complex128)
    num = np.linalg.norm(a - b, axis=1)
    den = np.linalg.norm(b, axis=1)
    return np.where(den > 0, num / np.where(den > 0, den, 1), num)


def main():
    manifest = {"numpy": np.__version__, "reference": dm.__version__, "cases": {}}
    for c in case_list():
        data = case_data(c)
        H, fov, valid = run_reference(c, data)
        o = orc.compute_channels(data, **oracle_kwargs(c))
        err = float(rel_fro(o["H"], H).max()) if H.shape[0] else 0.0
        fov_equal = (fov is None and o["fov_mask"] is None) or bool(np.array_equal(fov, o["fov_mask"]))
        valid_equal = bool(np.array_equal(valid, o["valid"]))
        bits_equal = bool(np.array_equal(H.view(np.uint32), o["H"].view(np.uint32)))
        out = dict(H=H, valid=valid, has_fov=np.array(fov is not None))
        if fov is not None:
            out["fov_mask"] = fov
        np.savez_compressed(os.path.join(HERE, f"{c['name']}.npz"), **out)
        manifest["cases"][c["name"]] = dict(shape=list(H.shape), oracle_max_rel_fro=err, oracle_bits_equal=bits_equal,
                                            fov_mask_equal=fov_equal, valid_equal=valid_equal,
                                            nonzero_users=int((np.abs(H).reshape(H.shape[0], -1).sum(1) > 0).sum()))
        print(f"{c['name']:28s} H{tuple(H.shape)} oracle-vs-ref max rel {err:.2e} bits_equal={bits_equal} "
              f"fov_equal={fov_equal} valid_equal={valid_equal}")
        assert fov_equal and valid_equal and err < 1e-12, c["name"]
    with open(os.path.join(HERE, "manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1)


if __name__ == "__main__":
    main()
