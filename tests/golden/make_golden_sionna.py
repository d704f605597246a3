"""Golden (a, tau) arrays from the LIVE reference Sionna adapter (deepmimo/integrations/sionna_adapter.py:21-200).

    python tests/golden/make_golden_sionna.py        # build container only (needs /root/reference)

The adapter consumes a v3-style dataset dict (`dataset[bs]['user']['channel'][ue]`, `['paths'][ue]['num_paths' | 'ToA']`).  That dict
is assembled here from the live v4 reference: `channel` is `dm.Dataset.compute_channels` with freq_domain = 0 (time domain),
`num_paths` / `ToA` are the valid paths of each user in column order (the slots the time-domain channel fills, channel.py:285-287).
The unmodified adapter class then produces the samples; tests/test_sionna_adapter.py replays them against the CUDA path.
Inputs are regenerated from the seeds by `sionna_case()`; only the adapter's outputs are stored.
"""
import os
import sys
from unittest.mock import MagicMock

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, HERE):
    if p not in sys.path:
        sys.path.insert(0, p)


def sionna_case():
    from deepmimo_b200.synth import make_paths
    n = 40
    data = [make_paths(n, 61 + b, n_sc=256, bandwidth=10e6, zero_frac=0.2) for b in range(2)]
    params = {"bs_antenna": {"shape": np.array([4, 2]), "spacing": 0.5, "rotation": np.array([10, 20, 30]), "radiation_pattern": "isotropic"},
              "ue_antenna": {"shape": np.array([2, 1]), "spacing": 0.5, "rotation": np.array([0, 0, 0]), "radiation_pattern": "isotropic"},
              "enable_doppler": 0, "enable_dual_polar": 0, "num_paths": 10, "freq_domain": 0,
              "ofdm": {"subcarriers": 256, "selected_subcarriers": np.arange(1), "bandwidth": 10e6, "rx_filter": 0}}
    ue_idx = np.array([[0, 1], [2, 3], [5, 7], [39, 11]])      # 4 samples x 2 receivers
    bs_idx = np.array([[0, 1], [1, 0]])                        # 2 samples x 2 transmitters
    return data, params, bs_idx, ue_idx


def main():
    for m in ("matplotlib", "matplotlib.pyplot", "matplotlib.figure", "matplotlib.axes", "matplotlib.colorbar",
              "matplotlib.colors", "mpl_toolkits", "mpl_toolkits.mplot3d", "mpl_toolkits.mplot3d.art3d"):
        sys.modules.setdefault(m, MagicMock())
    sys.path.insert(0, "/root/reference")
    os.environ.setdefault("TQDM_DISABLE", "1")
    import deepmimo as dm                                                        # the reference, unmodified
    from deepmimo.integrations.sionna_adapter import DeepMIMOSionnaAdapter

    data, params, bs_idx, ue_idx = sionna_case()
    v3 = []
    for d in data:
        ds = dm.Dataset({k: v.copy() for k, v in d.items()})
        H = ds.compute_channels(dm.ChannelGenParameters(params))                 # time domain [n, M_r, M_t, num_paths]
        valid = ~np.isnan(d["power"][:, :params["num_paths"]])
        paths = [{"num_paths": int(valid[i].sum()), "ToA": d["delay"][i, :params["num_paths"]][valid[i]]} for i in range(len(valid))]
        v3.append({"user": {"channel": H, "paths": paths}})
    out = {}
    for name, kw in (("multi", dict(bs_idx=bs_idx, ue_idx=ue_idx)), ("default", dict())):
        ad = DeepMIMOSionnaAdapter(v3, **kw)
        samples = list(ad())
        out[f"{name}_a"] = np.stack([s[0] for s in samples])
        out[f"{name}_tau"] = np.stack([s[1] for s in samples])
        out[f"{name}_len"] = np.array(len(ad))
    np.savez_compressed(os.path.join(HERE, "sionna.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
