"""Golden vectors for row a11 (Doppler) from the LIVE v3 generator that ships inside the reference tree.

    python tests/golden/make_golden_doppler_v3.py        # build container only (needs /root/reference)

v4 (deepmimo/generator/channel.py:50) only carries the `enable_doppler` flag; the one Doppler definition in the tree is v3's
`OFDM_PathGenerator.generate` (deepmimo_v3/generator/python/construct_deepmimo.py:267-280): without the LPF every path gain is
multiplied by the constant phase exp(-j 2 pi f_c (v tau / c + a tau^2 / (2 c))), tau = ToA, v/a = per-path `Doppler_vel` /
`Doppler_acc` of a dynamic scenario.  This script runs that code (unmodified, imported from /root/reference) on seeded synthetic
ray data and stores its channel; tests/test_doppler_v3.py feeds the same paths to the oracle and to the CUDA path through
`enable_doppler=1` (which maps v3's phase onto the per-path Doppler shift of the time-axis kernels with a single snapshot,
f_D * t == -f_c (v tau / c + a tau^2 / 2c)) and compares.

v3 and v4 are bit-identical on the no-Doppler part for isotropic elements, so the difference measured by the test is
the Doppler factor alone.  Inputs are regenerated from the seed by `doppler_case()`; only v3's outputs are stored.
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

CARRIER_HZ = 3.5e9
LIGHTSPEED = 299792458          # deepmimo_v3/consts.py:112


def doppler_case(n=48, seed=31):
    """Seeded inputs: v4-style path matrices (float32 [n,25], NaN padded) + per-path radial velocity / acceleration."""
    from deepmimo_b200.synth import make_paths
    d = make_paths(n, seed, n_sc=256, bandwidth=10e6)
    rng = np.random.default_rng(seed + 1000)
    pad = np.isnan(d["power"])
    vel = rng.uniform(-30, 30, d["power"].shape).astype(np.float32)
    acc = rng.uniform(-5, 5, d["power"].shape).astype(np.float32)
    vel[pad] = np.nan
    acc[pad] = np.nan
    d["doppler_vel"], d["doppler_acc"] = vel, acc
    cfg = dict(bs_shape=np.array([8, 2]), ue_shape=np.array([2, 1]), bs_rot=np.array([30, 40, 30]),
               ue_rot=np.random.default_rng(seed + 2000).uniform(0, 45, (n, 3)), n_sc=256, sel=np.arange(0, 256, 5),
               bandwidth=10e6, carrier_hz=CARRIER_HZ)
    return d, cfg


def main():
    for m in ("matplotlib", "matplotlib.pyplot", "matplotlib.figure", "matplotlib.axes", "matplotlib.colorbar",
              "matplotlib.colors", "mpl_toolkits", "mpl_toolkits.mplot3d", "mpl_toolkits.mplot3d.art3d"):
        sys.modules.setdefault(m, MagicMock())
    sys.path.insert(0, "/root/reference")
    os.environ.setdefault("TQDM_DISABLE", "1")
    from deepmimo_v3.generator.python.construct_deepmimo import generate_MIMO_channel      # the v3 reference, unmodified

    d, cfg = doppler_case()
    n = d["power"].shape[0]
    out = {}
    for name, dop in (("H_doppler", 1), ("H_static", 0)):
        raydata = []
        for i in range(n):
            m = ~np.isnan(d["power"][i])
            raydata.append({"num_paths": int(m.sum()),
                            "DoD_theta": d["aod_el"][i, m].copy(), "DoD_phi": d["aod_az"][i, m].copy(),
                            "DoA_theta": d["aoa_el"][i, m].copy(), "DoA_phi": d["aoa_az"][i, m].copy(),
                            "phase": d["phase"][i, m].copy(), "ToA": d["delay"][i, m].copy(),
                            "power": (10 ** (d["power"][i, m] / 10)).copy(),        # v3 stores linear power; v4: generator_utils.py:35
                            "LoS": np.zeros(int(m.sum()), dtype=np.int8),
                            "Doppler_vel": d["doppler_vel"][i, m].copy(), "Doppler_acc": d["doppler_acc"][i, m].copy()})
        params = {"ofdm": {"bandwidth": cfg["bandwidth"] / 1e9, "subcarriers": cfg["n_sc"], "selected_subcarriers": cfg["sel"],
                           "rx_filter": 0},
                  "freq_domain": 1, "num_paths": 25, "enable_doppler": dop,
                  "scenario_params": {"doppler_available": 1, "carrier_freq": cfg["carrier_hz"]}}
        tx = {"shape": cfg["bs_shape"], "spacing": 0.5, "rotation": cfg["bs_rot"], "radiation_pattern": "isotropic",
              "fov": np.array([360, 180])}
        rx = {"shape": cfg["ue_shape"], "spacing": 0.5, "rotation": cfg["ue_rot"], "radiation_pattern": "isotropic",
              "fov": np.array([360, 180])}
        H, _los = generate_MIMO_channel(raydata, params, tx, rx)
        out[name] = H
    assert not np.array_equal(out["H_doppler"], out["H_static"])
    np.savez_compressed(os.path.join(HERE, "doppler_v3.npz"), **out)
    print("doppler_v3.npz:", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
