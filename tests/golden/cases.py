"""Golden-case definitions shared by make_golden.py (runs the live reference) and the tests.

Parameter grid mined from the reference's own correspondence tests
(test/test_v3_correspondence.py:21-34,65-76: num_paths 5/10/25, subcarriers 64/512,
selected_subcarriers arange(1) / arange(3)*3, UE shape [1,1]/[3,2], freq_domain T/F,
BS rotation None/[30,40,30]/per-user, bs_fov [140,120], ue_fov [90,80]) and from
test/test_fov.py:74-155 (FoV [360,180], [180,90]), docs/manual.ipynb cell 85
(rotation [0,30,-135]).
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from deepmimo_b200.synth import make_paths  # noqa: E402


def _interleave_nans(d: dict, seed: int) -> dict:
    """Knock out a few non-trailing paths (all seven matrices together)."""
    rng = np.random.default_rng(seed)
    hole = rng.random(d["power"].shape) < 0.15
    for k in ("power", "phase", "delay", "aoa_az", "aoa_el", "aod_az", "aod_el", "inter"):
        d[k] = d[k].copy()
        d[k][hole] = np.nan
    return d


def case_list():
    """Each case: name, data kwargs, channel params, FoV.  Small enough to commit (< 1 MB each)."""
    A = np.array
    cases = []

    def add(name, n, seed, n_sc, bw, sel, bs_shape=(8, 1), ue_shape=(1, 1), bs_rot=(0, 0, 0), ue_rot=(0, 0, 0),
            bs_pat="isotropic", ue_pat="isotropic", bs_fov=None, ue_fov=None, num_paths=25, fd=1,
            bs_sp=0.5, ue_sp=0.5, holes=False, n_cols=25, zero_frac=0.10, lpf=0, f64=False):
        cases.append(dict(name=name, n=n, seed=seed, n_sc=n_sc, bw=bw, sel=np.asarray(sel), bs_shape=A(bs_shape), f64=f64,
                          ue_shape=A(ue_shape), bs_rot=np.asarray(bs_rot), ue_rot=np.asarray(ue_rot), bs_pat=bs_pat,
                          ue_pat=ue_pat, bs_fov=None if bs_fov is None else A(bs_fov),
                          ue_fov=None if ue_fov is None else A(ue_fov), num_paths=num_paths, fd=fd,
                          bs_sp=bs_sp, ue_sp=ue_sp, holes=holes, n_cols=n_cols, zero_frac=zero_frac, lpf=lpf))

    per_user = lambda n, s, lo=0, hi=45: np.random.default_rng(s).uniform(lo, hi, (n, 3))

    add("defaults", 64, 11, 512, 10e6, np.arange(1))
    add("cfg1_shape", 64, 12, 64, 10e6, np.arange(64))
    add("cfg2_shape", 6, 13, 512, 50e6, np.r_[np.arange(0, 512, 37), 511], bs_shape=(32, 8), ue_shape=(2, 2),
        bs_rot=(30, 40, 30), ue_rot=per_user(6, 42))
    add("cfg3_shape", 64, 14, 1024, 100e6, np.r_[np.arange(0, 1024, 128), 1023], bs_shape=(64, 4),
        bs_rot=(0, 30, -135), bs_pat="halfwave-dipole", ue_pat="halfwave-dipole", bs_fov=(140, 120), ue_fov=(90, 80))
    add("td_basic", 32, 15, 512, 10e6, np.arange(1), bs_shape=(8, 4), ue_shape=(2, 1), fd=0)
    add("td_fov_np10", 128, 16, 512, 10e6, np.arange(1), bs_shape=(4, 2), ue_shape=(3, 2), fd=0, num_paths=10,
        bs_rot=(30, 40, 30), ue_rot=per_user(128, 43), bs_fov=(140, 120), ue_fov=(90, 80))
    add("sel_stride_np5", 64, 17, 64, 50e6, np.arange(3) * 3, ue_shape=(3, 2), num_paths=5)
    add("ue_rot_random", 64, 18, 512, 10e6, np.arange(8), ue_shape=(2, 2), ue_rot=A([[0, 30], [-20, 20], [0, 90]]))
    add("mixed_pattern_bsfov", 64, 19, 512, 20e6, np.arange(0, 512, 64), bs_shape=(8, 2), ue_shape=(2, 1),
        bs_pat="halfwave-dipole", bs_fov=(180, 90), ue_fov=(360, 180), bs_rot=(10, 20, 30))
    add("npow2_spacing", 48, 20, 600, 30e6, A([0, 1, 7, 299, 599]), bs_shape=(6, 3), ue_shape=(1, 2), bs_sp=0.7,
        ue_sp=0.35, bs_rot=(-15, 5, 170))
    add("holes_fd", 64, 21, 256, 10e6, np.arange(0, 256, 16), bs_shape=(4, 4), holes=True, ue_fov=(120, 90),
        bs_fov=(360, 180))
    add("holes_td", 64, 22, 256, 10e6, np.arange(1), bs_shape=(4, 1), ue_shape=(2, 2), fd=0, holes=True,
        ue_pat="halfwave-dipole")
    add("all_empty_and_single", 1, 23, 512, 10e6, np.arange(4), zero_frac=1.0)
    add("ue_rot_tiled_fov_full", 32, 24, 128, 10e6, np.arange(128), ue_shape=(2, 2), ue_rot=(30, 45, 60),
        bs_fov=(360, 180), ue_fov=(360, 180))
    add("np10_fd_perusr_bsrot", 32, 25, 512, 10e6, np.arange(0, 512, 100), bs_shape=(16, 1), num_paths=10,
        bs_rot=(0, 0, 90), ue_rot=per_user(32, 44, -90, 90), ue_shape=(2, 2), bs_fov=(120, 90), ue_fov=(180, 90))
    # receive low-pass filter (ofdm.rx_filter = 1, channel.py:193-194): power-of-two N (FFT route on the device), a
    # non-power-of-two N with a strided selection (direct DFT route), and dipole + FoV (float64 power branch)
    add("lpf_pow2", 24, 26, 256, 20e6, np.arange(256), bs_shape=(4, 2), ue_shape=(2, 1), bs_rot=(10, 20, 30), lpf=1)
    add("lpf_n600_sel", 24, 27, 600, 30e6, A([0, 1, 7, 64, 299, 598, 599]), bs_shape=(6, 1), ue_shape=(1, 2), lpf=1, num_paths=10)
    add("lpf_dipole_fov", 32, 28, 128, 10e6, np.arange(0, 128, 3), bs_shape=(8, 1), bs_pat="halfwave-dipole",
        ue_pat="halfwave-dipole", bs_fov=(140, 120), ue_fov=(180, 120), holes=True, lpf=1)
    # float64 path matrices (last paragraph): NumPy runs every step in float64; the values are NOT
    # float32-representable, so a float32 flow would miss the reference by ~1e-3 (delay x frequency reaches thousands of cycles)
    add("f64_fd_rot", 48, 29, 4096, 400e6, np.r_[np.arange(0, 4096, 293), 4095], bs_shape=(16, 4), ue_shape=(2, 1), bs_rot=(30, 40, 30),
        ue_rot=per_user(48, 45), f64=True)
    add("f64_small_dipole_fov", 64, 30, 64, 10e6, np.arange(64), bs_shape=(8, 1), bs_pat="halfwave-dipole", ue_pat="halfwave-dipole",
        bs_fov=(140, 120), ue_fov=(180, 120), holes=True, f64=True)
    add("f64_td", 32, 31, 512, 10e6, np.arange(1), bs_shape=(4, 2), ue_shape=(2, 2), fd=0, num_paths=12, bs_rot=(5, -10, 100), f64=True)
    add("f64_lpf", 16, 32, 128, 10e6, np.arange(0, 128, 5), bs_shape=(4, 1), lpf=1, f64=True)
    return cases


def case_data(c: dict) -> dict:
    d = _case_data32(c)
    if c.get("f64"):
        rng = np.random.default_rng(c["seed"] + 5000)
        for k in ("power", "phase", "delay", "aoa_az", "aoa_el", "aod_az", "aod_el"):
            d[k] = d[k].astype(np.float64) * (1.0 + 1e-6 * rng.standard_normal(d[k].shape))      # NaN padding stays NaN
    return d


def _case_data32(c: dict) -> dict:
    d = make_paths(c["n"], c["seed"], n_sc=c["n_sc"], bandwidth=c["bw"], n_cols=c["n_cols"], zero_frac=c["zero_frac"],
                   clip_frac=0.02)
    if c["holes"]:
        d = _interleave_nans(d, c["seed"] + 1000)
    return d


def oracle_kwargs(c: dict) -> dict:
    return dict(bs_shape=c["bs_shape"], ue_shape=c["ue_shape"], bs_spacing=c["bs_sp"], ue_spacing=c["ue_sp"],
                bs_rotation=c["bs_rot"], ue_rotation=c["ue_rot"], bs_pattern=c["bs_pat"], ue_pattern=c["ue_pat"],
                bs_fov=c["bs_fov"], ue_fov=c["ue_fov"], num_paths=c["num_paths"], freq_domain=bool(c["fd"]),
                subcarriers=c["n_sc"], selected_subcarriers=c["sel"], bandwidth=c["bw"], rx_filter=c.get("lpf", 0))


def params_dict(c: dict) -> dict:
    """Nested dict in the reference's ChannelGenParameters layout (deepmimo/generator/channel.py:33-63)."""
    return {
        "bs_antenna": {"shape": c["bs_shape"], "spacing": c["bs_sp"], "rotation": c["bs_rot"],
                       "radiation_pattern": c["bs_pat"]},
        "ue_antenna": {"shape": c["ue_shape"], "spacing": c["ue_sp"], "rotation": c["ue_rot"],
                       "radiation_pattern": c["ue_pat"]},
        "enable_doppler": 0, "enable_dual_polar": 0, "num_paths": c["num_paths"], "freq_domain": c["fd"],
        "ofdm": {"subcarriers": c["n_sc"], "selected_subcarriers": c["sel"], "bandwidth": c["bw"], "rx_filter": c.get("lpf", 0)},
    }
