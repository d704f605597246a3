"""Synthetic ray-tracing scenarios of the BASELINE.json shapes.

A DeepMIMO `Dataset` is a dict of float32 `[n_ue, 25]` path matrices with valid
paths leading and NaN padding trailing (deepmimo/converter/wireless_insite/
p2m_parser.py:84-123, deepmimo/consts.py:65,180).  These generators produce such
dicts with fixed seeds so that the oracle, the CUDA path and the reference see
the same bytes.  No scenario files are needed.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional

import numpy as np

MAX_PATHS = 25          # deepmimo/consts.py:180
C_LIGHT = 299792458.0


def make_paths(n_ue: int, seed: int, *, n_sc: int = 512, bandwidth: float = 10e6, n_cols: int = MAX_PATHS,
               clip_frac: float = 0.005, zero_frac: float = 0.10, dense: bool = False) -> dict:
    """Path matrices for `n_ue` users (SURVEY.md 8d "Synthetic inputs").

    n_paths: `zero_frac` users with 0 paths, else U{1..n_cols}; power dBW U(-160,-60) sorted
    descending; phase deg U(-180,180); delay s U(3e-8,4e-6) ascending, `clip_frac` of valid
    paths moved to U(1,1.5)*N/B (exercises the delay clip); el deg U(0,180); az deg U(-180,180).
    `dense`: every user has all `n_cols` paths (the P = 25 case north_star's hypothesis H4 reasons about).
    """
    rng = np.random.default_rng(seed)
    n_paths = rng.integers(1, n_cols + 1, n_ue)
    n_paths[rng.random(n_ue) < zero_frac] = 0
    if dense:
        n_paths[:] = n_cols
    pad = np.arange(n_cols)[None, :] >= n_paths[:, None]

    def fill(lo, hi):
        return rng.uniform(lo, hi, (n_ue, n_cols))

    power = -np.sort(-fill(-160, -60), axis=1)
    delay = np.sort(fill(3e-8, 4e-6), axis=1)
    over = (rng.random((n_ue, n_cols)) < clip_frac)
    delay = np.where(over, rng.uniform(1.0, 1.5, (n_ue, n_cols)) * n_sc / bandwidth, delay)
    out = dict(power=power, phase=fill(-180, 180), delay=delay,
               aoa_az=fill(-180, 180), aoa_el=fill(0, 180), aod_az=fill(-180, 180), aod_el=fill(0, 180))
    for k, v in out.items():
        v = v.astype(np.float32)
        v[pad] = np.nan
        out[k] = np.ascontiguousarray(v)
    out["rx_pos"] = rng.uniform(0, 500, (n_ue, 3)).astype(np.float32)
    out["tx_pos"] = np.zeros((1, 3), np.float32)
    inter = np.zeros((n_ue, n_cols), np.float32)
    inter[pad] = np.nan
    out["inter"] = inter
    return out


@dataclass
class Scenario:
    """One benchmark / parity configuration: path data + channel parameters + FoV."""
    name: str
    data: dict
    params: dict                      # nested dict in ChannelGenParameters layout
    bs_fov: Optional[np.ndarray] = None
    ue_fov: Optional[np.ndarray] = None
    doppler_hz: Optional[np.ndarray] = None
    times: Optional[np.ndarray] = None
    carrier_hz: float = 3.5e9
    notes: str = ""
    extra: dict = field(default_factory=dict)

    @property
    def n_ue(self) -> int:
        return self.data["power"].shape[0]


def _params(bs_shape, ue_shape, n_sc, n_sel, bandwidth, *, bs_rot=(0, 0, 0), ue_rot=(0, 0, 0),
            bs_pat="isotropic", ue_pat="isotropic", freq_domain=1, num_paths=MAX_PATHS) -> dict:
    return {
        "bs_antenna": {"shape": np.array(bs_shape), "spacing": 0.5, "rotation": np.asarray(bs_rot),
                       "radiation_pattern": bs_pat},
        "ue_antenna": {"shape": np.array(ue_shape), "spacing": 0.5, "rotation": np.asarray(ue_rot),
                       "radiation_pattern": ue_pat},
        "enable_doppler": 0, "enable_dual_polar": 0, "num_paths": num_paths, "freq_domain": freq_domain,
        "ofdm": {"subcarriers": n_sc, "selected_subcarriers": np.arange(n_sel), "bandwidth": bandwidth,
                 "rx_filter": 0},
    }


def doppler_from_velocity(data: dict, seed: int, carrier_hz: float, vmax: float = 30.0) -> np.ndarray:
    """Per-path Doppler shift f_D = (f_c/c) v . u_aoa (row a11 definition), float32 [n,P], NaN padded."""
    rng = np.random.default_rng(seed)
    n = data["power"].shape[0]
    speed = rng.uniform(0, vmax, n)
    heading = rng.uniform(-np.pi, np.pi, n)
    v = np.stack([speed * np.cos(heading), speed * np.sin(heading), np.zeros(n)], axis=1)
    th = np.deg2rad(data["aoa_el"].astype(np.float64))
    ph = np.deg2rad(data["aoa_az"].astype(np.float64))
    u = np.stack([np.sin(th) * np.cos(ph), np.sin(th) * np.sin(ph), np.cos(th)], axis=-1)   # [n,P,3]
    fd = carrier_hz / C_LIGHT * np.einsum("npk,nk->np", u, v)
    return fd.astype(np.float32)


def scenario(cfg: int, n_ue: Optional[int] = None, *, bs_index: int = 0, shard: int = 0, dense: bool = False,
             fov: bool = True) -> Scenario:
    """The five BASELINE.json configurations, optionally with fewer users.

    `shard` offsets every seed by 100*shard (distinct data per rank in weak-scaling runs); `bs_index`
    selects the base station of config 5 (seed 1005 + 10*b).  Sensitivity variants for the benchmark: `dense` gives every user
    all 25 paths, `fov=False` drops the field-of-view filter of config 3 (its dipole patterns stay)."""
    so = 100 * shard
    s = _scenario(cfg, n_ue, bs_index, so, dense)
    if not fov:
        s.bs_fov = s.ue_fov = None
        s.name += "_nofov"
        s.notes += " -- FoV filter off"
    if dense:
        s.name += "_dense"
        s.notes += " -- dense: 25 valid paths for every user"
    return s


def _scenario(cfg: int, n_ue: Optional[int], bs_index: int, so: int, dense: bool) -> Scenario:
    if cfg == 1:
        n = 80_000 if n_ue is None else n_ue
        d = make_paths(n, 1001 + so, n_sc=64, bandwidth=10e6, dense=dense)
        return Scenario("cfg1_asu_8x1_K64", d, _params([8, 1], [1, 1], 64, 64, 10e6),
                        notes="1 BS, 8x1 ULA, 1 UE antenna, N=K=64, B=10 MHz, FD, isotropic")
    if cfg == 2:
        n = 4096 if n_ue is None else n_ue
        d = make_paths(n, 1002 + so, n_sc=512, bandwidth=50e6, dense=dense)
        ue_rot = np.random.default_rng(42 + so).uniform(0, 45, (n, 3))
        return Scenario("cfg2_32x8_2x2_K512", d,
                        _params([32, 8], [2, 2], 512, 512, 50e6, bs_rot=[30, 40, 30], ue_rot=ue_rot),
                        notes="32x8 rotated BS UPA + 2x2 UE (per-user rotation), N=K=512, B=50 MHz, 3.5 GHz, isotropic")
    if cfg == 3:
        n = 8192 if n_ue is None else n_ue
        d = make_paths(n, 1003 + so, n_sc=1024, bandwidth=100e6, dense=dense)
        return Scenario("cfg3_64x4_dipole_fov_K1024", d,
                        _params([64, 4], [1, 1], 1024, 1024, 100e6, bs_rot=[0, 30, -135],
                                bs_pat="halfwave-dipole", ue_pat="halfwave-dipole"),
                        bs_fov=np.array([140, 120]), ue_fov=np.array([90, 80]),
                        notes="64x4 UPA, half-wave dipole, BS FoV [140,120], UE FoV [90,80], N=K=1024, B=100 MHz")
    if cfg == 4:
        n = 50_000 if n_ue is None else n_ue
        d = make_paths(n, 1004 + so, n_sc=512, bandwidth=10e6, dense=dense)
        fd = doppler_from_velocity(d, 2004 + so, 3.5e9)
        return Scenario("cfg4_td_doppler_T16", d,
                        _params([8, 4], [2, 1], 512, 1, 10e6, freq_domain=0),
                        doppler_hz=fd, times=np.arange(16) * 1e-3,
                        notes="time domain, 8x4 BS, 2x1 UE, 25 path slots, 16 snapshots of 1 ms, Doppler from UE velocity")
    if cfg == 5:
        n = 200_000 if n_ue is None else n_ue
        d = make_paths(n, 1005 + 10 * bs_index + so, n_sc=1024, bandwidth=100e6, dense=dense)
        return Scenario(f"cfg5_city_bs{bs_index}_8x8_K1024", d, _params([8, 8], [1, 1], 1024, 1024, 100e6),
                        notes="city-scale shard: one BS x 200k users, 8x8 BS UPA, 1 UE antenna, N=K=1024, B=100 MHz")
    if cfg == 6:
        n = 131_072 if n_ue is None else n_ue
        d = make_paths(n, 1006 + so, n_sc=64, bandwidth=10e6, dense=dense)
        return Scenario("mid_8x8_K64", d, _params([8, 8], [1, 1], 64, 64, 10e6, bs_rot=[5, 10, 15]),
                        notes="not a BASELINE config: the common 64 antennas x 64 subcarriers dataset (32 KB per user), 8x8 rotated BS UPA, "
                              "1 UE antenna, N=K=64, B=10 MHz (no path beyond the OFDM symbol)")
    if cfg == 7:
        n = 200_000 if n_ue is None else n_ue
        d = make_paths(n, 1007 + so, n_sc=512, bandwidth=10e6, dense=dense)
        return Scenario("default_8x8_K1", d, _params([8, 8], [1, 1], 512, 1, 10e6),
                        notes="not a BASELINE config: the reference's default OFDM parameters (512 subcarriers, ONE selected, channel.py:58-62) "
                              "on an 8x8 BS panel: 512 bytes of output per user")
    if cfg == 8:
        n = 200_000 if n_ue is None else n_ue
        d = make_paths(n, 1008 + so, n_sc=512, bandwidth=10e6, dense=dense)
        return Scenario("td_8x8_static", d, _params([8, 8], [1, 1], 512, 1, 10e6, freq_domain=0),
                        notes="not a BASELINE config: the reference's own time-domain mode (freq_domain = 0, no time axis), 8x8 BS panel, "
                              "25 path slots: 12.8 KB of output per user")
    raise ValueError(f"unknown config {cfg}")


def coef_count(s: Scenario) -> int:
    """Complex64 coefficients in the returned array for a scenario."""
    p = s.params
    m = int(np.prod(p["bs_antenna"]["shape"][:2])) * int(np.prod(p["ue_antenna"]["shape"][:2]))
    last = len(p["ofdm"]["selected_subcarriers"]) if p["freq_domain"] else min(p["num_paths"], s.data["power"].shape[1])
    t = 1 if s.times is None else len(s.times)
    return s.n_ue * m * last * t
