"""oracle/channel_oracle.py -- TEST INFRASTRUCTURE (the oracle). NOT product code.

CPU / NumPy restatement of the channel-generation path of jmoraispk/DeepMIMO
v4.0.0a3 (``dataset.compute_channels(params)``).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py`` (``cpu_baseline`` leg and
``--impl reference``) may import this file; ``deepmimo_b200`` never does.

Pinning: the restatement is checked bit-for-bit (masks) / to <= 1e-12 (values)
against the *live* reference imported from /root/reference in the build
container (``tests/golden/make_golden.py`` writes the fixtures under
``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` replays them).  The
reference's own tests store no golden vectors for this path, so
the reference-generated fixtures are the pin.  Row a11 (Doppler / time
snapshots) has no implementation in v4; the oracle's `doppler_hz` / `times`
extension is pinned against the one Doppler definition in the reference tree,
v3's constant per-path phase (deepmimo_v3/.../construct_deepmimo.py:267-280), by
tests/golden/doppler_v3.npz (produced by running v3 itself) and against the v4
reference at t = 0; the multi-snapshot time axis beyond that is this repo's
definition (exp(+j 2 pi f_D t) per path).

Every function cites the reference file:line it follows (paths relative to
/root/reference).  dtype flow is deliberately the reference's (float32 inputs,
NumPy >= 2 promotion rules): see SURVEY.md Appendix A.
"""
from __future__ import annotations

import numpy as np

MAX_GAIN_DIPOLE = 1.643  # deepmimo/generator/ant_patterns.py:51


# --------------------------------------------------------------------------
# geometry
# --------------------------------------------------------------------------
def element_grid(shape) -> np.ndarray:
    """Antenna element indices (x, y, z) of a panel, y fastest, x == 0.

    Follows deepmimo/generator/geometry.py:105-120 (`_ant_indices`): element
    n sits at (0, n mod shape[0], n div shape[0]); only shape[0], shape[1] are read.
    """
    m_h, m_v = int(shape[0]), int(shape[1])
    n = np.arange(m_h * m_v)
    return np.stack([np.zeros_like(n), n % m_h, n // m_h], axis=1)


def rotate_angles(rotation_deg: np.ndarray, el_deg: np.ndarray, az_deg: np.ndarray):
    """Global -> array-local angles.  deepmimo/generator/geometry.py:244-319.

    rotation_deg: [3] or [n,3] = rotations about x, y, z in degrees.
    el_deg/az_deg: float32 [n,P] degrees (zenith / azimuth).
    Returns theta', phi' in radians (float64 for float32 inputs and a float64/int
    rotation), NaN preserved.
    """
    rot = np.asarray(rotation_deg)
    if rot.ndim == 1:
        rot = rot[None, :]
    elif rot.ndim == 3:
        rot = rot.reshape(-1, 3)
    theta = np.deg2rad(el_deg)          # :284  float32 stays float32
    phi = np.deg2rad(az_deg)            # :285
    rot = np.deg2rad(rot)               # :286  int/float64 -> float64
    rx, ry, rz = rot[:, 0:1], rot[:, 1:2], rot[:, 2:3]
    s_dphi = np.sin(phi - rz)           # :294  float32 - float64 -> float64
    c_dphi = np.cos(phi - rz)           # :297
    s_y, c_y = np.sin(ry), np.cos(ry)   # :295, :298
    s_x, c_x = np.sin(rx), np.cos(rx)   # :296, :299
    s_t = np.sin(theta)                 # :301  float32 SIMD sin (rounding point R2)
    c_t = np.cos(theta)                 # :302
    # :305-306 (same association order as the reference expression)
    theta_rot = np.arccos(c_y * c_x * c_t + s_t * (s_y * c_x * c_dphi - s_x * s_dphi))
    # :308-310
    phi_rot = np.angle(c_y * s_t * c_dphi - s_y * c_t
                       + 1j * (c_y * s_x * c_t + s_t * (s_y * s_x * c_dphi + c_x * s_dphi)))
    return theta_rot, phi_rot


def fov_inclusion(fov_deg, theta: np.ndarray, phi: np.ndarray) -> np.ndarray:
    """Boolean in-FoV mask.  deepmimo/generator/geometry.py:162-195."""
    theta = np.mod(theta, 2 * np.pi)    # :180
    phi = np.mod(phi, 2 * np.pi)        # :181
    fov = np.deg2rad(fov_deg)           # :184
    in_h = np.logical_or(phi <= 0 + fov[0] / 2, phi >= 2 * np.pi - fov[0] / 2)          # :187
    in_v = np.logical_and(theta <= np.pi / 2 + fov[1] / 2, theta >= np.pi / 2 - fov[1] / 2)  # :190
    return np.logical_and(in_h, in_v)   # :193


def is_full_fov(fov) -> bool:
    """deepmimo/generator/dataset.py:450-459."""
    return bool(fov[0] >= 360 and fov[1] >= 180)


def steering_batch(grid: np.ndarray, theta: np.ndarray, phi: np.ndarray, kd: float) -> np.ndarray:
    """Array response [n, M, P] complex128, exact zeros where theta is NaN.

    deepmimo/generator/geometry.py:38-102 (`_array_response_batch`,
    `_array_response_phase`).
    """
    n, p = theta.shape
    ok = ~np.isnan(theta)                                   # :65
    th, ph = theta[ok], phi[ok]
    g = np.vstack([1j * kd * np.sin(th) * np.cos(ph),       # :99
                   1j * kd * np.sin(th) * np.sin(ph),       # :100
                   1j * kd * np.cos(th)]).T                 # :101-102  [nv,3]
    out = np.zeros((n, len(grid), p), dtype=np.complex128)  # :71
    bi, pi_ = np.nonzero(ok)                                # :74
    out[bi, :, pi_] = np.exp(grid @ g.T).T                  # :77-80
    return out


# --------------------------------------------------------------------------
# element patterns
# --------------------------------------------------------------------------
def pattern_gain(name: str, theta: np.ndarray):
    """Element power gain.  deepmimo/generator/ant_patterns.py:21-78.

    'isotropic' returns the Python scalar 1. (so float32 power stays float32,
    :31); 'halfwave-dipole' returns float64 1.643*cos^2(pi/2 cos th)/sin th where
    |sin th| > 1e-10, else 0 (NaN theta -> 0) (:51-71).
    """
    if name == "isotropic":
        return 1.
    if name == "halfwave-dipole":
        theta = np.asarray(theta)
        g = np.zeros_like(theta, dtype=np.float64)          # :57
        ok = np.abs(np.sin(theta)) > 1e-10                  # :60
        tv = theta[ok]
        g[ok] = MAX_GAIN_DIPOLE * (np.cos(np.pi / 2 * np.cos(tv)) ** 2 / np.sin(tv))  # :65-69
        return g
    raise NotImplementedError(f"The given '{name}' antenna radiation pattern is not applicable.")  # :119-122


# --------------------------------------------------------------------------
# per-path OFDM gains
# --------------------------------------------------------------------------
def ofdm_path_gains(power, toa, phase, n_sc: int, sel_sc: np.ndarray, ts: float, rx_filter: int = 0):
    """[P_i, K] complex path gains.  deepmimo/generator/channel.py:155-198.

    rx_filter = 0: c_p exp(-j 2 pi k delay_n / N) (:196-197).
    rx_filter = 1: receive low-pass filter (:166-168, :193-194): (c_p sinc(d - delay_n)) @ exp(-j 2 pi d k / N), d = 0..N-1;
    `*` and `@` have equal precedence, so the product with sinc is formed first.
    Returns (gains, over) where `over` marks paths with delay_n >= N (:187).
    """
    power = power.reshape(-1, 1)
    delay_n = toa.reshape(-1, 1) / ts                       # :183  float32 / float32(Ts)  (R11)
    phase = phase.reshape(-1, 1)
    over = delay_n >= n_sc                                  # :187
    power[over] = 0                                         # :188
    delay_n[over] = n_sc                                    # :189
    c = np.sqrt(power / n_sc) * np.exp(1j * np.deg2rad(phase))          # :192
    if rx_filter:
        delay_d = np.arange(n_sc)                                                               # :165
        delay_to_ofdm = np.exp(-1j * 2 * np.pi / n_sc * np.outer(delay_d, sel_sc))              # :166-167
        g = c * np.sinc(delay_d - delay_n) @ delay_to_ofdm                                      # :194
    else:
        g = c * np.exp(-1j * (2 * np.pi / n_sc) * np.outer(delay_n.ravel(), sel_sc))  # :196-197
    return g, over.ravel()


# --------------------------------------------------------------------------
# the path
# --------------------------------------------------------------------------
def resolve_ue_rotation(ue_rot, n_ue: int, seed_numpy: bool = True) -> np.ndarray:
    """UE rotation -> per-user [n,3].  deepmimo/generator/dataset.py:328-338, :250.

    (3,) is tiled; (3,2) draws U(lo,hi) per user from the *global* NumPy RNG after
    np.random.seed(1001) (dataset.py:250); (n,3) is taken as is.
    """
    ue_rot = np.asarray(ue_rot)
    if ue_rot.ndim == 1 and ue_rot.shape[0] == 3:
        return np.tile(ue_rot, (n_ue, 1))
    if ue_rot.ndim == 2 and ue_rot.shape == (3, 2):
        if seed_numpy:
            np.random.seed(1001)
        return np.random.uniform(ue_rot[:, 0], ue_rot[:, 1], (n_ue, 3))
    return ue_rot


def compute_channels(data: dict, *, bs_shape=(8, 1), ue_shape=(1, 1), bs_spacing=0.5, ue_spacing=0.5,
                     bs_rotation=(0, 0, 0), ue_rotation=(0, 0, 0),
                     bs_pattern="isotropic", ue_pattern="isotropic",
                     bs_fov=None, ue_fov=None, num_paths=25, freq_domain=True,
                     subcarriers=512, selected_subcarriers=(0,), bandwidth=10e6, rx_filter=0,
                     doppler_hz=None, times=None, user_range=None) -> dict:
    """Restatement of Dataset.compute_channels (deepmimo/generator/dataset.py:224-268).

    data: dict with float32 [n,P0] arrays power(dBW) phase(deg) delay(s) aoa_az aoa_el
    aod_az aod_el (deg).  Returns dict(H, fov_mask|None, valid, clip, path_slot).

    Extension (row a11, parity unpinned): doppler_hz [n,P0] and times [T] multiply
    each path by exp(+j 2 pi f_D t) and append a trailing T axis.
    """
    sl = slice(None) if user_range is None else slice(*user_range)
    power_db = np.asarray(data["power"])[sl]
    phase = np.asarray(data["phase"])[sl]
    delay = np.asarray(data["delay"])[sl]
    n_ue = power_db.shape[0]
    sel_sc = np.asarray(selected_subcarriers)

    # --- rotated angles: dataset.py:310-356
    ue_rot = resolve_ue_rotation(ue_rotation, np.asarray(data["power"]).shape[0])
    if ue_rot.ndim == 2 and ue_rot.shape[0] != 1:
        ue_rot = ue_rot[sl]
    aod_th, aod_ph = rotate_angles(np.asarray(bs_rotation), np.asarray(data["aod_el"])[sl], np.asarray(data["aod_az"])[sl])
    aoa_th, aoa_ph = rotate_angles(ue_rot, np.asarray(data["aoa_el"])[sl], np.asarray(data["aoa_az"])[sl])

    # --- FoV: dataset.py:461-512
    bs_full = bs_fov is not None and is_full_fov(bs_fov)
    ue_full = ue_fov is not None and is_full_fov(ue_fov)
    if (bs_fov is None and ue_fov is None) or (bs_full and ue_full):
        fov_mask = None                                     # :484-491
    else:
        fov_mask = np.ones_like(aod_th, dtype=bool)         # :494
        if not bs_full:
            # bs_fov may be None here when only ue_fov was given: the reference would raise
            # inside np.deg2rad(None); apply_fov always sets both (dataset.py:447-448).
            fov_mask = np.logical_and(fov_mask, fov_inclusion(bs_fov, aod_th, aod_ph))   # :497-499
        if not ue_full:
            fov_mask = np.logical_and(fov_mask, fov_inclusion(ue_fov, aoa_th, aoa_ph))   # :502-504
        aod_th = np.where(fov_mask, aod_th, np.nan)         # :508-511
        aod_ph = np.where(fov_mask, aod_ph, np.nan)
        aoa_th = np.where(fov_mask, aoa_th, np.nan)
        aoa_ph = np.where(fov_mask, aoa_ph, np.nan)

    # --- array responses and their product: dataset.py:380-417
    a_tx = steering_batch(element_grid(bs_shape), aod_th, aod_ph, 2 * np.pi * bs_spacing)
    a_rx = steering_batch(element_grid(ue_shape), aoa_th, aoa_ph, 2 * np.pi * ue_spacing)

    # --- power: dataset.py:694-696 (generator_utils.py:35), :665-691, ant_patterns.py:167-168
    p_lin = 10 ** (power_db / 10)
    pw = p_lin * (pattern_gain(bs_pattern, aod_th) * pattern_gain(ue_pattern, aoa_th))

    # --- [:num_paths] slicing: dataset.py:255-262
    P = min(int(num_paths), pw.shape[1])
    a_tx, a_rx = a_tx[..., :P], a_rx[..., :P]
    pw, delay, phase = pw[:, :P], delay[:, :P], phase[:, :P]
    dop = None if doppler_hz is None else np.asarray(doppler_hz)[sl][:, :P]
    tt = None if times is None else np.atleast_1d(np.asarray(times, dtype=np.float64))

    # --- accumulation: channel.py:200-289
    ts = 1 / bandwidth                                      # :223
    m_rx, m_tx = a_rx.shape[1], a_tx.shape[1]
    last = len(sel_sc) if freq_domain else P                # :256
    shape = (n_ue, m_rx, m_tx, last) + (() if tt is None else (len(tt),))
    H = np.zeros(shape, dtype=np.csingle)                   # :257
    valid = ~np.isnan(pw)                                   # :260
    clip = np.zeros_like(valid)
    path_slot = np.full(valid.shape, -1, dtype=np.int32)
    for i in range(n_ue):                                   # :264
        m = valid[i]
        n_i = int(m.sum())
        path_slot[i, m] = np.arange(n_i)
        if n_i == 0:
            continue                                        # :270-271
        arp = a_rx[i][:, None, :][..., m] * a_tx[i][None, :, :][..., m]     # dataset.py:417 + channel.py:274
        if dop is None or tt is None:
            dphase = None
        else:
            dphase = np.exp(1j * 2 * np.pi * np.outer(dop[i, m].astype(np.float64), tt))   # [P_i, T] (a11 definition)
        if freq_domain:
            g, over = ofdm_path_gains(pw[i, m], delay[i, m], phase[i, m], int(subcarriers), sel_sc, ts, int(rx_filter))
            clip[i, m] = over
            if dphase is None:
                H[i] = np.nansum(arp[..., None, :] * g.T[None, None, :, :], axis=-1)        # channel.py:283-284
            else:
                for it in range(len(tt)):
                    H[i, ..., it] = np.nansum(arp[..., None, :] * (g * dphase[:, it:it + 1]).T[None, None, :, :], axis=-1)
        else:
            pg = np.sqrt(pw[i, m]) * np.exp(1j * np.deg2rad(phase[i, m]))                   # channel.py:286
            if dphase is None:
                H[i, ..., :n_i] = arp * pg[None, None, :]                                   # :287
            else:
                H[i, :, :, :n_i, :] = (arp * pg[None, None, :])[..., None] * dphase[None, None, :, :]
    return dict(H=H, fov_mask=fov_mask, valid=valid, clip=clip, path_slot=path_slot)


# --------------------------------------------------------------------------
# row a11: v3's constant per-path Doppler phase (the only Doppler definition in the reference tree)
# --------------------------------------------------------------------------
LIGHTSPEED = 299792458   # deepmimo_v3/consts.py:112


def v3_doppler_shift(toa, vel, acc, carrier_hz: float) -> np.ndarray:
    """deepmimo_v3/generator/python/construct_deepmimo.py:267-280 (no-LPF branch): path gains are multiplied by
    exp(-j 2 pi f_c (v tau / c + a tau^2 / (2 c))), tau = ToA.  Returned as the float32 per-path Doppler shift f_D [Hz] for which
    `compute_channels(..., doppler_hz=f_D, times=[1.0])` applies exactly that phase (exp(+j 2 pi f_D t) at t = 1 s).
    Pinned by tests/golden/doppler_v3.npz (generated by running v3 itself)."""
    tau = np.asarray(toa, dtype=np.float64)
    v, a = np.asarray(vel, dtype=np.float64), np.asarray(acc, dtype=np.float64)
    return (-carrier_hz * (v * tau / LIGHTSPEED + a * tau ** 2 / (2 * LIGHTSPEED))).astype(np.float32)


# --------------------------------------------------------------------------
# row f3: beamforming codebooks and beam amplitude maps (adjacent consumer of H)
# --------------------------------------------------------------------------
def steering_vec(array, phi: float = 0, theta: float = 0, spacing: float = 0.5) -> np.ndarray:
    """Normalised steering vector [M, 1] complex128.  deepmimo/generator/geometry.py:322-339.

    The reference passes (phi*pi/180, theta*pi/180 + pi/2) to `_array_response(ant_ind, theta, phi, kd)`
    (geometry.py:338, :19), i.e. the azimuth lands in the `theta` slot and the shifted elevation in the `phi`
    slot; the phases follow geometry.py:99-101 with those arguments.  Restated as is.
    """
    idx = element_grid(array)                                   # :337
    th, ph, kd = phi * np.pi / 180, theta * np.pi / 180 + np.pi / 2, 2 * np.pi * spacing
    gamma = np.vstack([1j * kd * np.sin(th) * np.cos(ph), 1j * kd * np.sin(th) * np.sin(ph), 1j * kd * np.cos(th)]).T   # :99-102
    resp = np.exp(idx @ gamma.T)                                # :35
    return resp / np.linalg.norm(resp)                          # :339


def beam_amplitude(H: np.ndarray, F: np.ndarray) -> np.ndarray:
    """Mean amplitude per (user, beam) of the beamformed channel, docs/manual.ipynb cell 105:
    `np.abs(F1 @ channel).mean(axis=1).mean(axis=-1)` with F1 [n_beams, M_t], channel [n, M_r, M_t, K]."""
    return np.abs(F @ H).mean(axis=1).mean(axis=-1)
