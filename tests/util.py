"""Shared helpers for the parity tests."""
import numpy as np

TOL_REL_FRO = 1e-5      # north_star: per-user relative Frobenius error of the channel coefficients


def per_user_rel_fro(got: np.ndarray, ref: np.ndarray) -> np.ndarray:
    """||got_u - ref_u||_F / ||ref_u||_F per user (absolute norm where the reference user is all-zero)."""
    assert got.shape == ref.shape, (got.shape, ref.shape)
    n = got.shape[0]
    g = got.reshape(n, -1).astype(np.complex128)
    r = ref.reshape(n, -1).astype(np.complex128)
    num = np.linalg.norm(g - r, axis=1)
    den = np.linalg.norm(r, axis=1)
    return np.where(den > 0, num / np.where(den > 0, den, 1.0), num)


def assert_channels_close(got, ref, tol=TOL_REL_FRO, what=""):
    assert got.dtype == np.complex64 and got.shape == ref.shape, (what, got.dtype, got.shape, ref.shape)
    assert not np.isnan(got.view(np.float32)).any(), f"{what}: NaN in output"
    err = per_user_rel_fro(got, ref)
    zero_ref = np.abs(ref.reshape(ref.shape[0], -1)).sum(1) == 0
    assert np.all(np.abs(got.reshape(got.shape[0], -1))[zero_ref] == 0), f"{what}: users without paths must be exactly zero"
    worst = int(np.argmax(err)) if err.size else -1
    assert err.size == 0 or err.max() <= tol, f"{what}: max per-user rel. Frobenius {err.max():.3e} at user {worst} > {tol}"
    return float(err.max()) if err.size else 0.0


def make_dataset(dmb, scen_or_data, bs_fov=None, ue_fov=None):
    data = scen_or_data if isinstance(scen_or_data, dict) else scen_or_data.data
    ds = dmb.Dataset({k: v for k, v in data.items()})
    if bs_fov is not None or ue_fov is not None:
        kw = {}
        if bs_fov is not None:
            kw["bs_fov"] = bs_fov
        if ue_fov is not None:
            kw["ue_fov"] = ue_fov
        ds.apply_fov(**kw)
    return ds


def oracle_kwargs_from_params(p: dict, bs_fov=None, ue_fov=None) -> dict:
    return dict(bs_shape=p["bs_antenna"]["shape"], ue_shape=p["ue_antenna"]["shape"],
                bs_spacing=p["bs_antenna"]["spacing"], ue_spacing=p["ue_antenna"]["spacing"],
                bs_rotation=p["bs_antenna"]["rotation"], ue_rotation=p["ue_antenna"]["rotation"],
                bs_pattern=p["bs_antenna"]["radiation_pattern"], ue_pattern=p["ue_antenna"]["radiation_pattern"],
                bs_fov=bs_fov, ue_fov=ue_fov, num_paths=p["num_paths"], freq_domain=bool(p["freq_domain"]),
                subcarriers=p["ofdm"]["subcarriers"], selected_subcarriers=p["ofdm"]["selected_subcarriers"],
                bandwidth=p["ofdm"]["bandwidth"], rx_filter=int(p["ofdm"].get("rx_filter", 0)))


def scenario_subset(s, idx):
    """The users `idx` of a synth.Scenario as (data, oracle kwargs, doppler): every per-user array follows the selection."""
    idx = np.asarray(idx)
    n = s.n_ue
    data = {k: (v[idx] if getattr(v, "shape", (0,))[0] == n and k != "tx_pos" else v) for k, v in s.data.items()}
    kw = oracle_kwargs_from_params(s.params, s.bs_fov, s.ue_fov)
    rot = np.asarray(kw["ue_rotation"])
    if rot.ndim == 2 and rot.shape[0] == n:
        kw["ue_rotation"] = rot[idx]
    dop = None if s.doppler_hz is None else s.doppler_hz[idx]
    return data, kw, dop


def _oracle_job(args):
    from oracle import channel_oracle as orc
    data, kw, dop, times = args
    o = orc.compute_channels(data, **kw, doppler_hz=dop, times=times)
    return o["H"], o["valid"], o["clip"], o["fov_mask"]


def oracle_on_users(s, idx, procs=None):
    """Oracle channels + masks of the users `idx` (any order) of scenario `s`, spread over `procs` processes (the NumPy path is
    single-threaded like the reference's per-user loop; default: all host cores, at most one job per 8 users)."""
    import multiprocessing as mp
    import os
    idx = np.asarray(idx)
    procs = min(os.cpu_count() or 1, 32, max(1, len(idx) // 8)) if procs is None else procs
    parts = [p for p in np.array_split(np.arange(len(idx)), procs) if len(p)]
    jobs = []
    for p in parts:
        data, kw, dop = scenario_subset(s, idx[p])
        jobs.append((data, kw, dop, s.times))
    if len(jobs) == 1:
        res = [_oracle_job(jobs[0])]
    else:
        with mp.get_context("fork").Pool(len(jobs)) as pool:
            res = pool.map(_oracle_job, jobs)
    H = np.concatenate([r[0] for r in res])
    valid = np.concatenate([r[1] for r in res])
    clip = np.concatenate([r[2] for r in res])
    fov = None if res[0][3] is None else np.concatenate([r[3] for r in res])
    return dict(H=H, valid=valid, clip=clip, fov_mask=fov)
