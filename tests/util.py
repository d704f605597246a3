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
