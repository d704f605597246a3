"""Per-path by-products of the prologue kernel: the Dataset caches the reference fills lazily.

Rotated angles (deepmimo/generator/dataset.py:310-356), FoV mask and FoV-filtered angles (:461-512)
and linear power with antenna gain (:665-691), computed by `dmk_path_prologue` (include/dmk.h) with the
same rounding flow as the channel kernels, returned under the reference's cache keys
(deepmimo/consts.py:212-226).
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib
from . import channels as _ch
from .params import ChannelGenParameters


def path_byproducts(dataset, params=None, *, device=None, seed_numpy_rng: bool = True) -> dict:
    import torch
    if params is None:
        params = ChannelGenParameters()
    arrays = _ch._dataset_arrays(dataset)
    n_ue = int(arrays["power"].shape[0])
    params.validate(n_ue)
    if seed_numpy_rng:
        np.random.seed(1001)
    spec = _ch.parse_spec(params, n_ue, bs_fov=_ch._get(dataset, "bs_fov"), ue_fov=_ch._get(dataset, "ue_fov"),
                          seed_numpy_rng=False)
    plan = _ch.ChannelPlan(spec, arrays, device=device)
    n, p0 = plan.n_users, plan.n_cols
    ang = torch.empty((4, n, p0), dtype=torch.float64, device=plan.device)
    pw = torch.empty((n, p0), dtype=torch.float64, device=plan.device)
    fov = torch.empty((n, p0), dtype=torch.uint8, device=plan.device)
    t = plan.t
    with torch.cuda.device(plan.device):
        rc = plan.lib.dmk_path_prologue(ctypes.byref(plan.desc), t["power"].data_ptr(), t["aoa_az"].data_ptr(),
                                        t["aoa_el"].data_ptr(), t["aod_az"].data_ptr(), t["aod_el"].data_ptr(),
                                        None if plan.ue_rot is None else plan.ue_rot.data_ptr(), n, p0,
                                        ang.data_ptr(), pw.data_ptr(), fov.data_ptr(),
                                        torch.cuda.current_stream(plan.device).cuda_stream)
    _lib.check(rc)
    ang_h = ang.cpu().numpy()
    mask = fov.cpu().numpy().astype(bool)
    out = {"_aod_el_rot": ang_h[0], "_aod_az_rot": ang_h[1], "_aoa_el_rot": ang_h[2], "_aoa_az_rot": ang_h[3]}
    if spec.fov_any:
        out["_fov_mask"] = mask
        for k in ("_aod_el_rot", "_aod_az_rot", "_aoa_el_rot", "_aoa_az_rot"):
            out[k + "_fov"] = np.where(mask, out[k], np.nan)
    else:
        out["_fov_mask"] = None
        for k in ("_aod_el_rot", "_aod_az_rot", "_aoa_el_rot", "_aoa_az_rot"):
            out[k + "_fov"] = out[k]
    pg = pw.cpu().numpy()
    iso = spec.patterns == (0, 0)
    out["_power_linear_ant_gain"] = pg.astype(np.float32) if (iso and not plan.f64) else pg     # float32 stays float32 when isotropic
    return out


def user_byproducts(dataset, params=None, *, device=None, seed_numpy_rng: bool = True, want=("num_paths", "los", "pathloss")) -> dict:
    """Per-user by-products on the device (`dmk_user_byproducts`): `num_paths` after FoV filtering (dataset.py:613-619), `los`
    (:569-611, needs dataset['inter']), `pathloss` coherent and `pathloss_noncoherent` (:541-566), with the reference's dtypes
    (int64 / int64 / float32).  One warp per user runs the same rotation / FoV prologue as the channel kernels."""
    import torch
    if params is None:
        params = ChannelGenParameters()
    arrays = _ch._dataset_arrays(dataset)
    n_ue = int(arrays["power"].shape[0])
    params.validate(n_ue)
    if seed_numpy_rng:
        np.random.seed(1001)
    spec = _ch.parse_spec(params, n_ue, bs_fov=_ch._get(dataset, "bs_fov"), ue_fov=_ch._get(dataset, "ue_fov"),
                          seed_numpy_rng=False)
    plan = _ch.ChannelPlan(spec, arrays, device=device)
    n, p0 = plan.n_users, plan.n_cols
    dev = plan.device
    inter = None
    if "los" in want:
        inter = plan._to_dev(np.ascontiguousarray(dataset["inter"], dtype=np.float32), torch.float32, (n, p0), "inter")
    out_np = torch.empty(n, dtype=torch.int32, device=dev) if "num_paths" in want else None
    out_los = torch.empty(n, dtype=torch.int32, device=dev) if "los" in want else None
    out_plc = torch.empty(n, dtype=torch.float32, device=dev) if "pathloss" in want else None
    out_pln = torch.empty(n, dtype=torch.float32, device=dev) if "pathloss" in want else None
    ptr = lambda t: None if t is None else t.data_ptr()
    t = plan.t
    with torch.cuda.device(dev):
        rc = plan.lib.dmk_user_byproducts(ctypes.byref(plan.desc), t["power"].data_ptr(), t["phase"].data_ptr(),
                                          t["aoa_az"].data_ptr(), t["aoa_el"].data_ptr(), t["aod_az"].data_ptr(),
                                          t["aod_el"].data_ptr(), ptr(inter), ptr(plan.ue_rot), n, p0,
                                          ptr(out_np), ptr(out_los), ptr(out_plc), ptr(out_pln),
                                          torch.cuda.current_stream(dev).cuda_stream)
    _lib.check(rc)
    res = {}
    if out_np is not None:
        res["num_paths"] = out_np.cpu().numpy().astype(np.int64)
    if out_los is not None:
        res["los"] = out_los.cpu().numpy().astype(np.int64)
    if out_plc is not None:
        res["pathloss"] = out_plc.cpu().numpy()
        res["pathloss_noncoherent"] = out_pln.cpu().numpy()
    return res
