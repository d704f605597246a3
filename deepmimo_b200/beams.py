"""Beamforming consumer of the channel path.

`steering_vec` mirrors the reference's public codebook helper (deepmimo/generator/geometry.py:322-339,
exported at deepmimo/__init__.py:36-38); it is host-side parameter preparation ([M, 1] complex128, like the
reference).  `beam_amplitude` is the GPU part: the amplitude map of docs/manual.ipynb cell 105,

    np.abs(F1 @ dataset.channel).mean(axis=1).mean(axis=-1)            # [n_ue, n_beams]

computed by `dmk_beam_amplitude_fd` (include/dmk.h, csrc/dmk_bf.cuh) straight from the path matrices: the
codebook is folded into the TX steering of every path and the product is reduced in registers, so H is
never written.  No CPU fallback: without libdmk.so / a CUDA device this raises.
"""
from __future__ import annotations

import numpy as np

from .channels import _torch, make_plan


def steering_vec(array, phi: float = 0, theta: float = 0, spacing: float = 0.5) -> np.ndarray:
    """Normalised steering vector of a panel, [M, 1] complex128 (geometry.py:322-339).

    Bug-compatible with the reference: it calls `_array_response(idxs, phi*pi/180, theta*pi/180 + pi/2, kd)`
    whose signature is `(ant_ind, theta, phi, kd)` (geometry.py:19, :338), so the azimuth is used as the polar
    angle and the shifted elevation as the azimuth of geometry.py:99-101.  Element order: y fastest, x == 0
    (`_ant_indices`, :105-120).
    """
    m_h, m_v = int(array[0]), int(array[1])
    n = np.arange(m_h * m_v)
    idx = np.stack([np.zeros_like(n), n % m_h, n // m_h], axis=1)
    th, ph, kd = phi * np.pi / 180, theta * np.pi / 180 + np.pi / 2, 2 * np.pi * spacing
    gamma = np.vstack([1j * kd * np.sin(th) * np.cos(ph), 1j * kd * np.sin(th) * np.sin(ph), 1j * kd * np.cos(th)]).T
    resp = np.exp(idx @ gamma.T)
    return resp / np.linalg.norm(resp)


def beam_amplitude(dataset, beams, params=None, *, out: str = "numpy", device=None, return_info: bool = False,
                   seed_numpy_rng: bool = True, warn: bool = True):
    """Mean amplitude of the beamformed frequency-domain channel per (user, beam): float32 `[n_ue, n_beams]`,
    equal to `np.abs(beams @ H).mean(axis=1).mean(axis=-1)` for `H = dataset.compute_channels(params)`.

    beams: complex `[n_beams, M_t]` (rows e.g. `steering_vec(bs_shape, phi=az).squeeze()`).  Parameters, FoV and
    rotations are taken from `params` / the dataset exactly as `compute_channels` does.  `out='torch'` returns the
    CUDA tensor.
    """
    torch = _torch()
    plan, params = make_plan(dataset, params, device=device, seed_numpy_rng=seed_numpy_rng, warn=warn)
    s = plan.spec
    if not s.freq_domain:
        raise ValueError("beam_amplitude works on the frequency-domain channel (freq_domain=1)")
    F = np.ascontiguousarray(np.asarray(beams.detach().cpu().numpy() if isinstance(beams, torch.Tensor) else beams),
                             dtype=np.complex64)
    if F.ndim != 2 or F.shape[1] != s.m_tx:
        raise ValueError(f"beams must have shape [n_beams, {s.m_tx}] (M_t = prod(bs_antenna.shape)), got {F.shape}")
    Fd = torch.from_numpy(F).to(plan.device)
    n, nb = plan.n_users, int(F.shape[0])
    res = torch.empty((n, nb), dtype=torch.float32, device=plan.device)
    masks = plan.alloc_masks()
    if n:
        plan.run_beams(Fd, res, masks)
    info = plan.info_from_masks(masks) if return_info else None
    val = res if out == "torch" else res.cpu().numpy()
    return (val, info) if return_info else val
