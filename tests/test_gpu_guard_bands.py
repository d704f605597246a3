"""Out-of-bounds write detector (compute-sanitizer is not available on the GPU pool): every kernel variant writes
into an output that sits between two canary-filled guard bands; the guards and the bytes between users must come back
untouched, and every element of the output must have been written (no canary left inside)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GUARD = 1 << 16          # complex64 elements on each side (512 KiB)
CANARY = np.complex64(complex(-7.0e33, 3.0e-33))


def _plan(bs, ue, n_sc, sel, n, fd=1, times=None, n_cols=25, seed=5, rx_filter=0):
    import deepmimo_b200 as dmb
    from deepmimo_b200.synth import doppler_from_velocity, make_paths
    d = make_paths(n, seed, n_sc=n_sc, bandwidth=50e6, n_cols=n_cols)
    p = dmb.ChannelGenParameters()
    p.bs_antenna.shape = np.array(bs); p.ue_antenna.shape = np.array(ue)
    p.bs_antenna.rotation = np.array([5, 10, 15])
    p.ofdm.subcarriers = n_sc; p.ofdm.selected_subcarriers = np.asarray(sel); p.ofdm.bandwidth = 50e6
    p.freq_domain = fd; p.num_paths = n_cols; p.ofdm.rx_filter = rx_filter
    dop = doppler_from_velocity(d, 1, 3.5e9) if times is not None else None
    plan, _ = dmb.make_plan(dmb.Dataset(d), p, times=times, doppler=dop, warn=False)
    return plan


def _run_guarded(plan):
    import torch
    shape = plan.out_shape()
    n = int(np.prod(shape))
    big = torch.full((n + 2 * GUARD,), complex(CANARY), dtype=torch.complex64, device=plan.device)
    out = big[GUARD:GUARD + n].view(shape)
    masks = plan.alloc_masks()
    mbig = {k: torch.full((v.numel() + 2 * 4096,), 77, dtype=v.dtype, device=v.device) for k, v in masks.items()}
    mview = {k: mbig[k][4096:4096 + masks[k].numel()].view(masks[k].shape) for k in masks}
    plan.run(out, 0, plan.n_users, mview)
    torch.cuda.synchronize()
    h = big.cpu().numpy()
    assert np.all(h[:GUARD].view(np.uint64) == np.array([CANARY]).view(np.uint64)[0]), "write before the output"
    assert np.all(h[GUARD + n:].view(np.uint64) == np.array([CANARY]).view(np.uint64)[0]), "write past the output"
    inner = h[GUARD:GUARD + n]
    assert not np.any(inner.view(np.uint64) == np.array([CANARY]).view(np.uint64)[0]), "an output element was never written"
    assert np.isfinite(inner.view(np.float32)).all()
    for k, v in mbig.items():
        hv = v.cpu().numpy()
        assert np.all(hv[:4096] == 77) and np.all(hv[4096 + masks[k].numel():] == 77), f"mask {k} written out of bounds"
    return inner.reshape(shape)


@pytest.mark.parametrize("variant", ["tc", "tc1", "ffma", "tile"])
@pytest.mark.parametrize("bs,ue,k", [((12, 11), (1, 1), 192), ((5, 3), (3, 1), 64), ((3, 3), (1, 1), 320), ((9, 8), (2, 1), 128)])
def test_fd_kernels_stay_in_bounds(variant, bs, ue, k, monkeypatch):
    monkeypatch.setenv("DMK_FD_KERNEL", variant)
    H = _run_guarded(_plan(bs, ue, 1024, np.arange(k), 37))
    assert H.shape == (37, ue[0] * ue[1], bs[0] * bs[1], k)


@pytest.mark.parametrize("sel", [np.arange(3) * 3, np.array([0, 1, 5, 9, 600]), np.arange(1), np.arange(70)])
def test_fd_odd_column_counts_stay_in_bounds(sel):
    _run_guarded(_plan((7, 3), (1, 2), 1024, sel, 29))


@pytest.mark.parametrize("variant", ["small", "small1"])
@pytest.mark.parametrize("bs,ue,sel", [((8, 1), (1, 1), np.arange(64)), ((3, 1), (1, 1), np.arange(7)), ((4, 2), (2, 1), np.arange(130)),
                                        ((1, 1), (1, 1), np.arange(1)), ((5, 1), (1, 3), 2 + 5 * np.arange(33)),
                                        ((4, 4), (1, 1), np.arange(300)), ((2, 1), (2, 1), 7 + 3 * np.arange(1000))])
def test_small_array_kernel_stays_in_bounds(bs, ue, sel, variant, monkeypatch):
    from deepmimo_b200 import _lib
    monkeypatch.setenv("DMK_FD_KERNEL", variant)
    _run_guarded(_plan(bs, ue, 4096, sel, 53))
    assert _lib.last_kernel().startswith("fd_small2_kernel" if variant == "small" else "fd_small_kernel<"), _lib.last_kernel()


@pytest.mark.parametrize("times", [None, np.arange(16) * 1e-3, np.arange(5) * 1e-3, np.arange(70) * 1e-4])
def test_td_kernel_stays_in_bounds(times):
    H = _run_guarded(_plan((7, 3), (1, 2), 512, np.arange(1), 41, fd=0, times=times, n_cols=23))
    assert H.shape[:4] == (41, 2, 21, 23)


def test_fd_time_axis_stays_in_bounds():
    _run_guarded(_plan((4, 2), (1, 1), 256, np.arange(48), 19, times=np.arange(3) * 1e-3))


@pytest.mark.parametrize("n_sc,sel", [(256, np.arange(256)), (600, np.arange(0, 600, 7)), (1024, np.arange(130))])
def test_fd_lowpass_filter_stays_in_bounds(n_sc, sel):
    H = _run_guarded(_plan((5, 3), (2, 1), n_sc, sel, 21, rx_filter=1))
    assert H.shape == (21, 2, 15, len(sel))
