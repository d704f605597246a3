"""Edge cases of the tensor-core FD kernels (fd_ws_kernel / fd_tc_kernel, forced with DMK_FD_KERNEL=tc / tc1) against the oracle:
tile raggedness in M and K, 32 path columns, FP16 operand scaling under a 150 dB power spread, sub-tiled small arrays,
single-path users, K split across CTAs (few users)."""
import numpy as np
import pytest

from util import assert_channels_close, make_dataset

pytestmark = pytest.mark.gpu


def _case(n, bs, ue, n_sc, k_sel, seed, n_cols=25, bs_rot=(10, 20, 30), power_lo=-160.0, power_hi=-60.0, zero_frac=0.1):
    from deepmimo_b200.synth import make_paths
    d = make_paths(n, seed, n_sc=n_sc, bandwidth=50e6, n_cols=n_cols, zero_frac=zero_frac)
    if (power_lo, power_hi) != (-160.0, -60.0):
        rng = np.random.default_rng(seed + 1)
        pw = -np.sort(-rng.uniform(power_lo, power_hi, d["power"].shape), axis=1).astype(np.float32)
        pw[np.isnan(d["power"])] = np.nan
        d["power"] = pw
    p = {"bs_antenna": {"shape": np.array(bs), "spacing": 0.5, "rotation": np.array(bs_rot), "radiation_pattern": "isotropic"},
         "ue_antenna": {"shape": np.array(ue), "spacing": 0.5, "rotation": np.array([0, 0, 0]), "radiation_pattern": "isotropic"},
         "enable_doppler": 0, "enable_dual_polar": 0, "num_paths": n_cols, "freq_domain": 1,
         "ofdm": {"subcarriers": n_sc, "selected_subcarriers": np.asarray(k_sel), "bandwidth": 50e6, "rx_filter": 0}}
    return d, p


def _run(d, p, monkeypatch, expect=("fd_ws_kernel", "fd_tc_kernel")):
    """Both tensor-core kernels: DMK_FD_KERNEL=tc is the warp-specialised persistent kernel (fd_ws_kernel; it hands shapes whose
    double-buffered tables do not fit to fd_tc_kernel), tc1 the one-CTA-per-user kernel.  Returns the worst error and the last info."""
    import deepmimo_b200 as dmb
    from oracle import channel_oracle as orc
    o = orc.compute_channels(d, bs_shape=p["bs_antenna"]["shape"], ue_shape=p["ue_antenna"]["shape"],
                             bs_rotation=p["bs_antenna"]["rotation"], num_paths=p["num_paths"],
                             subcarriers=p["ofdm"]["subcarriers"], selected_subcarriers=p["ofdm"]["selected_subcarriers"],
                             bandwidth=p["ofdm"]["bandwidth"])
    worst, info = 0.0, None
    for variant in ("tc", "tc1"):
        monkeypatch.setenv("DMK_FD_KERNEL", variant)
        H, info = make_dataset(dmb, d).compute_channels(dmb.ChannelGenParameters(p), return_info=True, warn=False)
        assert info.kernel.startswith(tuple(expect) if variant == "tc" else (expect[-1],)), info.kernel
        worst = max(worst, assert_channels_close(H, o["H"], what=info.kernel))
        assert np.array_equal(info.valid, o["valid"]) and np.array_equal(info.clip, o["clip"])
    return worst, info


@pytest.mark.parametrize("bs,ue,k", [((8, 8), (1, 1), 64),          # one segment, one 64-row tile
                                      ((12, 11), (1, 1), 128),       # M = 132: ragged second 128-row tile
                                      ((5, 3), (3, 1), 192),         # M = 45 -> 64-row tile, 3 segments (odd count with 2 sub-tiles)
                                      ((3, 3), (1, 1), 320),         # M = 9 -> 16-row tile
                                      ((16, 16), (2, 2), 64)])       # M = 1024, 8 row tiles
def test_tc_tile_raggedness(bs, ue, k, monkeypatch):
    d, p = _case(40, bs, ue, 1024, np.arange(k), 31)
    err, info = _run(d, p, monkeypatch)
    print(info.kernel, f"{err:.2e}")


def test_tc_32_path_columns_and_offset_stride_selection(monkeypatch):
    d, p = _case(64, (8, 8), (2, 1), 2048, 5 + 3 * np.arange(128), 32, n_cols=32)      # affine selection start 5, step 3
    err, _ = _run(d, p, monkeypatch)
    assert err < 2e-6


def test_tc_k4096_and_fallback_beyond(monkeypatch):
    d, p = _case(6, (8, 8), (1, 1), 4096, np.arange(4096), 33)
    _run(d, p, monkeypatch)
    d, p = _case(4, (8, 8), (1, 1), 8192, np.arange(8192), 34)
    _run(d, p, monkeypatch, expect=("fd_",))                 # K > 4096: the tensor-core kernel is not eligible, another FD kernel runs


def test_tc_fp16_scaling_under_150_db_spread(monkeypatch):
    """Path powers from -200 to -50 dBW inside one user: the per-user scale keeps the strong paths exact and the weak
    ones cost nothing against the per-user Frobenius criterion."""
    d, p = _case(128, (16, 8), (1, 1), 512, np.arange(512), 35, power_lo=-200.0, power_hi=-50.0, zero_frac=0.0)
    err, _ = _run(d, p, monkeypatch)
    assert err < 2e-6
    d, p = _case(128, (16, 8), (1, 1), 512, np.arange(512), 36, power_lo=-300.0, power_hi=-250.0, zero_frac=0.0)   # tiny everywhere
    err, _ = _run(d, p, monkeypatch)
    assert err < 2e-6


def test_tc_single_path_users_and_few_users_split_over_ctas(monkeypatch):
    d, p = _case(3, (16, 8), (1, 1), 1024, np.arange(1024), 37, n_cols=1, zero_frac=0.0)       # ksplit > 1, np = 1
    err, info = _run(d, p, monkeypatch)
    assert "ksplit=" in info.kernel and int(info.kernel.split("ksplit=")[1].split()[0]) > 1


def test_ws_many_users_per_cta_chunked_independent_launches_and_ticket_reuse():
    """Persistent kernel: far more users than resident CTAs (several users per CTA, zero-path users interleaved), the
    user range streamed through iter_channels' ring (chunks after the first are launched with DMK_FLAG_INDEPENDENT_LAUNCH
    and may overlap the previous chunk's tail), and the whole thing twice (ticket counters must return to zero)."""
    import torch
    import deepmimo_b200 as dmb
    from oracle import channel_oracle as orc
    d, p = _case(2500, (8, 8), (1, 1), 1024, np.arange(512), 41, zero_frac=0.2)      # 256 KB per user: above the fd_mma_kernel range
    o = orc.compute_channels(d, bs_shape=p["bs_antenna"]["shape"], ue_shape=p["ue_antenna"]["shape"],
                             bs_rotation=p["bs_antenna"]["rotation"], num_paths=p["num_paths"],
                             subcarriers=p["ofdm"]["subcarriers"], selected_subcarriers=p["ofdm"]["selected_subcarriers"],
                             bandwidth=p["ofdm"]["bandwidth"])
    plan, _ = dmb.make_plan(make_dataset(dmb, d), dmb.ChannelGenParameters(p), warn=False)
    for rep in range(2):
        H = np.empty(plan.out_shape(), dtype=np.complex64)
        chunks = list(dmb.iter_channels(plan, chunk_users=700, n_buffers=4))     # 4 chunks back to back, one buffer each
        torch.cuda.synchronize()
        for start, stop, buf in chunks:
            H[start:stop] = buf.cpu().numpy()
        from deepmimo_b200 import _lib
        assert _lib.last_kernel().startswith("fd_ws_kernel"), _lib.last_kernel()
        err = assert_channels_close(H, o["H"], what=f"ws chunked rep {rep}")
        assert err < 2e-6
    torch.cuda.synchronize()


@pytest.mark.parametrize("helpers", ["1", "4"])
def test_ws_helper_counts_agree(helpers, monkeypatch):
    """The persistent kernel runs with one helper warp (two CTAs per SM) or four (one CTA per SM, eight user buffers);
    DMK_WS_HELPERS pins the choice.  Both must meet the bar on a shape that streams many users through every CTA."""
    monkeypatch.setenv("DMK_WS_HELPERS", helpers)
    d, p = _case(1500, (8, 8), (1, 1), 1024, np.arange(192), 43, zero_frac=0.15)
    import deepmimo_b200 as dmb
    from deepmimo_b200 import _lib
    from oracle import channel_oracle as orc
    monkeypatch.setenv("DMK_FD_KERNEL", "tc")
    H = make_dataset(dmb, d).compute_channels(dmb.ChannelGenParameters(p), warn=False)
    assert f"{helpers} helper" in _lib.last_kernel(), _lib.last_kernel()
    o = orc.compute_channels(d, bs_shape=p["bs_antenna"]["shape"], ue_shape=p["ue_antenna"]["shape"],
                             bs_rotation=p["bs_antenna"]["rotation"], num_paths=p["num_paths"],
                             subcarriers=p["ofdm"]["subcarriers"], selected_subcarriers=p["ofdm"]["selected_subcarriers"],
                             bandwidth=p["ofdm"]["bandwidth"])
    assert assert_channels_close(H, o["H"], what=_lib.last_kernel()) < 2e-6
