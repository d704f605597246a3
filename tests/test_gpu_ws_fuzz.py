"""Randomised cross-check of the persistent tensor-core kernel (fd_ws_kernel, both helper configurations) against the packed-FP32
kernel on the same inputs: random panel shapes, subcarrier counts/offsets/strides, user counts from fewer-than-CTAs to thousands,
zero-path fractions from 0 to 1, path-column counts.  Both results come from the GPU; the FP32 kernel is itself pinned to the oracle
by the other tests.  The point is the synchronisation protocol (tickets, mbarrier parities, sentinel handling, K split across CTAs)."""
import numpy as np
import pytest

from util import per_user_rel_fro

pytestmark = pytest.mark.gpu


def _one(rng, monkeypatch, helpers, ragged=False):
    import torch
    import deepmimo_b200 as dmb
    from deepmimo_b200 import _lib
    from deepmimo_b200.synth import make_paths
    while True:
        bs = (int(rng.integers(1, 33)), int(rng.integers(1, 17)))
        ue = (int(rng.integers(1, 3)), int(rng.integers(1, 3)))
        m = bs[0] * bs[1] * ue[0] * ue[1]
        if 16 <= m <= 1024 and bs[0] <= 32:
            break
    nseg = int(rng.integers(1, 9))
    step = int(rng.choice([1, 1, 2, 3]))
    start = int(rng.integers(0, 5))
    k = 64 * nseg
    if ragged:                                          # K not a multiple of 64: cut-off last chunk of every antenna row (padding <= K / 2)
        k = int(rng.choice([88, 100, 120, 180, 257, 300, 330, 383, 450, 600, 624, 1000]))
    n_sc = int(2 ** np.ceil(np.log2(start + step * k + 1)))
    n = int(rng.choice([1, 3, 50, 290, 300, 700, 2500]))
    n = max(1, min(n, (1 << 28) // (8 * m * k)))
    n_cols = int(rng.choice([1, 7, 25, 32]))
    zero_frac = float(rng.choice([0.0, 0.1, 0.5, 1.0]))
    d = make_paths(n, int(rng.integers(1, 10 ** 6)), n_sc=n_sc, bandwidth=50e6, n_cols=n_cols, zero_frac=zero_frac)
    p = dmb.ChannelGenParameters()
    p.bs_antenna.shape = np.array(bs); p.ue_antenna.shape = np.array(ue)
    p.bs_antenna.rotation = np.array([5, 10, 15]); p.num_paths = n_cols
    p.ofdm.subcarriers = n_sc; p.ofdm.selected_subcarriers = start + step * np.arange(k); p.ofdm.bandwidth = 50e6
    plan, _ = dmb.make_plan(dmb.Dataset(d), p, warn=False)
    monkeypatch.setenv("DMK_FD_KERNEL", "ffma")
    ref = plan.run(plan.alloc_out()).cpu().numpy()
    monkeypatch.setenv("DMK_FD_KERNEL", "tc")
    monkeypatch.setenv("DMK_WS_HELPERS", helpers)
    out = plan.alloc_out()
    for _ in range(2):                                   # twice: ticket counters must be back at zero
        out.fill_(complex(float("nan"), 0.0))
        got = plan.run(out).cpu().numpy()
        torch.cuda.synchronize()
        kern = _lib.last_kernel()
        desc = f"bs{bs} ue{ue} K={k} start={start} step={step} n={n} cols={n_cols} zero={zero_frac} -> {kern}"
        assert not np.isnan(got.view(np.float32)).any(), "unwritten output: " + desc
        err = per_user_rel_fro(got, ref)
        assert err.size == 0 or err.max() < 2e-6, f"{err.max():.2e} " + desc
    return kern


@pytest.mark.parametrize("helpers", ["1", "4"])
def test_ws_kernel_random_shapes_match_fp32_kernel(helpers, monkeypatch):
    rng = np.random.default_rng(2024 + int(helpers))
    seen = set()
    for _ in range(24):
        seen.add(_one(rng, monkeypatch, helpers).split("<")[0])
    assert "fd_ws_kernel" in seen


@pytest.mark.parametrize("helpers", ["1", "4"])
def test_ws_kernel_subcarrier_counts_that_are_not_multiples_of_64(helpers, monkeypatch):
    """12 x n resource blocks (300, 600, 624 ...) and odd counts: the persistent kernel cuts the last chunk of every row off
    (ws_store_chunks_rag); zero-path users go through the same store route."""
    rng = np.random.default_rng(4048 + int(helpers))
    seen = set()
    for _ in range(16):
        seen.add(_one(rng, monkeypatch, helpers, ragged=True).split("<")[0])
    assert "fd_ws_kernel" in seen
