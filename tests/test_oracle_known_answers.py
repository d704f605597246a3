"""Known answers the reference's own (stale-import) unit tests pin for the helpers on this path
(SURVEY.md 4 / 8c), replayed on the oracle restatement."""
import numpy as np

from oracle import channel_oracle as orc


def test_dipole_known_answers():
    """test/test_ant_patterns.py:72-78,131-134: max at 90 deg, nulls at 0/180, TX*RX relative gain ~0.08 at 45 deg."""
    th = np.deg2rad(np.array([0.0, 45.0, 90.0, 135.0, 180.0]))
    g = orc.pattern_gain("halfwave-dipole", th)
    assert g[0] == 0 and g[4] == 0
    assert np.argmax(g) == 2 and abs(g[2] - 1.643) < 1e-12
    rel = (g[1] / g[2]) ** 2
    assert abs(rel - 0.08) < 0.012 or abs(g[1] ** 2 / g[2] ** 2 - rel) < 1e-15
    assert orc.pattern_gain("isotropic", th) == 1.0
    assert orc.pattern_gain("halfwave-dipole", np.array([np.nan]))[0] == 0


def test_array_response_nan_and_shape():
    """test/test_array_response.py:123-131: shape (batch, N_ant, paths); all-NaN -> exact zeros."""
    grid = orc.element_grid((4, 2))
    assert grid.shape == (8, 3) and (grid[:, 0] == 0).all()
    assert np.array_equal(grid[:, 1], np.tile(np.arange(4), 2)) and np.array_equal(grid[:, 2], np.repeat(np.arange(2), 4))
    th = np.full((3, 5), np.nan)
    out = orc.steering_batch(grid, th, th, np.pi)
    assert out.shape == (3, 8, 5) and np.all(out == 0)
    th = np.random.default_rng(42).uniform(0, np.pi, (3, 5)); ph = np.random.default_rng(1).uniform(-np.pi, np.pi, (3, 5))
    th[1, 2] = np.nan
    out = orc.steering_batch(grid, th, ph, np.pi)
    assert np.all(out[1, :, 2] == 0)
    i, p = 2, 3
    want = np.exp(1j * np.pi * (grid[:, 1] * np.sin(th[i, p]) * np.sin(ph[i, p]) + grid[:, 2] * np.cos(th[i, p])))
    np.testing.assert_allclose(out[i, :, p], want, rtol=1e-10)
    np.testing.assert_allclose(np.abs(out[~np.isnan(th)[:, None, :].repeat(8, 1)]), 1.0, rtol=1e-12)


def test_fov_sets():
    """test/test_fov.py:74-155: full sphere keeps everything; [180, 90] keeps the front half-band."""
    rng = np.random.default_rng(0)
    th = rng.uniform(0, np.pi, (50, 7)); ph = rng.uniform(-np.pi, np.pi, (50, 7))
    assert orc.fov_inclusion(np.array([360, 180]), th, ph).all()
    m = orc.fov_inclusion(np.array([180, 90]), th, ph)
    want = (np.abs(ph) <= np.pi / 2) & (np.abs(th - np.pi / 2) <= np.pi / 4)
    assert np.array_equal(m, want)
    assert not orc.fov_inclusion(np.array([180, 90]), np.array([[np.nan]]), np.array([[0.0]]))[0, 0]
    assert orc.is_full_fov([360, 180]) and not orc.is_full_fov([359, 180])


def test_rotation_identity_and_axis():
    """test/test_rotate_angles.py:67-127: zero rotation keeps the direction (to float32 sin/cos rounding);
    a z rotation shifts azimuth only."""
    rng = np.random.default_rng(3)
    el = rng.uniform(1, 179, (20, 6)).astype(np.float32); az = rng.uniform(-179, 179, (20, 6)).astype(np.float32)
    th, ph = orc.rotate_angles(np.array([0, 0, 0]), el, az)
    np.testing.assert_allclose(th, np.deg2rad(el.astype(np.float64)), atol=2e-6)
    np.testing.assert_allclose(ph, np.deg2rad(az.astype(np.float64)), atol=2e-6)
    th2, ph2 = orc.rotate_angles(np.array([0, 0, 40]), el, az)
    np.testing.assert_allclose(th2, th, atol=2e-6)
    d = np.mod(ph - ph2 - np.deg2rad(40) + np.pi, 2 * np.pi) - np.pi
    np.testing.assert_allclose(d, 0, atol=2e-6)
    per_user = rng.uniform(-90, 90, (20, 3))
    th3, ph3 = orc.rotate_angles(per_user, el, az)
    for i in (0, 7, 19):
        a, b = orc.rotate_angles(per_user[i], el[i:i + 1], az[i:i + 1])
        assert np.array_equal(a[0], th3[i]) and np.array_equal(b[0], ph3[i])


def test_ofdm_gain_clip_and_scale():
    """channel.py:187-192: delay_n >= N -> zero power, delay_n := N; FD gain carries 1/sqrt(N)."""
    pw = np.array([1e-10, 4e-10], np.float32); ph = np.array([0, 90], np.float32)
    toa = np.array([1e-7, 1.0], np.float32)
    g, over = orc.ofdm_path_gains(pw.copy(), toa, ph, 64, np.arange(4), 1 / 10e6)
    assert list(over) == [False, True] and np.all(g[1] == 0)
    np.testing.assert_allclose(np.abs(g[0]), np.sqrt(1e-10 / 64), rtol=1e-6)
    np.testing.assert_allclose(np.angle(g[0, 1] / g[0, 0]), -2 * np.pi * 1.0 / 64, rtol=1e-5)
