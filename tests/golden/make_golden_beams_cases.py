"""Beam codebooks of the row-f3 golden vectors (shared by make_golden_beams.py and tests/test_beams.py)."""
import numpy as np

BEAM_CASES = {"cfg2_shape": dict(phis=np.around(np.linspace(-60, 60, 16), 2), thetas=[0.0]),
              "mixed_pattern_bsfov": dict(phis=[-45.0, -10.0, 0.0, 30.0, 75.0], thetas=[-20.0, 15.0]),
              "cfg1_shape": dict(phis=np.around(np.linspace(-60, 60, 16), 2), thetas=[0.0])}


def codebook(steer, shape, spacing, phis, thetas):
    return np.array([steer(shape, phi=p, theta=t, spacing=spacing).squeeze() for t in thetas for p in phis])


