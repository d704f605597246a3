"""Golden vectors for row f3 (beamforming consumer), from the LIVE reference: `dm.steering_vec` codebooks
(deepmimo/generator/geometry.py:322-339) and the beam amplitude map of docs/manual.ipynb cell 105,
`np.abs(F1 @ dataset.channel).mean(axis=1).mean(axis=-1)`, on two of the cases of cases.py.

    python tests/golden/make_golden_beams.py        # build container only (needs /root/reference)
"""
import os
import sys
from unittest.mock import MagicMock

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
for m in ("matplotlib", "matplotlib.pyplot", "matplotlib.figure", "matplotlib.axes", "matplotlib.colorbar",
          "matplotlib.colors", "mpl_toolkits", "mpl_toolkits.mplot3d", "mpl_toolkits.mplot3d.art3d"):
    sys.modules.setdefault(m, MagicMock())
sys.path.insert(0, "/root/reference")
os.environ.setdefault("TQDM_DISABLE", "1")

import deepmimo as dm  # noqa: E402

from cases import case_list, case_data, params_dict  # noqa: E402
from make_golden import run_reference  # noqa: E402
from oracle import channel_oracle as orc  # noqa: E402

from make_golden_beams_cases import BEAM_CASES, codebook  # noqa: E402


def main():
    out = {}
    cases = {c["name"]: c for c in case_list()}
    for name, b in BEAM_CASES.items():
        c = cases[name]
        F = codebook(dm.steering_vec, c["bs_shape"], c["bs_sp"], b["phis"], b["thetas"])          # [n_beams, M_t] complex128
        Fo = codebook(orc.steering_vec, c["bs_shape"], c["bs_sp"], b["phis"], b["thetas"])
        assert np.array_equal(F.view(np.float64), Fo.view(np.float64)), name                      # oracle codebook: bit-identical
        H, _, _ = run_reference(c, case_data(c))
        amp = np.abs(F @ H).mean(axis=1).mean(axis=-1)                                             # manual.ipynb cell 105
        assert np.array_equal(amp, orc.beam_amplitude(H, F))
        out[f"{name}__F"] = F
        out[f"{name}__amp"] = amp
        print(f"{name}: F{F.shape} amp{amp.shape} max {amp.max():.3e}")
    np.savez_compressed(os.path.join(HERE, "beams.npz"), **out)


if __name__ == "__main__":
    main()
