"""Sionna `(a, tau)` layout straight from the path matrices.

Mirror of the reference adapter `DeepMIMOSionnaAdapter` (deepmimo/integrations/sionna_adapter.py:21-200).  The reference walks a
v3-style dataset dict, copies each user's time-domain channel `dataset[bs]['user']['channel'][ue]` into
`a[i_rx, :, i_tx, :, :, 0]` and the path delays into `tau[i_rx, i_tx, :num_paths]`, one sample at a time in Python.
Here the time-domain kernel produces both arrays for every user of a base station in one launch
(`dmk_channels_td_tau`, include/dmk.h) and the samples are gathered on the device.

    adapter = DeepMIMOSionnaAdapter(datasets, params, bs_idx=..., ue_idx=...)   # datasets: Dataset | MacroDataset | list
    a, tau = adapter.arrays()            # [num_samples, num_rx, num_rx_ant, num_tx, num_tx_ant, num_paths, 1], [num_samples, num_rx, num_tx, num_paths]
    for a_i, tau_i in adapter():         # the reference's generator protocol, same order (UE samples outer, BS samples inner)
        ...

`bs_idx` / `ue_idx` follow the reference (:60-75, :99-164): an int, a list / range / 1-D array (one receiver or transmitter per
sample) or a 2-D array [samples, receivers-or-transmitters per sample]; defaults: BS 0, all users.
"""
from __future__ import annotations

from typing import Optional

import numpy as np

from . import channels as _ch
from .params import ChannelGenParameters


class DeepMIMOSionnaAdapter:
    def __init__(self, datasets, params: Optional[ChannelGenParameters] = None, bs_idx=None, ue_idx=None, *, device=None):
        if hasattr(datasets, "datasets"):
            datasets = datasets.datasets
        elif not isinstance(datasets, (list, tuple)):
            datasets = [datasets]
        self.datasets = list(datasets)
        p = (params if params is not None else ChannelGenParameters()).deepcopy()
        p["freq_domain"] = 0                                       # the adapter consumes time-domain channels (:76-77, :195)
        self.params = p
        self.device = device
        self.bs_idx = self._verify_idx(np.array([[0]]) if bs_idx is None else bs_idx)
        n_ue = int(np.asarray(self.datasets[0]["power"]).shape[0])
        self.ue_idx = self._verify_idx(np.arange(n_ue) if ue_idx is None else ue_idx)
        ue_shape = np.asarray(p["ue_antenna"]["shape"]).ravel()[:2]
        bs_shape = np.asarray(p["bs_antenna"]["shape"]).ravel()[:2]
        self.num_rx_ant = int(ue_shape[0] * ue_shape[1])
        self.num_tx_ant = int(bs_shape[0] * bs_shape[1])
        self.num_samples_bs = self.bs_idx.shape[0]
        self.num_samples_ue = self.ue_idx.shape[0]
        self.num_samples = self.num_samples_bs * self.num_samples_ue
        self.num_rx = self.ue_idx.shape[1]
        self.num_tx = self.bs_idx.shape[1]
        n_cols = int(np.asarray(self.datasets[0]["power"]).shape[1])
        self.num_paths = min(int(p["num_paths"]), n_cols)
        self.num_time_steps = 1
        self.ch_shape = (self.num_rx, self.num_rx_ant, self.num_tx, self.num_tx_ant, self.num_paths, self.num_time_steps)
        self.t_shape = (self.num_rx, self.num_tx, self.num_paths)
        self._cache = None

    # -- index handling, as the reference (:99-164)
    @staticmethod
    def _verify_idx(idx) -> np.ndarray:
        if isinstance(idx, (int, np.integer)):
            idx = np.array([[int(idx)]])
        elif isinstance(idx, (list, range)):
            idx = np.array(idx)
        elif not isinstance(idx, np.ndarray):
            raise TypeError("The index input type must be an integer, list, or numpy array!")
        if idx.ndim == 1:
            idx = idx.reshape((-1, 1))
        elif idx.ndim != 2:
            raise ValueError("The index input must be integer, vector or 2D matrix!")
        return idx

    def __len__(self) -> int:
        return self.num_samples

    # -- device side
    def _per_bs(self, b: int):
        """(H [n, M_r, M_t, P], tau [n, P]) CUDA tensors of base station `b`."""
        torch = _ch._torch()
        plan, _ = _ch.make_plan(self.datasets[b], self.params, device=self.device, warn=False)
        n = plan.n_users
        H = plan.alloc_out()
        tau = torch.empty((n, self.num_paths), dtype=torch.float32, device=plan.device)
        if n:
            plan.run(H, 0, n, {"tau": tau})
        return H, tau

    def arrays(self, out: str = "numpy"):
        """All samples at once: a complex64 [num_samples, *ch_shape], tau float32 [num_samples, *t_shape]; sample s = i * num_samples_bs + j
        for UE sample i and BS sample j (the reference's loop order, :175-176).  `out='torch'` keeps them on the device."""
        torch = _ch._torch()
        per_bs = {int(b): self._per_bs(int(b)) for b in np.unique(self.bs_idx)}
        dev = next(iter(per_bs.values()))[0].device
        a = torch.zeros((self.num_samples_ue, self.num_samples_bs) + self.ch_shape, dtype=torch.complex64, device=dev)
        tau = torch.zeros((self.num_samples_ue, self.num_samples_bs) + self.t_shape, dtype=torch.float32, device=dev)
        ue = torch.as_tensor(self.ue_idx, device=dev, dtype=torch.long)
        for j in range(self.num_samples_bs):
            for j_ch in range(self.num_tx):
                H, t = per_bs[int(self.bs_idx[j, j_ch])]
                for i_ch in range(self.num_rx):
                    a[:, j, i_ch, :, j_ch, :, :, 0] = H[ue[:, i_ch]]           # a[i_ch, :, j_ch, :, :, 0] = channel[i_ue]   (:195)
                    tau[:, j, i_ch, j_ch, :] = t[ue[:, i_ch]]                   # tau[i_ch, j_ch, :num_paths] = ToA          (:196-198)
        a = a.reshape((self.num_samples,) + self.ch_shape)
        tau = tau.reshape((self.num_samples,) + self.t_shape)
        if out == "torch":
            return a, tau
        return a.cpu().numpy(), tau.cpu().numpy()

    def __call__(self):
        if self._cache is None:
            self._cache = self.arrays()
        a, tau = self._cache
        for s in range(self.num_samples):
            yield a[s], tau[s]
