"""Scenario loader -> Dataset: the step in front of the channel path.

Reads the reference's v4 on-disk format: one MATLAB file per matrix and (TX set, TX index, RX set) named
`{key}_t{tx_set:03}_tx{tx_idx:03}_r{rx_set:03}.mat` (deepmimo/general_utils.py:296-323) holding the array under `key`,
plus `params.json` whose `txrx_sets` entry lists the TX/RX sets (deepmimo/generator/core.py:111-128, :261-338).  Semantics
follow `_load_tx_rx_raydata` (core.py:186-258): rows are filtered by the selected receiver indices (all matrices but
`tx_pos`), path matrices are trimmed to `max_paths` columns; one Dataset per (TX set, RX set, TX index), several pairs
come back as a MacroDataset (core.py:139-183).  Path matrices are returned float32 and C-contiguous (the storage type,
deepmimo/consts.py:65), optionally in pinned host memory so the H2D copy of `compute_channels` is asynchronous.

`save_scenario` writes the same format from in-memory datasets (used by the tests and for synthetic scenarios).
"""
from __future__ import annotations

import json
import os
from typing import Dict, Iterable, List, Optional, Union

import numpy as np

from .dataset import Dataset, MacroDataset

MATRIX_KEYS = ("aoa_az", "aoa_el", "aod_az", "aod_el", "power", "phase", "delay", "rx_pos", "tx_pos", "inter", "inter_pos")
PATH_MATRICES = ("aoa_az", "aoa_el", "aod_az", "aod_el", "power", "phase", "delay", "inter")
MAX_PATHS = 25


def txrx_str_id(tx_set: int, tx_idx: int, rx_set: int) -> str:
    return f"t{tx_set:03}_tx{tx_idx:03}_r{rx_set:03}"


def mat_filename(key: str, tx_set: int, tx_idx: int, rx_set: int) -> str:
    return f"{key}_{txrx_str_id(tx_set, tx_idx, rx_set)}.mat"


def _pinned(a: np.ndarray) -> np.ndarray:
    import torch
    if not torch.cuda.is_available():
        return a
    return torch.from_numpy(a).pin_memory().numpy()


class _DeviceFeed:
    """Host -> device pipeline of the loader: the H2D copy of matrix k (from one of two pinned staging buffers, on a copy stream)
    runs while scipy parses matrix k + 1.  `finish()` makes the caller's stream wait for the copies."""

    def __init__(self, device):
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("load_scenario(device=...) needs a CUDA device")
        self.torch = torch
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(device=self.device)
        self.stage = [None, None]
        self.busy = [None, None]
        self.k = 0
        self.bytes = 0

    def put(self, a: np.ndarray):
        torch = self.torch
        b = self.k & 1
        self.k += 1
        if self.busy[b] is not None:
            self.busy[b].synchronize()                       # the copy that last read this staging buffer has finished
        if self.stage[b] is None or self.stage[b].numel() < a.nbytes:
            self.stage[b] = torch.empty(max(a.nbytes, 1 << 20), dtype=torch.uint8, pin_memory=True)
        st = self.stage[b][: a.nbytes].view(torch.float32).view(a.shape)
        st.copy_(torch.from_numpy(a))                        # host memcpy into page-locked memory
        out = torch.empty(a.shape, dtype=torch.float32, device=self.device)
        with torch.cuda.stream(self.stream):
            out.copy_(st, non_blocking=True)
            self.busy[b] = torch.cuda.Event()
            self.busy[b].record(self.stream)
        self.bytes += a.nbytes
        return out

    def finish(self):
        self.torch.cuda.current_stream(self.device).wait_stream(self.stream)


def load_tx_rx_raydata(folder: str, tx_set_id: int, rx_set_id: int, tx_idx: int, rx_idxs=None, max_paths: int = MAX_PATHS,
                       matrices: Union[str, Iterable[str]] = "all", pin: bool = False, device=None, _feed=None) -> Dict[str, np.ndarray]:
    """Matrices of one TX-RX pair (core.py:186-258).  `device='cuda'` returns the seven path matrices as float32 CUDA tensors fed
    through the parse / H2D pipeline (`compute_channels` takes them as they are; positions and interaction data stay on the host)."""
    import scipy.io
    feed = _feed if _feed is not None else (_DeviceFeed(device) if device is not None else None)
    if matrices == "all":
        wanted = MATRIX_KEYS
    else:
        wanted = tuple(matrices or ())
        bad = set(wanted) - set(MATRIX_KEYS)
        if bad:
            raise ValueError(f"Invalid matrix names: {bad}. Valid names are: {set(MATRIX_KEYS)}")
    out: Dict[str, np.ndarray] = {}
    for key in MATRIX_KEYS:
        if key not in wanted:
            continue
        path = os.path.join(folder, mat_filename(key, tx_set_id, tx_idx, rx_set_id))
        if not os.path.exists(path):
            print(f"File {path} could not be found")
            continue
        a = scipy.io.loadmat(path)[key]
        if key != "tx_pos" and rx_idxs is not None:
            a = a[rx_idxs]
        if key not in ("rx_pos", "tx_pos"):
            a = a[:, :max_paths, ...]
        if key in PATH_MATRICES or key in ("rx_pos", "tx_pos"):
            a = np.ascontiguousarray(a, dtype=np.float32)
        else:
            a = np.ascontiguousarray(a)
        if feed is not None and key in PATH_MATRICES and key != "inter":
            out[key] = feed.put(a)
        else:
            out[key] = _pinned(a) if (pin and key in PATH_MATRICES) else a
    if feed is not None and _feed is None:
        feed.finish()
    return out


def _resolve_sets(sets, txrx: dict, role: str) -> Dict[int, np.ndarray]:
    """tx_sets / rx_sets argument -> {set id: indices} (core.py:261-338)."""
    flag = "is_tx" if role == "tx" else "is_rx"
    valid = [txrx[k] for k in sorted(txrx) if txrx[k].get(flag)]
    ids = [s["id"] for s in valid]
    npts = {s["id"]: int(s["num_points"]) for s in valid}
    name = "Tx" if role == "tx" else "Rx"
    if isinstance(sets, str):
        if sets != "all":
            raise ValueError(f"String '{sets}' not understood. Only 'all' is allowed")
        return {i: np.arange(npts[i]) for i in ids}
    if isinstance(sets, (list, tuple)):
        for i in sets:
            if i not in ids:
                raise ValueError(f"{name} set {i} not in allowed sets {ids}")
        return {i: np.arange(npts[i]) for i in sets}
    if isinstance(sets, dict):
        out = {}
        for i, idxs in sets.items():
            if i not in ids:
                raise ValueError(f"{name} set {i} not in allowed sets {ids}")
            if isinstance(idxs, str):
                if idxs != "all":
                    raise ValueError(f"String '{idxs}' not recognized for tx/rx indices")
                idxs = np.arange(npts[i])
            idxs = np.asarray(idxs)
            if idxs.size and (idxs.min() < 0 or idxs.max() >= npts[i]):
                raise ValueError(f"Some indices of {name} set {i} are outside [0, {npts[i]})")
            out[i] = idxs
        return out
    raise ValueError("tx_sets / rx_sets must be 'all', a list of set ids or a dict {set id: indices}")


def load_scenario(folder: str, max_paths: int = MAX_PATHS, tx_sets="all", rx_sets="all", matrices="all", pin: bool = False, device=None):
    """Load a scenario folder into a Dataset (one TX-RX pair) or a MacroDataset (several), core.py:63-183.
    `device='cuda'` (or 'cuda:1', ...) streams the path matrices to that GPU while the following files are parsed."""
    params_path = os.path.join(folder, "params.json")
    if not os.path.exists(params_path):
        raise ValueError(f"Parameters file not found in {folder}")
    with open(params_path) as f:
        params = json.load(f)
    if int(params.get("scene", {}).get("num_scenes", 1)) > 1:
        raise NotImplementedError("Dynamic scenarios not implemented yet")          # core.py:118-120
    txrx = params["txrx_sets"]
    tx = _resolve_sets(tx_sets, txrx, "tx")
    rx = _resolve_sets(rx_sets, txrx, "rx")
    datasets: List[Dataset] = []
    feed = _DeviceFeed(device) if device is not None else None
    for tx_set_id, tx_idxs in tx.items():
        for rx_set_id, rx_idxs in rx.items():
            for tx_idx in tx_idxs:
                d = load_tx_rx_raydata(folder, tx_set_id, rx_set_id, int(tx_idx), rx_idxs, max_paths, matrices, pin, _feed=feed)
                d["txrx"] = {"tx_set_id": tx_set_id, "rx_set_id": rx_set_id, "tx_idx": int(tx_idx)}
                d["name"] = os.path.basename(os.path.normpath(folder))
                datasets.append(Dataset(d))
    if feed is not None:
        feed.finish()
    if not datasets:
        raise ValueError("no TX-RX pair selected")
    return datasets[0] if len(datasets) == 1 else MacroDataset(datasets)


def save_scenario(folder: str, pairs: Dict[tuple, dict], *, name: Optional[str] = None) -> str:
    """Write datasets in the v4 format.  `pairs` maps (tx_set, tx_idx, rx_set) -> dict of matrices."""
    import scipy.io
    os.makedirs(folder, exist_ok=True)
    sets: Dict[int, dict] = {}
    for (ts, ti, rs), data in pairs.items():
        n_rx = int(np.asarray(data["power"]).shape[0])
        t = sets.setdefault(ts, {"id": ts, "name": f"tx_set_{ts}", "is_tx": False, "is_rx": False, "num_points": 0})
        t["is_tx"] = True
        t["num_points"] = max(t["num_points"], ti + 1)
        r = sets.setdefault(rs, {"id": rs, "name": f"rx_set_{rs}", "is_tx": False, "is_rx": False, "num_points": 0})
        r["is_rx"] = True
        r["num_points"] = max(r["num_points"], n_rx) if rs != ts else max(r["num_points"], n_rx, ti + 1)
        for key in MATRIX_KEYS:
            if key in data:
                scipy.io.savemat(os.path.join(folder, mat_filename(key, ts, ti, rs)), {key: np.asarray(data[key])})
    params = {"version": "4.0.0a3", "name": name or os.path.basename(os.path.normpath(folder)),
              "scene": {"num_scenes": 1}, "txrx_sets": {f"txrx_set_{i}": s for i, s in sorted(sets.items())}}
    with open(os.path.join(folder, "params.json"), "w") as f:
        json.dump(params, f, indent=1)
    return folder
