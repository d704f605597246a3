"""deepmimo_b200 -- B200-native (sm_100a) channel generation for DeepMIMO.

Drop-in for the reference's `dataset.compute_channels(dm.ChannelGenParameters())` hot path
(jmoraispk/DeepMIMO v4.0.0a3, deepmimo/generator/dataset.py:224).  Hand-written CUDA kernels behind a
C ABI (include/dmk.h, deepmimo_b200/libdmk.so); Python keeps the reference's parameter and dataset
interface.  No Triton, no multi-backend dispatch, no CPU fallback.
"""
from .params import ChannelGenParameters, ChannelParameters, DotDict
from .dataset import Dataset, MacroDataset
from .channels import (ChannelInfo, ChannelPlan, ChannelSpec, compute_channels, iter_channels, make_plan,
                       parse_spec)
from .beams import beam_amplitude, steering_vec
from .install import install, uninstall
from .loader import load_scenario, load_tx_rx_raydata, save_scenario

__version__ = "0.1.0"
__all__ = ["ChannelGenParameters", "ChannelParameters", "DotDict", "Dataset", "MacroDataset", "ChannelInfo",
           "ChannelPlan", "ChannelSpec", "compute_channels", "iter_channels", "make_plan", "parse_spec",
           "beam_amplitude", "steering_vec", "install", "uninstall", "load_scenario", "load_tx_rx_raydata", "save_scenario"]
