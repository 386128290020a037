"""dynamic-rs-segmentation_b200 -- B200-native (sm_100a) hot path of keillernogueira/dynamic-rs-segmentation.

Import as ``import drs_b200`` (alias module at the repo root).  Layout:
  csrc/       hand-written CUDA kernels + the C-ABI (libdrs.so, include/drs.h)
  lib.py      ctypes binding (no fallback: raises when libdrs.so is missing)
  session.py  ``Session`` -- the drop-in for the reference's ``tf.Session.run`` seam
  nets.py     specs / TF variable names / initialisation of the four dilated FCNs
  host.py     host-side mirror of the reference's L2/L3 logic (batch selection, patch-size policy,
              augmentation decisions, sliding-window grid) -- same names and argument meaning
  synth.py    synthetic Vaihingen / Potsdam / contest / coffee shaped scenes
  dist.py     one-process-per-GPU plumbing: stripe-sharded inference, data-parallel training
"""
from . import lib, nets  # noqa: F401
from .session import Session, grid_positions, npz_read, npz_write  # noqa: F401

__all__ = ["Session", "grid_positions", "npz_read", "npz_write", "lib", "nets"]
