"""packnet_sfm_b200: B200-native (sm_100a) implementation of the PackNet-SfM self-supervised hot path.

Public surface mirrors the reference plug-in points:
  packnet_sfm_b200.networks.PackNet01               <- packnet_sfm/networks/depth/PackNet01.py
  packnet_sfm_b200.losses.MultiViewPhotometricLoss  <- packnet_sfm/losses/multiview_photometric_loss.py
  packnet_sfm_b200.dropin.install()                 registers both under the reference's module paths
The compute lives in csrc/ (hand-written CUDA behind the C-ABI in include/packnet_b200.h)."""
__version__ = "0.1.0"
