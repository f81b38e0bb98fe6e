"""v2a_b200: B200-native (sm_100a) hot paths of video-to-action-release.

Importable as ``v2a_b200`` (the alias package at the repo root points its
``__path__`` here; the directory name itself is not a valid Python identifier).

Public surface mirrors the reference's call surface:
  * ``GoalGaussianDiffusion`` / ``Unet_Libero``  — video sampling path
  * ``ConditionalUnet1D``                         — policy network
The compute runs in hand-written CUDA loaded through the C ABI in
``include/v2a_b200.h``; there is no CPU or PyTorch-eager fallback.
"""
__version__ = "0.1.0"
