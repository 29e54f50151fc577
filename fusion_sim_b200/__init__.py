"""fusion_sim_b200 -- B200-native particle-step engine behind the simulation API of
kcdodd/fusion-sim (empic.makeCylindricalParticlePusher, public/javascripts/empic.js:30).

Contents: csrc/ (hand-written CUDA for sm_100a + the C ABI of include/fusionsim.h) and the
headless host driver that mirrors the reference's JavaScript object.  Nothing else.
"""
from ._lib import Error, LIB_PATH  # noqa: F401
from .pusher import CylindricalParticlePusher, makeCylindricalParticlePusher  # noqa: F401

__all__ = ["makeCylindricalParticlePusher", "CylindricalParticlePusher", "Error", "LIB_PATH"]
