"""sgnerf_b200 -- B200-native (sm_100a) implementation of SG-NeRF's per-ray render hot path:
neural-point query, K-neighbour feature aggregation (+MLPs) and alpha compositing, behind the reference's
NeuralPoints / PointAggregator / ray_march module API.  See DESIGN.md and include/sgnerf_b200.h.

Importing the package does not need a GPU; every compute call does (there is no CPU or PyTorch fallback).
"""
__version__ = "0.1.0"

from . import _lib, dist, modules, ops, pipeline, synth  # noqa: F401
