"""libwave_b200 - B200-native drop-in for libwave's wave_matching registration hot path.

Product code: the CUDA library (csrc/, built to libwavecu.so behind include/wavecu.h) and the
host-side mirror of the reference's matcher interface (matching.py).  ``synth`` is the synthetic
scan generator used by the tests and the bench.  Nothing here imports ``oracle/``.
"""
from .matching import (EST_POINT_TO_PLANE, EST_SVD, INFO_CENSI, INFO_LUM, INFO_LUMOLD, GICPMatcher,  # noqa: F401
                       GICPMatcherParams, ICPMatcher,
                       ICPMatcherParams, Matcher, NDTMatcher, NDTMatcherParams, NDT_LS_MORE_THUENTE, NDT_LS_PCL18, SEARCH_TILED, SEARCH_TREE,
                       NearestNeighbour, voxel_grid)
