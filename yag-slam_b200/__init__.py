"""yag-slam_b200: B200-native correlative scan matcher behind yag_slam's scan_matching API.

Replaces Karto's ScanMatcher::MatchScan hot path (reference boundary:
yag_slam/scan_matching.py:32-42 -> karto_scanmatcher.Wrapper.match_scan) and the numba
ray-walk (yag_slam/raytracing.py:63-92) with hand-written sm_100a CUDA kernels called
through the C ABI declared in include/ysm.h. There is no CPU fallback: importing the
compute entry points without the built CUDA library raises.
"""
__version__ = "0.1.0"
