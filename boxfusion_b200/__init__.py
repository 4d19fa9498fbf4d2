"""boxfusion_b200 - B200-native (sm_100a) implementation of BoxFusion's multi-view box-fusion hot path."""
__version__ = "0.1.0"
