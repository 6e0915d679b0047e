"""unipose_b200 - B200-native (sm_100a) UniPose hot path behind the reference's nn.Module API."""
__version__ = "0.1.0"
