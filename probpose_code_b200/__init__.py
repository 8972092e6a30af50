"""probpose_code_b200: B200-native ProbPose top-down inference hot path."""
__version__ = "0.1.0"
