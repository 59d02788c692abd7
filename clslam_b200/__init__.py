"""clslam_b200: B200-native (sm_100a) implementation of CL-SLAM's online-adaptation hot path behind
the reference's ``DepthPosePrediction`` API.  See DESIGN.md / INTEGRATION.md."""
from .config import Config, DatasetConfig
from .depth_pose_prediction import DepthPosePrediction

__all__ = ['DepthPosePrediction', 'Config', 'DatasetConfig']
