"""Synthetic inputs shared by the parity tests (re-exported from the package)."""
from unflow_b200.synthetic import *  # noqa: F401,F403
from unflow_b200.synthetic import KITTI_NORMALIZATION, KITTI_PARAMS, image_pair, level_inputs  # noqa: F401
