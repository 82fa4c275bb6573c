# Shadows Diff-Reg-2d3d/experiments/<exp>/matching.py (imported by model.py:29-30)
from diffreg_b200.matching import Matching2D3D as Matching, log_optimal_transport  # noqa: F401
