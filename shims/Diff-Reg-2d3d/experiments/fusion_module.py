# Shadows Diff-Reg-2d3d/experiments/<exp>/fusion_module.py (optional, inference only; imported by model.py:21)
from diffreg_b200.fusion import CrossModalFusionModule  # noqa: F401
