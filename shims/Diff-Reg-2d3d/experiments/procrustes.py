# Shadows Diff-Reg-2d3d/experiments/<exp>/procrustes.py (imported by model.py:31)
from diffreg_b200.procrustes import SoftProcrustesLayer  # noqa: F401
