# Shadows Diff-Reg-3dmatch/models/procrustes.py: the 3DMatch variant uses the padded lengths (procrustes.py:61-62)
from diffreg_b200.procrustes import SoftProcrustesLayer3DMatch as SoftProcrustesLayer  # noqa: F401
