# Shadows Diff-Reg-4dmatch/models/procrustes.py (imported by models/pipeline.py:5, models/transformer.py:8)
from diffreg_b200.procrustes import SoftProcrustesLayer  # noqa: F401
