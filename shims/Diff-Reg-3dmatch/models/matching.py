# Shadows Diff-Reg-3dmatch/models/matching.py (imported by models/pipeline.py:4)
from diffreg_b200.matching import Matching, log_optimal_transport, mutual_topk_select  # noqa: F401
