# Shadows Diff-Reg-4dmatch/models/matching.py (imported by models/pipeline.py:4, models/transformer.py:7, lib/tester.py:6)
from diffreg_b200.matching import Matching, log_optimal_transport, mutual_topk_select  # noqa: F401
