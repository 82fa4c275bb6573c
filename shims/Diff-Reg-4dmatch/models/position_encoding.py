# Shadows Diff-Reg-4dmatch/models/position_encoding.py (optional; imported by models/matching.py:4, models/transformer.py:6)
from diffreg_b200.position_encoding import VolumetricPositionEncoding  # noqa: F401
