# Shadows Diff-Reg-4dmatch/models/transformer.py (optional, inference only; imported by models/pipeline.py:3)
from diffreg_b200.transformer import GeometryAttentionLayer, RepositioningTransformer  # noqa: F401
