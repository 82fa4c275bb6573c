"""Import the reference's own modules (unmodified) for one flavour -- test / bench infrastructure only.

Root: /root/reference when present (the build container), else the copy under oracle/_ref (made by
oracle/make_ref.py; it travels to the GPU box).  Shims, as SURVEY.md 8c lists them:
  * the 2D-3D matching.py imports `vision3d.ops`, whose compiled extension is absent: a stub `vision3d.ops`
    exposing the reference's own pure-torch mutual_topk_select.py (loaded by file path) is registered first;
  * mutual_topk_select hard-codes `.cuda()`: `cpu_cuda()` makes Tensor.cuda the identity while it runs on CPU.
"""
import importlib
import importlib.util
import os
import sys
import types
from types import SimpleNamespace

HERE = os.path.dirname(os.path.abspath(__file__))
EXP_2D3D = "Diff-Reg-2d3d/experiments/2d3dmatr.rgbdv2.stage4.level3.stage1"


def reference_root():
    """(path, kind) of the reference sources, or (None, None)."""
    env = os.environ.get("DRG_REFERENCE_ROOT")
    for root, kind in ((env, "checkout"), ("/root/reference", "checkout"), (os.path.join(HERE, "_ref"), "oracle/_ref")):
        if root and os.path.exists(os.path.join(root, "Diff-Reg-4dmatch", "models", "matching.py")):
            return root, kind
    return None, None


def available():
    return reference_root()[0] is not None


def _purge(prefixes):
    for k in list(sys.modules):
        if any(k == p or k.startswith(p + ".") for p in prefixes):
            del sys.modules[k]


_PREFIXES = ["models", "matching", "procrustes", "position_encoding", "lib", "datasets", "vision3d", "kernels"]


def load_flavour(name):
    """'4d' -> ns.matching / procrustes / pipeline / pe / transformer; '3d' -> matching / procrustes;
    '2d3d' -> matching / procrustes / mts.  Modules of a previously loaded flavour are dropped from sys.modules."""
    root, _ = reference_root()
    if root is None:
        raise RuntimeError("reference sources not found (neither /root/reference nor oracle/_ref)")
    _purge(_PREFIXES)
    for p in list(sys.path):
        if p.startswith(root):
            sys.path.remove(p)
    ns = SimpleNamespace(root=root)
    if name == "4d":
        sys.path.insert(0, os.path.join(root, "Diff-Reg-4dmatch"))
        ns.matching = importlib.import_module("models.matching")
        ns.procrustes = importlib.import_module("models.procrustes")
        ns.pe = importlib.import_module("models.position_encoding")
        ns.transformer = importlib.import_module("models.transformer")
        ns.pipeline = importlib.import_module("models.pipeline")
    elif name == "3d":
        sys.path.insert(0, os.path.join(root, "Diff-Reg-3dmatch"))
        ns.matching = importlib.import_module("models.matching")
        ns.procrustes = importlib.import_module("models.procrustes")
    elif name == "2d3d":
        spec = importlib.util.spec_from_file_location(
            "_ref_mts", os.path.join(root, "Diff-Reg-2d3d", "vision3d", "ops", "mutual_topk_select.py"))
        mts = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mts)
        v3d = types.ModuleType("vision3d")
        ops = types.ModuleType("vision3d.ops")
        ops.mutual_topk_select = mts.mutual_topk_select
        v3d.ops = ops
        sys.modules["vision3d"] = v3d
        sys.modules["vision3d.ops"] = ops
        sys.path.insert(0, os.path.join(root, EXP_2D3D))
        ns.matching = importlib.import_module("matching")
        ns.procrustes = importlib.import_module("procrustes")
        ns.mts = mts
    else:
        raise ValueError(name)
    return ns


def load_fusion():
    """The 2D-3D fusion module and the two vision3d layer files it is built from (transformer.py, embedding.py), unmodified.
    `vision3d/layers/__init__.py` imports the whole layer zoo (and with it the compiled `vision3d.ext`, absent here), so the
    packages `vision3d` / `vision3d.layers` are registered as empty namespaces and only the needed files are imported into
    them.  Returns ns.fusion (CrossModalFusionModule), ns.transformer (TransformerLayer ...), ns.embedding (FourierEmbedding)."""
    root, _ = reference_root()
    if root is None or not os.path.exists(os.path.join(root, "Diff-Reg-2d3d", "vision3d", "layers", "transformer.py")):
        raise RuntimeError("reference sources of the 2D-3D fusion module not found")
    _purge(_PREFIXES + ["fusion_module"])
    base = os.path.join(root, "Diff-Reg-2d3d")
    v3d = types.ModuleType("vision3d")
    v3d.__path__ = [os.path.join(base, "vision3d")]
    lay = types.ModuleType("vision3d.layers")
    lay.__path__ = [os.path.join(base, "vision3d", "layers")]
    sys.modules["vision3d"] = v3d
    sys.modules["vision3d.layers"] = lay
    v3d.layers = lay
    ns = SimpleNamespace(root=root)
    ns.transformer = importlib.import_module("vision3d.layers.transformer")
    ns.embedding = importlib.import_module("vision3d.layers.embedding")
    lay.TransformerLayer = ns.transformer.TransformerLayer
    lay.FourierEmbedding = ns.embedding.FourierEmbedding
    spec = importlib.util.spec_from_file_location("fusion_module", os.path.join(base, "experiments", EXP_2D3D.split("/")[-1], "fusion_module.py"))
    ns.fusion = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ns.fusion)
    return ns


def fusion_available():
    root, _ = reference_root()
    return root is not None and os.path.exists(os.path.join(root, "Diff-Reg-2d3d", "vision3d", "layers", "basic_layers", "builder.py"))


def unload():
    root, _ = reference_root()
    _purge(_PREFIXES)
    if root:
        for p in list(sys.path):
            if p.startswith(root):
                sys.path.remove(p)


class cpu_cuda:
    """Make ``tensor.cuda()`` a no-op so reference code with a literal .cuda() runs on CPU tensors."""

    def __enter__(self):
        import torch
        self._orig = torch.Tensor.cuda
        torch.Tensor.cuda = lambda self, *a, **k: self

    def __exit__(self, *a):
        import torch
        torch.Tensor.cuda = self._orig
