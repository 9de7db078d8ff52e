"""Import the UNMODIFIED reference model (BlockBasedImgCompLossyNetv9) in this container.

TEST INFRASTRUCTURE ONLY; needs /root/reference (absent on the GPU box), so only fixture
generation (tests/golden/make_golden.py) and the optional `needs_reference` tests use it.

The reference's package __init__ files import every sibling module (graphs/layers/__init__.py:6-10,
utils/__init__.py:6-10) and so drag in matplotlib/easydict/bjontegaard/compressai, none of which
is installed.  We therefore register bare namespace packages and load the few files the codec path
needs BY PATH, with oracle/ref_shim/compressai standing in for CompressAI.
"""
import importlib.util
import os
import sys
import types

REF = os.environ.get("LBIC_REFERENCE_ROOT", "/root/reference")
_HERE = os.path.dirname(os.path.abspath(__file__))
_REPO = os.path.dirname(os.path.dirname(_HERE))


def available() -> bool:
    return os.path.isfile(os.path.join(REF, "graphs/models/BlockBasedImgCompLossy_net.py"))


def _load(name, relpath):
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, relpath))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def _ns(name):
    if name not in sys.modules:
        m = types.ModuleType(name)
        m.__path__ = []
        sys.modules[name] = m
    return sys.modules[name]


def load():
    """Returns the reference module graphs.models.BlockBasedImgCompLossy_net."""
    if not available():
        raise FileNotFoundError(f"reference tree not found at {REF}")
    for p in (_REPO, _HERE):
        if p not in sys.path:
            sys.path.insert(0, p)
    import compressai  # the stand-in in this directory

    assert os.path.dirname(compressai.__file__) == os.path.join(_HERE, "compressai"), compressai.__file__
    _ns("utils")
    bound_ops = _load("utils.bound_ops", "utils/bound_ops.py")
    _load("utils.parametrizers", "utils/parametrizers.py")
    ops = types.ModuleType("compressai.ops")
    ops.LowerBound = bound_ops.LowerBound
    sys.modules["compressai.ops"] = ops
    compressai.ops = ops
    _ns("graphs"); _ns("graphs.layers"); _ns("graphs.models")
    gdn = _load("graphs.layers.gdn_compressai", "graphs/layers/gdn_compressai.py")
    layers = types.ModuleType("compressai.layers")
    layers.GDN, layers.GDN1 = gdn.GDN, gdn.GDN1
    sys.modules["compressai.layers"] = layers
    compressai.layers = layers
    ent = _load("graphs.layers.entropy_layers_cai", "graphs/layers/entropy_layers_cai.py")
    em = types.ModuleType("compressai.entropy_models")
    em.GaussianConditional, em.EntropyBottleneck, em.EntropyModel = (
        ent.GaussianConditional, ent.EntropyBottleneck, ent.EntropyModel)
    sys.modules["compressai.entropy_models"] = em
    compressai.entropy_models = em
    _load("graphs.layers.masked_conv2d", "graphs/layers/masked_conv2d.py")
    return _load("graphs.models.BlockBasedImgCompLossy_net", "graphs/models/BlockBasedImgCompLossy_net.py")
