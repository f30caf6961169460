"""Import the UNMODIFIED reference modules staged under baseline/_ref (baseline/stage_reference.py) -- test / bench infrastructure.

Two entry points:

  reference_modules()            the reference's own PointAggregator, positional_encoding, ray_march, alpha_ray_march, ray generation,
                                 render / blend / tone-map finders (CPU or cuda torch code; what tests/golden/make_golden.py pins the
                                 oracle to, and what bench.py --impl reference times)
  volumetric_model(swap=True)    models/neural_points_volumetric_model.py imported as it is, with INTEGRATION.md's three-import swap
                                 applied from outside: sgnerf_b200's NeuralPoints / PointAggregator stand where the file's own
                                 `from .neural_points.neural_points import NeuralPoints` and
                                 `from .aggregators.point_aggregators import PointAggregator` resolve, and the module global `ray_march`
                                 (star-imported from base_rendering_model) is rebound to sgnerf_b200's.

Packages the file imports at module level that this image does not have (MinkowskiEngine, imageio, matplotlib) or that only BPNet
needs (models.bpneter.bpnet, bpnet_dataset.*) are replaced by empty stand-ins: nothing on the render path touches them.
"""
import importlib
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STAGED = os.path.join(ROOT, "baseline", "_ref")


def available():
    return os.path.exists(os.path.join(STAGED, "models", "neural_points_volumetric_model.py"))


class _Anything:
    """Attribute sink: any name resolves to a class that can be instantiated, called or subclassed."""
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, name):
        return _Anything()


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    m.__dict__.update(attrs)

    def _missing(attr):                               # PEP 562: any other public attribute is a stand-in class
        if attr.startswith("__"):
            raise AttributeError(attr)
        return _Anything
    m.__getattr__ = _missing
    m.__path__ = []                                   # lets `import a.b` treat it as a package
    sys.modules[name] = m
    return m


def _prepare():
    import scipy.special
    import torch  # noqa: F401  (before any stand-in module exists: torch's import inspects sys.modules)
    import torchvision  # noqa: F401
    for fn in ("sph_harm", "lpmn"):                   # removed from recent scipy; only the unused SH kernel calls them
        if not hasattr(scipy.special, fn):
            setattr(scipy.special, fn, lambda *a, **k: None)
    for name in ("MinkowskiEngine", "imageio", "matplotlib", "matplotlib.cm", "matplotlib.pyplot", "bpnet_dataset",
                 "bpnet_dataset.augmentation_2d", "bpnet_dataset.voxelizer"):
        try:
            importlib.import_module(name)
        except Exception:
            _stub(name)
    if STAGED not in sys.path:
        sys.path.insert(0, STAGED)
    # a top-level `utils` / `models` of another project must not shadow the staged packages
    for pkg in ("utils", "models"):
        m = sys.modules.get(pkg)
        if m is not None and not str(getattr(m, "__file__", "") or "").startswith(STAGED) and \
                not any(str(p).startswith(STAGED) for p in getattr(m, "__path__", [])):
            for k in [k for k in sys.modules if k == pkg or k.startswith(pkg + ".")]:
                del sys.modules[k]


def reference_modules():
    if not available():
        raise FileNotFoundError("baseline/_ref is not staged (python baseline/stage_reference.py in the build container)")
    _prepare()
    from models.aggregators.point_aggregators import PointAggregator
    from models.helpers.networks import positional_encoding
    from models.rendering import diff_ray_marching as drm
    from models.rendering import diff_render_func as drf
    return types.SimpleNamespace(PointAggregator=PointAggregator, positional_encoding=positional_encoding, ray_march=drm.ray_march,
                                 alpha_ray_march=drm.alpha_ray_march, near_far_linear_ray_generation=drm.near_far_linear_ray_generation,
                                 find_render_function=drf.find_render_function, find_blend_function=drf.find_blend_function,
                                 find_tone_map=drf.find_tone_map)


def volumetric_model(swap=True):
    """The reference's models.neural_points_volumetric_model module (NeuralPointsRayMarching, NeuralPointsVolumetricModel)."""
    if not available():
        raise FileNotFoundError("baseline/_ref is not staged (python baseline/stage_reference.py in the build container)")
    _prepare()
    from sgnerf_b200 import modules as ours
    _stub("models.bpneter")
    _stub("models.bpneter.bpnet")
    if swap:
        # where the file's relative imports resolve: the two classes come from sgnerf_b200
        _stub("models.neural_points")
        sys.modules.pop("models.neural_points.neural_points", None)
        sys.modules.pop("models.neural_points_volumetric_model", None)
        _stub("models.neural_points.neural_points", NeuralPoints=ours.NeuralPoints)
        real_agg = sys.modules.get("models.aggregators.point_aggregators")
        sys.modules["models.aggregators.point_aggregators"] = types.SimpleNamespace(PointAggregator=ours.PointAggregator)
        try:
            mod = importlib.import_module("models.neural_points_volumetric_model")
        finally:
            if real_agg is not None:
                sys.modules["models.aggregators.point_aggregators"] = real_agg
            else:
                sys.modules.pop("models.aggregators.point_aggregators", None)
        mod.ray_march, mod.alpha_ray_march = ours.ray_march, ours.alpha_ray_march
        assert mod.NeuralPoints is ours.NeuralPoints and mod.PointAggregator is ours.PointAggregator
    else:
        mod = importlib.import_module("models.neural_points_volumetric_model")
    return mod
