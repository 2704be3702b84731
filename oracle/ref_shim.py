"""Import shim for the UNMODIFIED reference (anibali/margipose) -- TEST INFRASTRUCTURE ONLY.

Only usable where /root/reference exists (the build container); it is used by
tests/golden/make_golden.py to generate golden vectors and by
tests/test_oracle_pin.py to pin oracle/ against the real reference.  Nothing on
the product path (margipose_b200/) and nothing that runs on the GPU box may
import this module.

The reference's model file imports a data stack whose third-party dependencies
are absent here.  We pre-register inert stand-ins for
them; none is touched on the numerical path.
"""
import os
import sys
import types

REFERENCE_SRC = os.environ.get('MARGIPOSE_REFERENCE_SRC', '/root/reference/src')


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_SRC, 'margipose'))


class _Anything(types.ModuleType):
    """A module whose every attribute is a harmless callable/class."""

    def __getattr__(self, name):
        if name.startswith('__'):
            raise AttributeError(name)
        value = type(name, (), {'__init__': lambda self, *a, **k: None})
        setattr(self, name, value)
        return value


def _stub(name):
    if name not in sys.modules:
        mod = _Anything(name)
        mod.__path__ = []
        sys.modules[name] = mod
    return sys.modules[name]


class _Version:
    def __init__(self, text):
        self.parts = tuple(int(p) for p in str(text).split('.')[:3])

    def __str__(self):
        return '.'.join(str(p) for p in self.parts)


class _Spec:
    """Caret specs only ('^6.0.0'), which is all the reference uses."""

    def __init__(self, text):
        assert text.startswith('^')
        self.base = _Version(text[1:])

    def __contains__(self, version):
        v = version if isinstance(version, _Version) else _Version(version)
        return v.parts[0] == self.base.parts[0] and v.parts >= self.base.parts


_loaded = None


def load_reference():
    """Returns the imported reference modules as a namespace."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not reference_available():
        raise RuntimeError('reference source tree not found at ' + REFERENCE_SRC)
    import torch
    import torchvision

    for name in ['pose3d_utils', 'pose3d_utils.camera', 'pose3d_utils.coords',
                 'pose3d_utils.skeleton_normaliser', 'pose3d_utils.transformers',
                 'pose3d_utils.transforms', 'pretrainedmodels', 'pretrainedmodels.models',
                 'pretrainedmodels.models.inceptionv4', 'h5py', 'importlib_resources',
                 'matplotlib', 'matplotlib.pylab', 'matplotlib.pyplot', 'mpl_toolkits',
                 'mpl_toolkits.mplot3d', 'plotly', 'plotly.graph_objs']:
        _stub(name)
    six = _stub('torch._six')
    six.string_classes = (str, bytes)
    six.int_classes = (int,)
    sv = _stub('semantic_version')
    sv.Version = _Version
    sv.Spec = _Spec

    # pretrained=True needs the network; use the architecture with its default init.
    for arch in ['resnet18', 'resnet34', 'resnet50']:
        orig = getattr(torchvision.models, arch)
        if getattr(orig, '_shimmed', False):
            continue

        def make(orig):
            def ctor(pretrained=False, **kwargs):
                return orig(weights=None, **kwargs)
            ctor._shimmed = True
            return ctor
        setattr(torchvision.models, arch, make(orig))

    if REFERENCE_SRC not in sys.path:
        sys.path.insert(0, REFERENCE_SRC)
    import margipose.dsntnn as ref_dsntnn
    import margipose.models.margipose_model as ref_model
    import margipose.models as ref_models
    from margipose.data.skeleton import CanonicalSkeletonDesc
    _loaded = types.SimpleNamespace(dsntnn=ref_dsntnn, model=ref_model, models=ref_models,
                                    CanonicalSkeletonDesc=CanonicalSkeletonDesc)
    return _loaded
