"""Import shim that lets the UNMODIFIED reference run in the build container.

Used only by ``make_golden.py`` (fixture generation).  /root/reference does not
exist on the GPU box, so nothing under tests/ imports this at test time.

The reference's hot-path files import ``mmcv`` (absent here).  The reference
tree vendors a pure-Python mmcv 1.3.17 under
``controlnet/annotator/uniformer/mmcv`` (which refers to itself as
``annotator.uniformer.mmcv``); we alias ``mmcv[.x]`` to it, stub its two absent
third-party imports (``addict``, ``yapf``) and its compiled ``_ext`` (never
called on CPU: MultiScaleDeformableAttention takes the
``multi_scale_deformable_attn_pytorch`` branch when ``value.is_cuda`` is False).
"""
import importlib
import importlib.abc
import importlib.machinery
import sys
import types

REF = "/root/reference"


class _Dict(dict):
    """Minimal stand-in for addict.Dict (attribute access + recursive wrap)."""

    def __init__(self, *args, **kwargs):
        super().__init__()
        for a in args:
            if isinstance(a, dict):
                for k, v in a.items():
                    self[k] = v
        for k, v in kwargs.items():
            self[k] = v

    @classmethod
    def _hook(cls, v):
        if isinstance(v, dict) and not isinstance(v, cls):
            return cls(v)
        if isinstance(v, (list, tuple)):
            return type(v)(cls._hook(i) for i in v)
        return v

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            return self.__missing__(k)

    def __missing__(self, k):
        raise KeyError(k)

    def __setattr__(self, k, v):
        self[k] = v

    def __setitem__(self, k, v):
        super().__setitem__(k, self._hook(v))

    def to_dict(self):
        return {k: (v.to_dict() if isinstance(v, _Dict) else v) for k, v in self.items()}

    def copy(self):
        return type(self)(self)

    def __deepcopy__(self, memo):
        import copy
        return type(self)({copy.deepcopy(k, memo): copy.deepcopy(v, memo) for k, v in self.items()})

    def update(self, *a, **k):
        for d in a:
            for kk, vv in dict(d).items():
                self[kk] = vv
        for kk, vv in k.items():
            self[kk] = vv


class _Ext(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)

        def _f(*a, **k):
            raise RuntimeError(f"mmcv._ext.{name} is not available (CPU shim)")
        return _f


class _AliasFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, name, path=None, target=None):
        if (name == "mmcv" or name.startswith("mmcv.")) and name != "mmcv._ext":
            return importlib.machinery.ModuleSpec(name, self)
        return None

    def create_module(self, spec):
        return importlib.import_module("annotator.uniformer." + spec.name)

    def exec_module(self, module):
        pass


_installed = False


def install(package="segmentation"):
    """Make ``import mmcv`` / ``import mmseg`` (or ``depth``) resolve to the reference tree."""
    global _installed
    if not _installed:
        addict = types.ModuleType("addict")
        addict.Dict = _Dict
        sys.modules["addict"] = addict
        for n in ("yapf", "yapf.yapflib", "yapf.yapflib.yapf_api"):
            sys.modules[n] = types.ModuleType(n)
        sys.modules["yapf.yapflib.yapf_api"].FormatCode = lambda s, **k: (s, True)
        ext = _Ext("mmcv._ext")
        sys.modules["mmcv._ext"] = ext
        sys.modules["annotator.uniformer.mmcv._ext"] = ext
        for name, path in (("annotator", f"{REF}/controlnet/annotator"),
                           ("annotator.uniformer", f"{REF}/controlnet/annotator/uniformer")):
            m = types.ModuleType(name)
            m.__path__ = [path]
            sys.modules[name] = m
        sys.meta_path.insert(0, _AliasFinder())
        _installed = True
    p = f"{REF}/{package}"
    if p not in sys.path:
        sys.path.insert(0, p)
