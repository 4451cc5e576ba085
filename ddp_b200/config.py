"""Loader for the reference's mmcv-style python config files (``_base_`` inheritance, ``_delete_`` keys,
attribute access), so that ``segmentation/configs/{ade,cityscapes}/ddp_*.py`` and
``depth/configs/ddp_*/*.py`` load unchanged without mmcv (reference: mmcv.utils.config.Config), and for the BEV
tree's yaml configs (``bev/configs/nuscenes/seg/ddp-*.yaml``: torchpack's ``configs.load(path, recursive=True)`` —
every ``default.yaml`` from the configs root down to the file's directory, then the file, ``${expr}`` interpolation)."""
import copy
import os
import re

BASE_KEY = "_base_"
DELETE_KEY = "_delete_"


class ConfigDict(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(f"'ConfigDict' object has no attribute '{k}'")

    def __setattr__(self, k, v):
        self[k] = _wrap(v)

    def __setitem__(self, k, v):
        super().__setitem__(k, _wrap(v))

    def __deepcopy__(self, memo):
        return ConfigDict({copy.deepcopy(k, memo): copy.deepcopy(v, memo) for k, v in self.items()})


def _wrap(v):
    if isinstance(v, dict) and not isinstance(v, ConfigDict):
        out = ConfigDict()
        for k, x in v.items():
            out[k] = x
        return out
    if isinstance(v, (list, tuple)):
        return type(v)(_wrap(i) for i in v)
    return v


def _merge(a, b):
    """b into a copy of a (mmcv Config._merge_a_into_b semantics incl. _delete_)."""
    out = copy.deepcopy(a)
    for k, v in b.items():
        if isinstance(v, dict) and k in out and isinstance(out[k], dict) and not v.get(DELETE_KEY, False):
            out[k] = _merge(out[k], v)
        else:
            if isinstance(v, dict):
                v = {kk: vv for kk, vv in v.items() if kk != DELETE_KEY}
            out[k] = copy.deepcopy(v)
    return out


def _load_file(path):
    path = os.path.abspath(path)
    scope = {"__file__": path}
    with open(path) as f:
        exec(compile(f.read(), path, "exec"), scope)
    cfg = {k: v for k, v in scope.items() if not k.startswith("__") and not callable(v) and type(v).__name__ != "module"}
    bases = cfg.pop(BASE_KEY, [])
    if isinstance(bases, str):
        bases = [bases]
    merged = {}
    for b in bases:
        base_cfg = _load_file(os.path.join(os.path.dirname(path), b))
        dup = set(merged) & set(base_cfg)
        if dup:
            raise KeyError(f"Duplicate key is not allowed among bases: {dup}")
        merged.update(base_cfg)
    return _merge(merged, cfg)


class Config(ConfigDict):
    @staticmethod
    def fromfile(path):
        cfg = Config()
        for k, v in _load_file(path).items():
            cfg[k] = v
        cfg.__dict__["filename"] = path
        return cfg


# ---- torchpack-style yaml configs of the BEV tree -------------------------------------------------------------------
def _merge_yaml(a, b):
    out = dict(a)
    for k, v in b.items():
        out[k] = _merge_yaml(out[k], v) if isinstance(v, dict) and isinstance(out.get(k), dict) else copy.deepcopy(v)
    return out


_INTERP = re.compile(r"^\$\{(.*)\}$", re.S)


def _resolve(node, root):
    """Evaluate ``${expr}`` strings against the whole config (attribute access, e.g. ``${augment2d.resize[0]}``)."""
    if isinstance(node, dict):
        return {k: _resolve(v, root) for k, v in node.items()}
    if isinstance(node, list):
        return [_resolve(v, root) for v in node]
    if isinstance(node, str):
        m = _INTERP.match(node.strip())
        if m:
            return eval(m.group(1), {"__builtins__": {}}, _wrap(root))      # noqa: S307 - config expressions, as torchpack does
    return node


def load_yaml(path, recursive=True):
    """The reference BEV tree's ``configs.load(path, recursive=True)``: defaults of every directory level, then the file."""
    import yaml
    path = os.path.abspath(path)
    chain = [path]
    if recursive:
        d = os.path.dirname(path)
        while True:
            f = os.path.join(d, "default.yaml")
            if os.path.exists(f) and f != path:
                chain.append(f)
            if os.path.basename(d) == "configs" or os.path.dirname(d) == d:
                break
            d = os.path.dirname(d)
    cfg = {}
    for f in reversed(chain):
        with open(f) as fh:
            cfg = _merge_yaml(cfg, yaml.safe_load(fh) or {})
    for _ in range(8):                     # references to values that are themselves interpolated
        new = _resolve(cfg, cfg)
        if new == cfg:
            break
        cfg = new
    out = Config()
    for k, v in cfg.items():
        out[k] = v
    out.__dict__["filename"] = path
    return out
