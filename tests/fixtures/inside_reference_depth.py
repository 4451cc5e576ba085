"""Run in a SUBPROCESS by tests/test_plugin_cpu.py (build container only: needs /root/reference).

INTEGRATION.md scenario A inside the reference's DEPTH tree: `register_into_mmseg()` swaps the depth DDP classes into
`depth.models.builder`, the reference's `build_depther` builds the reference's UNCHANGED NYU config, the reference's own
Swin backbone comes from the reference's registry, and a reference-format state dict loads by key."""
import os
import sys
import types
import warnings

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
warnings.filterwarnings("ignore")
import refshim  # noqa: E402

refshim.install("depth")


class _Any(types.ModuleType):          # timm / matplotlib are absent here and unused on this path
    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        return _Any(self.__name__ + "." + k)

    def __call__(self, *a, **k):
        return None


for n in ("timm", "timm.models", "timm.models.layers", "matplotlib", "matplotlib.pyplot", "matplotlib.cm", "matplotlib.colors"):
    sys.modules[n] = _Any(n)
import mmcv  # noqa: E402,F401
# depth/depth/models/depther/__init__.py imports a module that is not in the tree (regulardepth)
dp = types.ModuleType("depth.models.depther")
dp.__path__ = [f"{refshim.REF}/depth/depth/models/depther"]
sys.modules["depth.models.depther"] = dp
from depth.models import build_depther  # noqa: E402
from mmcv import Config  # noqa: E402
import ddp_b200.registry as R  # noqa: E402
from ddp_b200.models import depth_ddp  # noqa: E402
from ddp_b200.neck import FusedNeck  # noqa: E402

assert "depth" in R.register_into_mmseg()
cfg = Config.fromfile(f"{refshim.REF}/depth/configs/ddp_nyu/ddp_swint_1k_w7_nyu_bs2x8_scale01.py")
cfg.model.backbone.init_cfg = None
model = build_depther(cfg.model)                      # the REFERENCE's builder and registry
assert type(model) is depth_ddp.DDP, type(model)
assert isinstance(model.neck, FusedNeck)
assert type(model.backbone).__module__.startswith("depth.models.backbones"), type(model.backbone)
assert model.max_depth == cfg.model.max_depth and model.timesteps == cfg.model.timesteps
with torch.no_grad():
    feats = model.backbone(torch.randn(1, 3, 64, 96))
assert [tuple(f.shape[1:]) for f in feats] == [(96, 16, 24), (192, 8, 12), (384, 4, 6), (768, 2, 3)], [f.shape for f in feats]
print("INSIDE-REFERENCE-DEPTH-OK")
