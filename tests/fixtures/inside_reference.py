"""Run in a SUBPROCESS by tests/test_plugin_cpu.py (build container only: needs /root/reference).

Scenario A of INTEGRATION.md, end to end on the host side: inside the reference's own segmentation tree (imported through
the golden-fixture shim, tests/golden/refshim.py), `register_into_mmseg()` swaps the DDP classes in, the reference's
`build_segmentor` builds the reference's UNCHANGED config, and a reference-format state dict loads by key."""
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

refshim.install("segmentation")
for n in ("mmcls", "mmcls.models"):
    sys.modules[n] = types.ModuleType(n)
from mmcv import Config  # noqa: E402
from mmseg.models import build_segmentor  # noqa: E402
import ddp_b200.registry as R  # noqa: E402
import ddp_b200.models as M  # noqa: E402
from ddp_b200.neck import FusedNeck  # noqa: E402

assert R.register_into_mmseg() == ["mmseg"]
cfg = Config.fromfile(f"{refshim.REF}/segmentation/configs/cityscapes/ddp_swin_t_4x4_512x1024_160k_cityscapes.py")
cfg.model.pretrained = None
cfg.model.backbone.init_cfg = None
model = build_segmentor(cfg.model)                       # the REFERENCE's builder and registry
assert type(model) is M.DDP, type(model)
assert isinstance(model.neck, FusedNeck)
assert type(model.backbone).__module__.startswith("mmseg.models.backbones"), type(model.backbone)   # the reference's own Swin
assert type(model.decode_head).__module__.startswith("ddp_b200.")

# a reference-format state dict: build the reference's ORIGINAL classes too and move their parameters over by key
import mmseg.models.segmentors.ddp as ref_ddp  # noqa: E402
from mmseg.models.builder import SEGMENTORS, HEADS  # noqa: E402
import mmseg.models.decode_heads.deformable_head_with_time as ref_head  # noqa: E402
SEGMENTORS.register_module(name="DDP", force=True, module=ref_ddp.DDP)
HEADS.register_module(name="DeformableHeadWithTime", force=True, module=ref_head.DeformableHeadWithTime)
cfg.model.auxiliary_head.norm_cfg = dict(type="BN")
cfg.model.decode_head.norm_cfg = dict(type="BN")
ref = build_segmentor(cfg.model)
assert type(ref) is ref_ddp.DDP
sd = ref.state_dict()
missing, unexpected = model.load_state_dict(sd, strict=False)
assert not missing, missing[:5]
assert all(k.startswith("auxiliary_head.") for k in unexpected), [k for k in unexpected if not k.startswith("auxiliary_head.")][:5]
ours = model.state_dict()
assert all(torch.equal(ours[k], sd[k]) for k in ours)

# the backbone (reference code) runs on the CPU; the neck / decode loop have no CPU path and say so
with torch.no_grad():
    feats = model.backbone(torch.randn(1, 3, 64, 64))
assert [tuple(f.shape[1:]) for f in feats] == [(96, 16, 16), (192, 8, 8), (384, 4, 4), (768, 2, 2)]
if not torch.cuda.is_available():
    try:
        model.neck(feats)
        raise SystemExit("the neck ran without a GPU")
    except RuntimeError as e:
        assert "no CPU fallback" in str(e)
print("INSIDE-REFERENCE-OK")
