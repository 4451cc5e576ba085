"""GPU sanity of the plug-in callers that only stack work for the loop (depth aug_test with batched views, slide_inference
with batched windows): shapes, finiteness, and agreement with the unbatched form up to the sampling noise."""
import os
import sys
import warnings

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_plugin_cpu import _depth_model, HERE          # noqa: E402  (registers ToyBackbone)
from ddp_b200.config import Config                      # noqa: E402
from ddp_b200.registry import build_segmentor           # noqa: E402
from oracle import ddp_oracle as O                      # noqa: E402

torch.manual_seed(0)
dev = torch.device("cuda")
# ---- depth: two views (original + horizontally flipped), as the NYU test pipeline produces them
model = _depth_model(test_cfg=dict(mode="whole"))
model.load_state_dict(O.make_weights(O.OracleConfig(task="depth"), seed=0), strict=False)
model = model.to(dev).eval()
img = torch.randn(1, 3, 96, 128, device=dev)
meta = dict(ori_shape=(96, 128, 3), img_shape=(96, 128, 3), flip=False)
views, metas = [img, img.flip(3)], [[meta], [dict(meta, flip=True, flip_direction="horizontal")]]
a = model.forward_test(views, metas)[0]
model.test_cfg = dict(mode="whole", batch_views=False)
b = model.forward_test(views, metas)[0]
one = model.forward_test([img], [[meta]])[0]
print("depth aug_test", a.shape, float(abs(a - b).max()), float(abs(a - one).max()), float(a.min()), float(a.max()))
assert a.shape == (1, 96, 128) and (a == a).all() and abs(a - b).max() < 0.5
# ---- segmentation: slide inference, 4 windows in one call vs one call each
cfg = Config.fromfile(os.path.join(HERE, "fixtures", "ddp_toy_config.py"))
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    seg = build_segmentor(cfg.model)
seg.load_state_dict(O.make_weights(O.OracleConfig(task="seg", num_classes=19), seed=0), strict=False)
seg = seg.to(dev).eval()
img = torch.randn(1, 3, 96, 160, device=dev)
meta = [dict(ori_shape=(96, 160, 3), img_shape=(96, 160, 3), flip=False)]
seg.test_cfg = dict(mode="slide", crop_size=(64, 96), stride=(32, 64), window_batch=8)
p8 = seg.inference(img, meta, rescale=True)
seg.test_cfg = dict(mode="slide", crop_size=(64, 96), stride=(32, 64), window_batch=1)
p1 = seg.inference(img, meta, rescale=True)
agree = float((p8.argmax(1) == p1.argmax(1)).float().mean())
print("slide_inference", tuple(p8.shape), "argmax agreement batched vs per-window (different noise draws):", agree)
assert p8.shape == (1, 19, 96, 160) and bool(torch.isfinite(p8).all()) and abs(float(p8.sum(1).mean()) - 1.0) < 1e-4
print("CALLERS-SANITY-OK")
