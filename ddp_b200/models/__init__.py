from .deformable_head_with_time import DeformableHeadWithTime, DepthDeformableHeadWithTime  # noqa: F401
from .ddp import DDP, SelfAlignedDDP  # noqa: F401
from .depth_ddp import DDP as DepthDDP  # noqa: F401
from .. import neck as _neck  # noqa: F401,E402  (registers FPN / MultiStageMerging)
