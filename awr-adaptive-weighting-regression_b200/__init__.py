"""awr_b200 -- B200-native hot path of AWR (Adaptive Weighting Regression).

Public surface mirrors the reference's four imported symbols (train.py:13-17):
    get_deconv_net, PoseNet, My_SmoothL1Loss, FeatureModule
plus the fused trainer used by bench.py and the device-side EvalUtil (util/eval_tool.py).
"""
from . import _lib  # noqa: F401
from .eval_tool import EvalUtil  # noqa: F401
from .feature_tool import FeatureModule  # noqa: F401
from .loss import My_SmoothL1Loss  # noqa: F401
from .modules import get_deconv_net, PoseNet  # noqa: F401
