"""Stands in for the reference's model/resnet_deconv.py: same public symbol, same call (resnet_deconv.py:8)."""
from awr_b200.modules import get_deconv_net, ResnetDeconv  # noqa: F401
