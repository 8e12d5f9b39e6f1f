"""Stands in for the reference's model/hourglass.py: PoseNet(net, joint_num, ...) (hourglass.py:105-165)."""
from awr_b200.modules import PoseNet  # noqa: F401
