"""Stands in for the reference's util/feature_tool.py (feature_tool.py:10-65)."""
from awr_b200.feature_tool import FeatureModule  # noqa: F401
