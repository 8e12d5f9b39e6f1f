"""Stands in for the reference's util/eval_tool.py (eval_tool.py:5-135): same class, arithmetic on the device."""
from awr_b200.eval_tool import EvalUtil  # noqa: F401
