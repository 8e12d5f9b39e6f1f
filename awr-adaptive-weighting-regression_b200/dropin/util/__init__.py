"""`util` package shim: feature_tool comes from awr_b200; eval_tool / vis_tool / util keep resolving to the reference
checkout further down sys.path."""
from pkgutil import extend_path

__path__ = extend_path(__path__, __name__)
