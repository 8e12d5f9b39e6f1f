"""`util` package shim: feature_tool and eval_tool come from awr_b200; vis_tool / util keep resolving to the reference
checkout further down sys.path."""
from pkgutil import extend_path

__path__ = extend_path(__path__, __name__)
