"""The import shim makes the reference's own import lines (train.py:13-17) resolve to awr_b200 while the rest of the
reference's `util` package stays reachable.  Needs the reference checkout only for the last assertion (skipped without it)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "awr-adaptive-weighting-regression_b200", "dropin")


def test_reference_import_lines_resolve_to_awr_b200():
    ref = "/root/reference"
    code = (
        "from model.hourglass import PoseNet\n"
        "from model.resnet_deconv import get_deconv_net\n"
        "from model.loss import My_SmoothL1Loss\n"
        "from util.feature_tool import FeatureModule\n"
        "from util.eval_tool import EvalUtil\n"                                  # train.py:18 / test.py:16
        "import awr_b200\n"
        "assert EvalUtil is awr_b200.EvalUtil\n"
        "assert get_deconv_net is awr_b200.get_deconv_net and PoseNet is awr_b200.PoseNet\n"
        "assert My_SmoothL1Loss is awr_b200.My_SmoothL1Loss and FeatureModule is awr_b200.FeatureModule\n"
        "m = get_deconv_net(18, 14, 2); assert len(m.state_dict()) == 142\n"
    )
    if os.path.isdir(ref):
        code += "from util.util import xyz2uvd, uvd2xyz\nimport util.util as u; assert u.__file__.startswith('/root/reference')\n"
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([SHIM, ROOT] + ([ref] if os.path.isdir(ref) else [])))
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
