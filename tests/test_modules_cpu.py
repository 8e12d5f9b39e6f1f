"""Host-side contract of the nn.Module drop-ins (no GPU): names, shapes, dtypes and ORDER of state_dict() and
parameters() equal the reference's (golden schemas recorded from the reference classes), checkpoints round-trip."""
import io
import os

import pytest
import torch

import awr_b200
from oracle import awr_oracle as O

SCHEMAS = torch.load(os.path.join(os.path.dirname(__file__), "golden", "schemas.pt"))
CTORS = {"resnet_18_ds2": lambda: awr_b200.get_deconv_net(18, 14, 2), "resnet_50_ds2": lambda: awr_b200.get_deconv_net(50, 14, 2),
         "resnet_18_ds4": lambda: awr_b200.get_deconv_net(18, 14, 4), "resnet_18_ds1_J21": lambda: awr_b200.get_deconv_net(18, 21, 1),
         "hourglass_1": lambda: awr_b200.PoseNet("hourglass_1", 14), "hourglass_2": lambda: awr_b200.PoseNet("hourglass_2", 14)}


@pytest.mark.parametrize("name", list(CTORS))
def test_state_dict_schema_matches_reference(name):
    m = CTORS[name]()
    got = [(k, tuple(v.shape), str(v.dtype)) for k, v in m.state_dict().items()]
    assert got == SCHEMAS[name]["state"]
    assert [k for k, _ in m.named_parameters()] == SCHEMAS[name]["params"]


def test_checkpoint_roundtrip_and_oracle_weights_load_strict():
    m = awr_b200.get_deconv_net(18, 14, 2)
    sd = O.randomize_bn(O.resnet_deconv_init(18, 14, 2, 3, head_std=0.02), 4)
    m.load_state_dict(sd, strict=True)
    buf = io.BytesIO()
    torch.save({"model": m.state_dict(), "best_records": {"epoch": 1}}, buf)      # layout of train.py:165-172
    buf.seek(0)
    back = torch.load(buf)["model"]
    for k, v in sd.items():
        assert torch.equal(back[k], v), k
    h = awr_b200.PoseNet("hourglass_1", 14)
    h.load_state_dict(O.hourglass_init(1, 14, 5), strict=True)


def test_init_distributions_follow_reference():
    torch.manual_seed(0)
    m = awr_b200.get_deconv_net(18, 14, 2)
    sd = m.state_dict()
    w = sd["layer1.0.conv1.weight"]
    assert abs(w.std().item() - (2.0 / (9 * 64)) ** 0.5) < 2e-3          # resnet_deconv.py:96-97
    assert abs(sd["deconv_layers.0.weight"].std().item() - 1e-3) < 5e-5  # :104
    assert abs(sd["final1.weight"].std().item() - 1e-3) < 1e-4 and sd["final1.bias"].abs().max() == 0   # :108-115
    assert torch.all(sd["pre.1.weight"] == 1) and torch.all(sd["pre.1.bias"] == 0)
    assert torch.all(sd["pre.1.running_var"] == 1) and sd["pre.1.num_batches_tracked"].dtype == torch.int64


def test_cpu_forward_refuses():
    m = awr_b200.get_deconv_net(18, 14, 2)
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 1, 128, 128))
