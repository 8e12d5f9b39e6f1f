"""The reference's shipped checkpoint (results/hourglass_1.pth: legacy torch<=1.1 pickle, numpy scalars in best_records, Adam state keyed by
id(param) with 220 entries for 250 parameters) against the drop-in module and the trainer's optimizer-state mapping.  CPU only; needs the
reference checkout (47 MB file, not redistributable as a fixture): skipped where /root/reference is absent (the GPU box)."""
import os

import pytest
import torch

CKPT = "/root/reference/results/hourglass_1.pth"
pytestmark = pytest.mark.skipif(not os.path.exists(CKPT), reason="reference checkout with results/hourglass_1.pth not present")


def _load():
    return torch.load(CKPT, map_location="cpu", weights_only=False)          # the reference's own file: trusted


def test_shipped_weights_load_strict_into_dropin():
    import awr_b200
    ck = _load()
    m = awr_b200.PoseNet("hourglass_1", 14)
    missing = m.load_state_dict(ck["model"], strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    sd = m.state_dict()
    assert list(sd.keys()) == list(ck["model"].keys())                      # same order: optimizer state maps by position
    for k, v in ck["model"].items():
        assert sd[k].dtype == v.dtype and tuple(sd[k].shape) == tuple(v.shape), k
        assert torch.equal(sd[k], v), k
    assert m._params_dirty                                                  # a trainer built earlier would refresh its bf16 shadow


def test_shipped_optimizer_state_maps_onto_trained_parameters():
    import awr_b200
    from awr_b200.engine import hourglass_layout
    from awr_b200.trainer import map_optimizer_state
    ck = _load()
    m = awr_b200.PoseNet("hourglass_1", 14)
    names = [k for k, _ in m.named_parameters()]
    assert len(names) == 250
    st = map_optimizer_state(ck["optimizer"], names)
    assert len(st) == 220
    never = sorted(set(names) - set(st))
    assert len(never) == 30 and all(".skip_layer." in n for n in never)        # Residual blocks of equal width never call skip_layer
    lay = hourglass_layout(1, 14)
    for n, s in st.items():
        assert tuple(s["exp_avg"].shape) == lay.specs[n].shape == tuple(s["exp_avg_sq"].shape), n
        assert int(s["step"]) == 231948
    g = ck["optimizer"]["param_groups"][0]
    assert g["lr"] == pytest.approx(1e-4) and g["betas"] == (0.9, 0.999) and g["weight_decay"] == 0


def test_best_records_numpy_scalars():
    ck = _load()
    br = ck["best_records"]
    assert int(br["epoch"]) == 14 and float(br["MPE"]) == pytest.approx(7.700112, rel=1e-6) and float(br["AUC"]) == pytest.approx(0.85048, rel=1e-4)
    # StepLR resumes from the stored epoch (train.py:92) and the loop restarts at epoch + 1 (train.py:105)
    from awr_b200 import optim as AO

    class R:
        lr = 1e-3

        def set_lr(self, v):
            self.lr = v
    r = R()
    sch = AO.StepLR(r, 30, 0.1, last_epoch=int(br["epoch"]))
    sch.step(int(br["epoch"]) + 1)
    assert r.lr == pytest.approx(1e-3)                                        # epoch 15 < step 30: still the initial rate
