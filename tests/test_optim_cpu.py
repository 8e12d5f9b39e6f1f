"""Host logic of the trainer's learning-rate schedules (awr_b200.optim) against torch.optim.lr_scheduler on the call patterns of the
reference loop (train.py:89-96,157-160).  CPU only: the classes drive any object with `.lr` / `.set_lr()`."""
import random

import torch

from awr_b200 import optim as AO


class _Rate:
    def __init__(self, lr):
        self.lr = lr

    def set_lr(self, lr):
        self.lr = float(lr)


def _torch_opt(lr):
    return torch.optim.SGD([torch.nn.Parameter(torch.zeros(1))], lr=lr)


def test_steplr_epoch_form_matches_torch():
    """train.py:92,160: StepLR(step_size, gamma 0.1, last_epoch = resumed epoch), then scheduler.step(epoch) once per epoch."""
    import warnings
    for step_size, last in ((30, -1), (10, -1), (5, 14), (10, 9)):
        opt = _torch_opt(1e-3)
        if last >= 0:
            opt.param_groups[0]["initial_lr"] = 1e-3            # what a resumed optimizer state carries (results/hourglass_1.pth does)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ref = torch.optim.lr_scheduler.StepLR(opt, step_size=step_size, gamma=0.1, last_epoch=last)
            r = _Rate(1e-3)
            mine = AO.StepLR(r, step_size, 0.1, last_epoch=last)
            for g in opt.param_groups:                           # train.py:94-96 forces the configured rate after building the scheduler
                g["lr"] = 1e-3
            r.set_lr(1e-3)
            for epoch in range(max(last, 0), max(last, 0) + 45):
                ref.step(epoch)
                mine.step(epoch)
                assert abs(opt.param_groups[0]["lr"] - r.lr) <= 1e-12 * max(1.0, r.lr), (step_size, last, epoch)
                assert mine.last_epoch == ref.last_epoch


def test_steplr_chainable_form_matches_torch():
    for step_size in (1, 3, 7):
        opt = _torch_opt(0.05)
        ref = torch.optim.lr_scheduler.StepLR(opt, step_size=step_size, gamma=0.5)
        r = _Rate(0.05)
        mine = AO.StepLR(r, step_size, 0.5)
        for _ in range(25):
            opt.step(); ref.step(); mine.step()
            assert abs(opt.param_groups[0]["lr"] - r.lr) <= 1e-15 + 1e-12 * r.lr
        sd = mine.state_dict()
        r2 = _Rate(r.lr)
        m2 = AO.StepLR(r2, 99, 0.9)
        m2.load_state_dict(sd)
        assert (m2.step_size, m2.gamma, m2.last_epoch) == (mine.step_size, mine.gamma, mine.last_epoch)


def test_reduce_on_plateau_matches_torch():
    """train.py:90,158: ReduceLROnPlateau(optimizer, 'min', patience=2, min_lr=1e-8) stepped with the epoch's training MPE."""
    rng = random.Random(0)
    for trial in range(20):
        opt = _torch_opt(1e-3)
        ref = torch.optim.lr_scheduler.ReduceLROnPlateau(opt, "min", patience=2, min_lr=1e-8)
        r = _Rate(1e-3)
        mine = AO.ReduceLROnPlateau(r, "min", patience=2, min_lr=1e-8)
        v = 20.0
        for epoch in range(60):
            v = v * (1.0 - 0.05 * rng.random()) if rng.random() < 0.5 else v * (1.0 + 0.02 * rng.random())
            ref.step(v); mine.step(v)
            assert abs(opt.param_groups[0]["lr"] - r.lr) <= 1e-15 + 1e-12 * r.lr, (trial, epoch)
        assert r.lr >= 1e-8


def test_trainer_lr_group_view():
    """FusedTrainer.param_groups[0]['lr'] (train.py:155 prints it) reads and writes through set_lr; checked on a stand-in object."""
    from awr_b200.trainer import _LRGroup
    r = _Rate(1e-3)
    g = _LRGroup(r)
    assert g["lr"] == 1e-3
    g["lr"] = 5e-4
    assert r.lr == 5e-4 and g["lr"] == 5e-4
