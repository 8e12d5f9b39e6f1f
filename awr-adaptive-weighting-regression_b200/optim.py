"""Learning-rate schedules of the reference's training loop, driving FusedTrainer's device-side learning rate.

train.py:89-92 builds `ReduceLROnPlateau(optimizer, "min", patience=2, min_lr=1e-8)` ('auto') or
`StepLR(optimizer, step_size=config.step, gamma=0.1, last_epoch=best_records['epoch'])` ('step'), train.py:94-96 then forces
the rate back to config.lr, and train.py:157-160 calls `scheduler.step(train_mpe)` / `scheduler.step(epoch)` once per epoch.
These classes reproduce torch.optim.lr_scheduler's arithmetic for exactly those call patterns (tests/test_optim_cpu.py checks
them against the torch classes) on anything that has `.lr` and `.set_lr(lr)` -- the rate is a device scalar the captured
optimizer kernel reads, so a change costs one 4-byte fill and no CUDA-graph re-capture.
"""
import math


class StepLR:
    """torch.optim.lr_scheduler.StepLR(optimizer, step_size, gamma, last_epoch).

    step(epoch) -- the form train.py:160 uses -- sets the closed form initial_lr * gamma ** (epoch // step_size);
    step() advances one epoch with torch's chainable rule (multiply by gamma whenever the new epoch is a multiple of step_size).
    `initial_lr` is the rate the trainer holds at construction (torch records it as param_groups[0]['initial_lr'])."""

    def __init__(self, trainer, step_size, gamma=0.1, last_epoch=-1, initial_lr=None):
        if step_size <= 0:
            raise ValueError("step_size must be positive")
        self.trainer, self.step_size, self.gamma = trainer, int(step_size), float(gamma)
        self.initial_lr = float(trainer.lr if initial_lr is None else initial_lr)
        self.last_epoch = int(last_epoch)
        self.step()                      # torch performs one initial step at construction (last_epoch -1 -> 0 leaves the rate alone)

    def get_last_lr(self):
        return [self.trainer.lr]

    def step(self, epoch=None):
        if epoch is None:
            self.last_epoch += 1
            if self.last_epoch != 0 and self.last_epoch % self.step_size == 0:
                self.trainer.set_lr(self.trainer.lr * self.gamma)
        else:
            self.last_epoch = int(math.floor(epoch))
            self.trainer.set_lr(self.initial_lr * self.gamma ** (self.last_epoch // self.step_size))

    def state_dict(self):
        return {"step_size": self.step_size, "gamma": self.gamma, "last_epoch": self.last_epoch, "initial_lr": self.initial_lr}

    def load_state_dict(self, sd):
        self.step_size, self.gamma = int(sd["step_size"]), float(sd["gamma"])
        self.last_epoch, self.initial_lr = int(sd["last_epoch"]), float(sd["initial_lr"])


class ReduceLROnPlateau:
    """torch.optim.lr_scheduler.ReduceLROnPlateau (train.py:90: mode 'min', patience 2, min_lr 1e-8; torch defaults otherwise:
    factor 0.1, threshold 1e-4 relative, cooldown 0, eps 1e-8).  step(metric) once per epoch with the training MPE (train.py:158)."""

    def __init__(self, trainer, mode="min", factor=0.1, patience=10, threshold=1e-4, threshold_mode="rel", cooldown=0, min_lr=0.0, eps=1e-8):
        if factor >= 1.0:
            raise ValueError("Factor should be < 1.0.")
        if mode not in ("min", "max") or threshold_mode not in ("rel", "abs"):
            raise ValueError("unknown mode")
        self.trainer, self.mode, self.factor, self.patience = trainer, mode, float(factor), int(patience)
        self.threshold, self.threshold_mode, self.cooldown, self.min_lr, self.eps = float(threshold), threshold_mode, int(cooldown), float(min_lr), float(eps)
        self.best = math.inf if mode == "min" else -math.inf
        self.num_bad_epochs, self.cooldown_counter, self.last_epoch = 0, 0, 0

    def _is_better(self, a, best):
        if self.mode == "min":
            return a < best * (1.0 - self.threshold) if self.threshold_mode == "rel" else a < best - self.threshold
        return a > best * (self.threshold + 1.0) if self.threshold_mode == "rel" else a > best + self.threshold

    def get_last_lr(self):
        return [self.trainer.lr]

    def step(self, metrics):
        current = float(metrics)
        self.last_epoch += 1
        if self._is_better(current, self.best):
            self.best, self.num_bad_epochs = current, 0
        else:
            self.num_bad_epochs += 1
        if self.cooldown_counter > 0:
            self.cooldown_counter -= 1
            self.num_bad_epochs = 0
        if self.num_bad_epochs > self.patience:
            old = self.trainer.lr
            new = max(old * self.factor, self.min_lr)
            if old - new > self.eps:
                self.trainer.set_lr(new)
            self.cooldown_counter = self.cooldown
            self.num_bad_epochs = 0

    def state_dict(self):
        return {k: v for k, v in self.__dict__.items() if k != "trainer"}

    def load_state_dict(self, sd):
        self.__dict__.update(sd)
