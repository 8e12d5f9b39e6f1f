"""Bit-reproducibility of the training step in the deterministic build (libawr_b200_det.so, AWR_B200_DETERMINISTIC=1): every sum that
several CTAs share -- BatchNorm statistics, their backward sums, split-K weight gradients, bias gradients -- is accumulated
order-independently, so two runs from the same state give bit-identical losses, gradients and parameters (the reference is
bit-deterministic on CPU; fp32 atomics, as in the default build, are not).  The library is chosen at import, hence the subprocesses."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_RUN = r"""
import sys, hashlib, torch
sys.path.insert(0, %r)
import awr_b200
from awr_b200 import _lib as L
from awr_b200.trainer import FusedTrainer
from oracle import awr_oracle as O
assert L.deterministic() == (%d == 1)
net, ks = %r, %r
kind, n = net.split("_")
sd = O.randomize_bn(O.resnet_deconv_init(int(n), 14, 2, 81, head_std=0.02), 82) if kind == "resnet" else O.randomize_bn(O.hourglass_init(int(n), 14, 81, head_gain=1.0), 82)
batches = [O.synthetic_batch(4, 128, 14, 90 + i) for i in range(3)]
out = []
for run in range(2):
    m = awr_b200.get_deconv_net(int(n), 14, 2, precision="bf16") if kind == "resnet" else awr_b200.PoseNet(net, 14, precision="bf16")
    m.load_state_dict(sd, strict=True)
    tr = FusedTrainer(m.cuda(), 4, 128, ks, 1.0, 1.0, lr=1e-3, use_graph=(run == 1), keep_grads=True)
    losses = []
    for s in range(6):
        img, jt = batches[s %% 3]
        losses.append(tr.train_step(img.cuda(), jt.cuda()))
    torch.cuda.synchronize()
    h = hashlib.sha256(tr.store.params.cpu().numpy().tobytes()).hexdigest()
    hb = hashlib.sha256(b"".join(b.cpu().numpy().tobytes() for b in tr.store.buffers.values())).hexdigest()
    out.append((losses, h, hb))
print("LOSSES_EQUAL", out[0][0] == out[1][0])
print("PARAMS_EQUAL", out[0][1] == out[1][1])
print("BUFFERS_EQUAL", out[0][2] == out[1][2])
print("LOSS0", out[0][0][0], out[0][0][-1])
"""


def _run(det, net, ks):
    env = dict(os.environ, AWR_B200_DETERMINISTIC=str(det))
    env.pop("AWR_B200_LIB", None)
    r = subprocess.run([sys.executable, "-c", _RUN % (ROOT, det, net, ks)], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return dict(line.split(" ", 1) for line in r.stdout.strip().splitlines() if " " in line)


@pytest.mark.parametrize("net,ks", [("resnet_18", 1.0), ("hourglass_1", 0.4)])
def test_deterministic_build_is_bit_reproducible(net, ks):
    """Six bf16 training steps, eager launches vs CUDA-graph replays, from the same state: bit-identical losses, parameters, BN buffers."""
    if not os.path.exists(os.path.join(ROOT, "awr-adaptive-weighting-regression_b200", "libawr_b200_det.so")):
        pytest.fail("libawr_b200_det.so missing: run __graft_entry__.build() (make DET=1)")
    res = _run(1, net, ks)
    assert res["LOSSES_EQUAL"] == "True" and res["PARAMS_EQUAL"] == "True" and res["BUFFERS_EQUAL"] == "True", res
