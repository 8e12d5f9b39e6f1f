"""World-size-2 gloo test of the data-parallel host logic (awr_b200/dp.py): batch sharding, sum all-reduce of a flat gradient
buffer, 1/world folded into Adam == single-process step on the concatenated batch; parameter broadcast from rank 0."""
import os
import socket
import sys

import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _model_pred(w, feat):
    """Toy differentiable 'backbone' without BatchNorm (BN statistics are per replica by design): per-channel gain on fixed features."""
    return feat * w.view(1, -1, 1, 1)


def _grads(w, feat, img, jt, ks):
    from oracle import awr_oracle as O
    w = w.clone().requires_grad_(True)
    pred = _model_pred(w, feat)
    gt = O.joint2offset(jt, img, ks, feat.shape[-1])
    loss = O.smooth_l1(O.offset2joint_softmax(pred, img, ks), jt) + O.smooth_l1(pred, gt)
    loss.backward()
    return w.grad.detach()


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.set_num_threads(2)
    from awr_b200 import dp
    from oracle import awr_oracle as O
    r, l, w_ = dp.init_from_env(backend="gloo")
    assert (r, w_) == (rank, world)
    B, J, Fs, H, ks = 2, 14, 32, 128, 1.0
    img, jt = O.synthetic_batch(B * world, H, J, 5)
    feat = torch.randn(B * world, 4 * J, Fs, Fs, generator=torch.Generator().manual_seed(6))
    # rank-dependent initial parameters -> broadcast must make them equal to rank 0's
    w = torch.full((4 * J,), 1.0 + rank)
    dp.broadcast_([w], src=0)
    assert torch.all(w == 1.0)
    sl = dp.shard_slice(rank, B)
    g = _grads(w, feat[sl], img[sl], jt[sl], ks)
    dp.allreduce_sum_(g)
    m, v = torch.zeros_like(w), torch.zeros_like(w)
    O.adam_step(w, g / world, m, v, 1)                       # grad_scale = 1/world, as awr_adam_flat applies it
    t = dp.max_over_ranks(rank + 1.0, "cpu")
    ms, bw = dp.allreduce_busbw(torch.ones(1 << 16), reps=3)          # bench.py's all-reduce leg (wall clock under gloo)
    if rank == 0:
        torch.save({"w": w, "g": g / world, "tmax": t, "ar_ms": ms, "ar_bw": bw}, out)
    torch.distributed.destroy_process_group()


def test_dp_world2_equals_single_process_on_concatenated_batch(tmp_path):
    sys.path.insert(0, ROOT)
    from oracle import awr_oracle as O
    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = torch.load(out)
    B, J, Fs, H, ks = 2, 14, 32, 128, 1.0
    img, jt = O.synthetic_batch(4, H, J, 5)
    feat = torch.randn(4, 4 * J, Fs, Fs, generator=torch.Generator().manual_seed(6))
    w = torch.ones(4 * J)
    g = _grads(w, feat, img, jt, ks)
    assert torch.allclose(got["g"], g, rtol=1e-4, atol=1e-9)
    m, v = torch.zeros_like(w), torch.zeros_like(w)
    O.adam_step(w, g, m, v, 1)
    assert torch.allclose(got["w"], w, rtol=1e-5, atol=1e-7)
    assert got["tmax"] == 2.0
    assert got["ar_ms"] > 0 and got["ar_bw"] > 0
