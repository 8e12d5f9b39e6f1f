"""CPU oracle for the AWR dense hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT.

Only tests/ (+ tests/golden/make_*.py), __graft_entry__.smoke(), bench.py's cpu_baseline /
--impl reference / parity / gpu_eager_baseline / preprocess legs and the developer scripts under tools/
import this file, always as the checker, the timed CPU baseline or a seeded input generator.
The product path (awr_b200) never does -- tests/test_abi.py greps the package for it -- and
fails loudly when the CUDA library is missing.

What this is: a plain restatement, in functional torch-CPU fp32 arithmetic, of
the algorithm the reference runs on the path named by BASELINE.json:

  backbone (model/resnet_deconv.py, model/hourglass.py)
    -> adaptive-weighting head (util/feature_tool.py)
    -> SmoothL1 (model/loss.py) -> backward -> Adam (train.py:67,129-131)

The reference itself is pure Python on top of torch (pinned torch==1.1.0 in
requirements.txt:1; README.md:11 says 1.4.0).  The heavy arithmetic therefore
lives in a third-party dependency (torch conv2d / conv_transpose2d /
batch_norm / max_pool2d / softmax) whose semantics are stable 1.1 -> 2.11; the
restatement calls the same published primitives through torch.nn.functional on
a flat state_dict, and writes the head / loss / coordinate grid in closed form.

Also restated (numpy, same dtypes as the reference's code): util/eval_tool.py::EvalUtil and the
non-augmented depth preprocessing of dataloader/loader.py.

Parity pin: tests/golden/* were produced by tests/golden/make_golden.py, make_eval_golden.py and
make_preprocess_golden.py, which import the UNMODIFIED reference modules from /root/reference
(with the real cv2) in the build container and record their outputs on seeded inputs;
tests/test_oracle.py, test_eval.py and test_preprocess.py check every function here against
those vectors (the numpy restatements bit-exactly).  (The reference ships no unit
tests or known-answer vectors of its own: SURVEY.md section 4.)
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
BN_EPS = 1e-5
BN_MOMENTUM = 0.1          # model/resnet_deconv.py:6 ; hourglass uses the torch default (also 0.1)
HUBER_DELTA = 0.01         # model/loss.py:12-13
SOFTMAX_SCALE = 30.0       # util/feature_tool.py:60
DEPTH_BG = 0.99            # util/feature_tool.py:35,57

RESNET_SPEC = {18: ("basic", [2, 2, 2, 2]), 50: ("bottleneck", [3, 4, 6, 3]),
               101: ("bottleneck", [3, 4, 23, 3]), 152: ("bottleneck", [3, 8, 36, 3])}


# --------------------------------------------------------------------------------------
# a.4  coordinate grid   (util/feature_tool.py:23-27 and :50-55)
# --------------------------------------------------------------------------------------
def coord_axis(feature_size: int) -> Tensor:
    """1-D pixel-centre coordinate 2*(i+0.5)/F - 1, evaluated in fp32 in the reference's order."""
    i = torch.arange(feature_size).float()
    return 2.0 * (i + 0.5) / feature_size - 1.0


def sample_depth(img: Tensor, feature_size: int) -> Tensor:
    """F.interpolate(img, [F,F]) with the default 'nearest' mode (feature_tool.py:20,44).

    For integer ratios nearest picks src index floor(dst*H/F) = dst*(H/F), i.e. a strided slice."""
    B, C, H, W = img.shape
    assert C == 1 and H % feature_size == 0 and W % feature_size == 0
    return img[:, 0, :: H // feature_size, :: W // feature_size].contiguous()      # (B,F,F)


# --------------------------------------------------------------------------------------
# a.5  joint2offset   (util/feature_tool.py:12-39)
# --------------------------------------------------------------------------------------
def joint2offset(jt_uvd: Tensor, img: Tensor, kernel_size: float, feature_size: int) -> Tensor:
    B, J, _ = jt_uvd.shape
    Fs = feature_size
    d = sample_depth(img, Fs)                                                # (B,F,F)
    ax = coord_axis(Fs).to(jt_uvd.device)
    u = ax.view(1, 1, 1, Fs).expand(B, 1, Fs, Fs)
    v = ax.view(1, 1, Fs, 1).expand(B, 1, Fs, Fs)
    coord = torch.stack((u, v, d.unsqueeze(1)), dim=2)                        # (B,1,3,F,F)
    off = jt_uvd.view(B, J, 3, 1, 1) - coord                                  # :29
    dis = torch.sqrt((off * off).sum(dim=2) + 1e-8)                           # :31
    off_n = off / dis.unsqueeze(2)                                            # :33
    hm = (kernel_size - dis) / kernel_size                                    # :34
    mask = (hm >= 0).float() * (d < DEPTH_BG).float().unsqueeze(1)            # :35
    out_vec = (off_n * mask.unsqueeze(2)).reshape(B, 3 * J, Fs, Fs)           # :37
    out_hm = hm * mask                                                        # :38
    return torch.cat((out_vec, out_hm), dim=1).float()


# --------------------------------------------------------------------------------------
# a.6  offset2joint_softmax and its analytic backward   (util/feature_tool.py:41-65)
# --------------------------------------------------------------------------------------
def offset2joint_softmax(offset: Tensor, img: Tensor, kernel_size: float) -> Tensor:
    B, C, Fs, _ = offset.shape
    J = C // 4
    P = Fs * Fs
    d = sample_depth(img, Fs).view(B, 1, P)
    m = (d < DEPTH_BG).float()                                               # :57
    ax = coord_axis(Fs).to(offset.device)
    u = ax.view(1, Fs).expand(Fs, Fs).reshape(1, 1, P)
    v = ax.view(Fs, 1).expand(Fs, Fs).reshape(1, 1, P)
    vec = offset[:, : 3 * J].reshape(B, J, 3, P) * m.unsqueeze(1)             # :58
    h = offset[:, 3 * J:].reshape(B, J, P) * m                                # :59
    w = torch.softmax(h * SOFTMAX_SCALE, dim=-1)                              # :60 (masked pixels keep logit 0)
    dis = kernel_size - h * kernel_size                                      # :61
    coord = torch.stack((u.expand(B, 1, P), v.expand(B, 1, P), d), dim=2)     # (B,1,3,P)
    val = vec * dis.unsqueeze(2) + coord
    return (val * w.unsqueeze(2)).sum(dim=-1).float()                         # :63


def offset2joint_softmax_bwd(offset: Tensor, img: Tensor, kernel_size: float, g_uvd: Tensor) -> Tensor:
    """Closed-form d L / d offset given g_uvd = d L / d uvd  (derivation: SURVEY.md section 8 a.6)."""
    B, C, Fs, _ = offset.shape
    J = C // 4
    P = Fs * Fs
    d = sample_depth(img, Fs).view(B, 1, P)
    m = (d < DEPTH_BG).float()
    ax = coord_axis(Fs)
    u = ax.view(1, Fs).expand(Fs, Fs).reshape(1, 1, P)
    v = ax.view(Fs, 1).expand(Fs, Fs).reshape(1, 1, P)
    vec = offset[:, : 3 * J].reshape(B, J, 3, P) * m.unsqueeze(1)
    h = offset[:, 3 * J:].reshape(B, J, P) * m
    w = torch.softmax(h * SOFTMAX_SCALE, dim=-1)
    dis = kernel_size - h * kernel_size
    coord = torch.stack((u.expand(B, 1, P), v.expand(B, 1, P), d), dim=2)
    val = vec * dis.unsqueeze(2) + coord                                      # (B,J,3,P)
    uvd = (val * w.unsqueeze(2)).sum(dim=-1)                                  # (B,J,3)
    g = g_uvd.view(B, J, 3, 1)
    d_vec = g * (w * dis).unsqueeze(2) * m.unsqueeze(1)
    d_h = m * (g * (-kernel_size * w.unsqueeze(2) * vec
                    + SOFTMAX_SCALE * w.unsqueeze(2) * (val - uvd.unsqueeze(-1)))).sum(dim=2)
    return torch.cat((d_vec.reshape(B, 3 * J, Fs, Fs), d_h.reshape(B, J, Fs, Fs)), dim=1)


# --------------------------------------------------------------------------------------
# a.7  My_SmoothL1Loss   (model/loss.py:8-25)  ==  Huber(delta=0.01), mean-reduced
# --------------------------------------------------------------------------------------
def smooth_l1(x: Tensor, y: Tensor) -> Tensor:
    assert x.shape == y.shape                                                # :10
    z = (x - y).float()
    a = z.abs()
    quad = 0.5 * z * z                                                        # :21-22
    lin = HUBER_DELTA * (a - 0.5 * HUBER_DELTA)                               # :24-25
    return torch.where(a < HUBER_DELTA, quad, lin).mean()


def smooth_l1_grad(x: Tensor, y: Tensor) -> Tensor:
    z = (x - y).float()
    g = torch.where(z.abs() < HUBER_DELTA, z, HUBER_DELTA * torch.sign(z))
    return g / z.numel()


# --------------------------------------------------------------------------------------
# a.1/a.2  ResNet-deconv backbone   (model/resnet_deconv.py:19-215)
# --------------------------------------------------------------------------------------
def _bn(x: Tensor, sd: Dict[str, Tensor], prefix: str, training: bool, new_stats: Optional[dict]) -> Tensor:
    w, b = sd[prefix + ".weight"], sd[prefix + ".bias"]
    rm, rv = sd[prefix + ".running_mean"], sd[prefix + ".running_var"]
    if training:
        rm2, rv2 = rm.clone(), rv.clone()
        y = F.batch_norm(x, rm2, rv2, w, b, True, BN_MOMENTUM, BN_EPS)
        if new_stats is not None:
            new_stats[prefix + ".running_mean"] = rm2
            new_stats[prefix + ".running_var"] = rv2
            new_stats[prefix + ".num_batches_tracked"] = sd[prefix + ".num_batches_tracked"] + 1
        return y
    return F.batch_norm(x, rm, rv, w, b, False, BN_MOMENTUM, BN_EPS)


def resnet_layout(layers: int) -> Tuple[str, List[int]]:
    return RESNET_SPEC[layers]


def resnet_deconv_forward(sd: Dict[str, Tensor], x: Tensor, layers: int, downsample: int,
                          training: bool = False, new_stats: Optional[dict] = None) -> Tensor:
    """x (B,1,H,W) -> (B,4J,H/ds,W/ds).  Follows ResnetDeconv.forward (resnet_deconv.py:118-136)."""
    kind, counts = RESNET_SPEC[layers]
    bn = lambda t, p: _bn(t, sd, p, training, new_stats)
    c = F.conv2d(x, sd["pre.0.weight"], None, 1, 2)                           # :32  5x5 s1 p2, no bias
    c = F.max_pool2d(F.relu(bn(c, "pre.1")), 3, 2, 1)                         # :33-35
    for li, nblk in enumerate(counts, start=1):
        for bi in range(nblk):
            p = f"layer{li}.{bi}"
            stride = 2 if (li > 1 and bi == 0) else 1                         # :39-43,66
            res = c
            if kind == "basic":                                              # BasicBlock.forward :158-174
                o = F.relu(bn(F.conv2d(c, sd[p + ".conv1.weight"], None, stride, 1), p + ".bn1"))
                o = bn(F.conv2d(o, sd[p + ".conv2.weight"], None, 1, 1), p + ".bn2")
            else:                                                            # Bottleneck.forward :195-215
                o = F.relu(bn(F.conv2d(c, sd[p + ".conv1.weight"]), p + ".bn1"))
                o = F.relu(bn(F.conv2d(o, sd[p + ".conv2.weight"], None, stride, 1), p + ".bn2"))
                o = bn(F.conv2d(o, sd[p + ".conv3.weight"]), p + ".bn3")
            if (p + ".downsample.0.weight") in sd:                            # :58-64
                res = bn(F.conv2d(c, sd[p + ".downsample.0.weight"], None, stride, 0), p + ".downsample.1")
            c = F.relu(o + res)
    n_deconv = 4 - int(math.log(downsample, 2))                               # :45
    for i in range(n_deconv):                                                 # :73-91  k4 s2 p1, no bias
        c = F.conv_transpose2d(c, sd[f"deconv_layers.{3 * i}.weight"], None, 2, 1)
        c = F.relu(bn(c, f"deconv_layers.{3 * i + 1}"))
    vec = F.conv2d(c, sd["final1.weight"], sd["final1.bias"])                 # :133
    ht = F.conv2d(c, sd["final2.weight"], sd["final2.bias"])                  # :134
    return torch.cat([vec, ht], dim=1)                                        # :136


def resnet_deconv_init(layers: int, num_joints: int, downsample: int, seed: int,
                       head_std: Optional[float] = None) -> Dict[str, Tensor]:
    """Seeded state_dict with the reference's key schema / shapes and its init distributions
    (resnet_deconv.py:93-115).  NOT bit-identical to the reference's RNG stream (irrelevant:
    golden vectors are produced by loading THIS dict into the reference module).
    head_std overrides the 1e-3 std of deconv / final layers so heat-maps become peaked
    (a 1e-3 head gives logits ~1e-6 and a vacuous softmax test)."""
    g = torch.Generator().manual_seed(seed)
    kind, counts = RESNET_SPEC[layers]
    exp = 1 if kind == "basic" else 4
    sd: Dict[str, Tensor] = {}

    def conv(name, co, ci, k, std=None):
        std = math.sqrt(2.0 / (k * k * co)) if std is None else std
        sd[name] = torch.randn(co, ci, k, k, generator=g) * std

    def bnp(name, c):
        sd[name + ".weight"] = torch.ones(c)
        sd[name + ".bias"] = torch.zeros(c)
        sd[name + ".running_mean"] = torch.zeros(c)
        sd[name + ".running_var"] = torch.ones(c)
        sd[name + ".num_batches_tracked"] = torch.zeros((), dtype=torch.long)

    conv("pre.0.weight", 64, 1, 5)
    bnp("pre.1", 64)
    inpl = 64
    for li, (planes, nblk) in enumerate(zip([64, 128, 256, 512], counts), start=1):
        for bi in range(nblk):
            p = f"layer{li}.{bi}"
            stride = 2 if (li > 1 and bi == 0) else 1
            if kind == "basic":
                conv(p + ".conv1.weight", planes, inpl, 3); bnp(p + ".bn1", planes)
                conv(p + ".conv2.weight", planes, planes, 3); bnp(p + ".bn2", planes)
            else:
                conv(p + ".conv1.weight", planes, inpl, 1); bnp(p + ".bn1", planes)
                conv(p + ".conv2.weight", planes, planes, 3); bnp(p + ".bn2", planes)
                conv(p + ".conv3.weight", planes * 4, planes, 1); bnp(p + ".bn3", planes * 4)
            if bi == 0 and (stride != 1 or inpl != planes * exp):
                conv(p + ".downsample.0.weight", planes * exp, inpl, 1); bnp(p + ".downsample.1", planes * exp)
            inpl = planes * exp
    hs = 1e-3 if head_std is None else head_std
    for i in range(4 - int(math.log(downsample, 2))):
        sd[f"deconv_layers.{3 * i}.weight"] = torch.randn(inpl, 256, 4, 4, generator=g) * hs   # IOHW
        bnp(f"deconv_layers.{3 * i + 1}", 256)
        inpl = 256
    sd["final1.weight"] = torch.randn(3 * num_joints, 256, 1, 1, generator=g) * hs
    sd["final1.bias"] = torch.zeros(3 * num_joints)
    sd["final2.weight"] = torch.randn(num_joints, 256, 1, 1, generator=g) * hs
    sd["final2.bias"] = torch.zeros(num_joints)
    return sd


def randomize_bn(sd: Dict[str, Tensor], seed: int) -> Dict[str, Tensor]:
    """Give every BN non-trivial affine + running statistics so eval-mode parity is not vacuous."""
    g = torch.Generator().manual_seed(seed)
    out = dict(sd)
    for k in list(sd):
        if k.endswith(".running_mean"):
            p = k[: -len(".running_mean")]
            c = sd[k].numel()
            out[p + ".weight"] = 0.5 + torch.rand(c, generator=g)
            out[p + ".bias"] = 0.2 * torch.randn(c, generator=g)
            out[p + ".running_mean"] = 0.1 * torch.randn(c, generator=g)
            out[p + ".running_var"] = 0.5 + torch.rand(c, generator=g)
    return out


# --------------------------------------------------------------------------------------
# a.3  Hourglass backbone   (model/hourglass.py:6-165)
# --------------------------------------------------------------------------------------
def _hg_conv(sd, p, x, k):                       # Conv without bn/relu: biased conv (:10,:17-20)
    return F.conv2d(x, sd[p + ".conv.weight"], sd[p + ".conv.bias"], 1, (k - 1) // 2)


def _hg_residual(sd, p, x, bn):                  # Residual.forward :44-59
    cin = sd[p + ".bn1.weight"].numel()
    cout = sd[p + ".conv3.conv.weight"].shape[0]
    res = _hg_conv(sd, p + ".skip_layer", x, 1) if cin != cout else x
    o = _hg_conv(sd, p + ".conv1", F.relu(bn(x, p + ".bn1")), 1)
    o = _hg_conv(sd, p + ".conv2", F.relu(bn(o, p + ".bn2")), 3)
    o = _hg_conv(sd, p + ".conv3", F.relu(bn(o, p + ".bn3")), 1)
    return o + res


def _hg_hourglass(sd, p, x, n, bn):              # Hourglass.forward :79-88
    up1 = _hg_residual(sd, p + ".up1", x, bn)
    low1 = _hg_residual(sd, p + ".low1", F.max_pool2d(x, 2, 2), bn)
    low2 = _hg_hourglass(sd, p + ".low2", low1, n - 1, bn) if n > 1 else _hg_residual(sd, p + ".low2", low1, bn)
    low3 = _hg_residual(sd, p + ".low3", low2, bn)
    return up1 + F.interpolate(low3, scale_factor=2, mode="nearest")


def hourglass_forward(sd: Dict[str, Tensor], x: Tensor, nstack: int, training: bool = False,
                      new_stats: Optional[dict] = None) -> List[Tensor]:
    """PoseNet.forward (hourglass.py:144-165): returns the list of per-stack (B,4J,H/2,W/2) volumes."""
    bn = lambda t, p: _bn(t, sd, p, training, new_stats)
    c = F.relu(bn(_hg_conv(sd, "pre.0", x, 5), "pre.0.bn"))                   # :112
    c = _hg_residual(sd, "pre.1", c, bn)
    c = F.max_pool2d(c, 2, 2)
    c = _hg_residual(sd, "pre.3", c, bn)
    c = _hg_residual(sd, "pre.4", c, bn)
    outs = []
    for i in range(nstack):
        hg = _hg_hourglass(sd, f"hgs.{i}.0", c, 4, bn)
        f = _hg_residual(sd, f"features.{i}.0", hg, bn)
        f = F.relu(bn(_hg_conv(sd, f"features.{i}.1", f, 1), f"features.{i}.1.bn"))
        vec = F.conv2d(f, sd[f"outs_1.{i}.weight"], sd[f"outs_1.{i}.bias"])
        ht = F.conv2d(f, sd[f"outs_2.{i}.weight"], sd[f"outs_2.{i}.bias"])
        preds = torch.cat((vec, ht), dim=1)
        outs.append(preds)
        if i < nstack - 1:                                                    # :162-163
            c = c + _hg_conv(sd, f"merge_preds.{i}.conv", preds, 1) + _hg_conv(sd, f"merge_features.{i}.conv", f, 1)
    return outs


def hourglass_init(nstack: int, num_joints: int, seed: int, head_gain: float = 1.0) -> Dict[str, Tensor]:
    """Seeded state_dict with PoseNet's key schema (hourglass.py:105-142); torch-default-like
    uniform(-1/sqrt(fan_in), 1/sqrt(fan_in)) init.  head_gain scales outs_1/outs_2."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, Tensor] = {}

    def conv(p, ci, co, k, gain=1.0):
        bound = 1.0 / math.sqrt(ci * k * k)
        sd[p + ".weight"] = (torch.rand(co, ci, k, k, generator=g) * 2 - 1) * bound * gain
        sd[p + ".bias"] = (torch.rand(co, generator=g) * 2 - 1) * bound * gain

    def bnp(p, c):
        sd[p + ".weight"] = torch.ones(c)
        sd[p + ".bias"] = torch.zeros(c)
        sd[p + ".running_mean"] = torch.zeros(c)
        sd[p + ".running_var"] = torch.ones(c)
        sd[p + ".num_batches_tracked"] = torch.zeros((), dtype=torch.long)

    def residual(p, ci, co):
        bnp(p + ".bn1", ci); conv(p + ".conv1.conv", ci, co // 2, 1)
        bnp(p + ".bn2", co // 2); conv(p + ".conv2.conv", co // 2, co // 2, 3)
        bnp(p + ".bn3", co // 2); conv(p + ".conv3.conv", co // 2, co, 1)
        conv(p + ".skip_layer.conv", ci, co, 1)                               # always constructed (:38)

    def hourglass(p, n, f):
        residual(p + ".up1", f, f); residual(p + ".low1", f, f)
        if n > 1:
            hourglass(p + ".low2", n - 1, f)
        else:
            residual(p + ".low2", f, f)
        residual(p + ".low3", f, f)

    conv("pre.0.conv", 1, 64, 5); bnp("pre.0.bn", 64)
    residual("pre.1", 64, 128); residual("pre.3", 128, 256); residual("pre.4", 256, 256)
    for i in range(nstack):
        hourglass(f"hgs.{i}.0", 4, 256)
        residual(f"features.{i}.0", 256, 256)
        conv(f"features.{i}.1.conv", 256, 256, 1); bnp(f"features.{i}.1.bn", 256)
        conv(f"outs_1.{i}", 256, 3 * num_joints, 1, head_gain)
        conv(f"outs_2.{i}", 256, num_joints, 1, head_gain)
    for i in range(nstack - 1):
        conv(f"merge_features.{i}.conv.conv", 256, 256, 1)
        conv(f"merge_preds.{i}.conv.conv", 4 * num_joints, 256, 1)
    return sd


# --------------------------------------------------------------------------------------
# a.8  one training step   (train.py:107-131)  and Adam (train.py:67)
# --------------------------------------------------------------------------------------
def backbone_forward(sd, x, net: str, downsample: int, training: bool, new_stats=None) -> Tensor:
    """net = 'resnet_<L>' | 'hourglass_<S>'.  For hourglass returns the LAST stack (the one train.py:116-121 ends up supervising)."""
    kind, n = net.split("_")
    if kind == "resnet":
        return resnet_deconv_forward(sd, x, int(n), downsample, training, new_stats)
    return hourglass_forward(sd, x, int(n), training, new_stats)[-1]


def loss_and_grads(sd: Dict[str, Tensor], img: Tensor, jt_uvd_gt: Tensor, net: str, downsample: int,
                   kernel_size: float, coord_weight: float, dense_weight: float, all_stacks: bool = False):
    """Forward + backward of one train.py iteration (train-mode BN).  Returns
    (loss, loss_coord, loss_dense, uvd_pred, offset_pred, grads{name: tensor}, new_stats).
    all_stacks (hourglass_N only): supervise every stack and SUM the per-stack losses, the accumulation test.py:74-80 performs
    (`loss += loss_coord + loss_offset`); train.py:116-121 overwrites `loss` in that loop, so the default keeps the last stack only.
    loss_coord / loss_dense are then the sums over stacks, uvd_pred / offset_pred those of the last stack."""
    if all_stacks and net.startswith("hourglass"):
        params = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items() if v.is_floating_point()
                  and not k.endswith(("running_mean", "running_var"))}
        full = dict(sd); full.update(params)
        new_stats = {}
        Fs = img.shape[-1] // downsample
        offset_gt = joint2offset(jt_uvd_gt, img, kernel_size, Fs)
        preds = hourglass_forward(full, img, int(net.split("_")[1]), True, new_stats)
        l_coord = l_dense = 0.0
        for pred in preds:
            uvd = offset2joint_softmax(pred, img, kernel_size)
            l_coord = l_coord + smooth_l1(uvd, jt_uvd_gt)
            l_dense = l_dense + smooth_l1(pred, offset_gt)
        loss = coord_weight * l_coord + dense_weight * l_dense
        loss.backward()
        grads = {k: (p.grad if p.grad is not None else None) for k, p in params.items()}
        return loss.detach(), l_coord.detach(), l_dense.detach(), uvd.detach(), pred.detach(), grads, new_stats
    params = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items() if v.is_floating_point()
              and not k.endswith(("running_mean", "running_var"))}
    full = dict(sd); full.update(params)
    new_stats: dict = {}
    Fs = img.shape[-1] // downsample
    offset_gt = joint2offset(jt_uvd_gt, img, kernel_size, Fs)                 # train.py:113
    pred = backbone_forward(full, img, net, downsample, True, new_stats)      # :117/:123
    uvd = offset2joint_softmax(pred, img, kernel_size)                        # :118/:124
    l_coord = smooth_l1(uvd, jt_uvd_gt)
    l_dense = smooth_l1(pred, offset_gt)
    loss = coord_weight * l_coord + dense_weight * l_dense                    # :119-121
    loss.backward()
    grads = {k: (p.grad if p.grad is not None else None) for k, p in params.items()}
    return loss.detach(), l_coord.detach(), l_dense.detach(), uvd.detach(), pred.detach(), grads, new_stats


def adam_step(p: Tensor, g: Tensor, m: Tensor, v: Tensor, step: int, lr: float = 1e-3,
              b1: float = 0.9, b2: float = 0.999, eps: float = 1e-8, weight_decay: float = 0.0):
    """torch.optim.Adam (no amsgrad, L2-style weight decay) as used at train.py:67.  In-place; step is 1-based."""
    if weight_decay != 0.0:
        g = g + weight_decay * p
    m.mul_(b1).add_(g, alpha=1 - b1)
    v.mul_(b2).addcmul_(g, g, value=1 - b2)
    bc1 = 1 - b1 ** step
    bc2 = 1 - b2 ** step
    denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
    p.addcdiv_(m, denom, value=-lr / bc1)


# --------------------------------------------------------------------------------------
# mean 3-D error (util/eval_tool.py:20-58, util/util.py:13-20)  -- side metric for parity reports
# --------------------------------------------------------------------------------------
NYU_PARAS = (588.03, 587.07, 320.0, 240.0)       # dataloader/nyu_loader.py:23
NYU_FLIP = -1.0                                  # dataloader/nyu_loader.py:34


def uvd_norm_to_xyz(jt_uvd: Tensor, center_xyz: Tensor, M: Tensor, cube: Tensor, img_size: int,
                    paras=NYU_PARAS, flip: float = NYU_FLIP) -> Tensor:
    """Normalised crop UVD (B,J,3) -> camera XYZ in mm, batched form of EvalUtil.feed (eval_tool.py:34-43)."""
    uvd = jt_uvd.double().clone()
    uv = (uvd[..., :2] + 1) * img_size / 2.0
    dep = uvd[..., 2] * cube[:, None, 2].double() / 2.0 + center_xyz[:, None, 2].double()
    hom = torch.cat((uv, torch.ones_like(uv[..., :1])), dim=-1)               # (B,J,3)
    Minv = torch.linalg.inv(M.double())
    uv0 = torch.einsum("bij,bkj->bki", Minv, hom)[..., :2]
    x = (uv0[..., 0] - paras[2]) * dep / paras[0]
    y = (uv0[..., 1] - paras[3]) * dep / paras[1] * flip
    return torch.stack((x, y, dep), dim=-1).float()


def mean_3d_error_mm(jt_uvd_pred, jt_xyz_gt_norm, center_xyz, M, cube, img_size) -> float:
    xyz_pred = uvd_norm_to_xyz(jt_uvd_pred, center_xyz, M, cube, img_size)
    xyz_gt = jt_xyz_gt_norm * (cube[:, None, :] / 2.0) + center_xyz[:, None, :]  # eval_tool.py:45
    return (xyz_gt - xyz_pred).pow(2).sum(-1).sqrt().mean().item()


# --------------------------------------------------------------------------------------
# EvalUtil (util/eval_tool.py:5-122) restated in numpy with the reference's dtypes  -- SURVEY.md section 8 f.1
# --------------------------------------------------------------------------------------
def eval_feed_np(jt_uvd_pred, jt_xyz_gt, center_xyz, M, cube, img_size, paras=NYU_PARAS, flip=NYU_FLIP, return_xyz=False):
    """One sample (eval_tool.py:20-50).  Returns (jt_uvd_img (J,3) f32, euclidean_dist (J,) f32, diff_mean (3,) f32)."""
    import numpy as np
    uvd = np.array(jt_uvd_pred, dtype=np.float32).copy()
    gt = np.asarray(jt_xyz_gt, dtype=np.float32)
    center_xyz, M, cube = (np.asarray(t, dtype=np.float32) for t in (center_xyz, M, cube))
    M_inv = np.linalg.inv(M)                                                    # :33
    uvd[:, :2] = (uvd[:, :2] + 1) * img_size / 2.                               # :38
    uvd[:, 2] = uvd[:, 2] * cube[2] / 2. + center_xyz[2]                        # :39
    trans = np.hstack([uvd[:, :2], np.ones((uvd.shape[0], 1))])                 # :40 (float64)
    uvd[:, :2] = np.dot(M_inv, trans.T).T[:, :2]                                # :41
    xyz = uvd.copy()                                                            # util.py:15-19
    xyz[:, :2] = (xyz[:, :2] - paras[2:]) * xyz[:, 2:] / paras[:2]
    xyz[:, 1] *= flip
    gt_mm = gt * (cube / 2.) + center_xyz                                       # :45
    diff = gt_mm - xyz                                                          # :48
    if return_xyz:
        return xyz
    return uvd, np.sqrt(np.sum(np.square(diff), axis=1)), diff.mean(axis=0)     # :49-50


def eval_measures_np(dist, vis=None):
    """get_measures (eval_tool.py:80-122) over dist (N,J) float32 [with optional (N,J) visibility]."""
    import numpy as np
    trapz = getattr(np, "trapezoid", None) or np.trapz
    thresholds = np.linspace(0, 50, 100)
    norm_factor = trapz(np.ones_like(thresholds), thresholds)
    means, medians, aucs, curves = [], [], [], []
    for j in range(dist.shape[1]):
        data = np.array([dist[n, j] for n in range(dist.shape[0]) if vis is None or vis[n, j]])
        if len(data) == 0:
            continue
        means.append(np.mean(data)); medians.append(np.median(data))
        curve = np.array([np.mean((data <= t).astype('float')) for t in thresholds])
        curves.append(curve)
        aucs.append(trapz(curve, thresholds) / norm_factor)
    return np.mean(np.array(means)), np.mean(np.array(medians)), np.mean(np.array(aucs)), np.mean(np.array(curves), 0), thresholds


def eval_case_inputs(N: int, J: int, seed: int, img_size: int = 128):
    """Synthetic EvalUtil inputs: crop affines built like Loader.center2transmat (loader.py:210-240) with an in-plane rotation as the
    augmentation adds, hand centres around 760 mm, NYU cubes (300 mm, and 250 mm as for the second test subject, nyu_loader.py:32-33)."""
    import numpy as np
    rng = np.random.RandomState(seed)
    uvd = rng.uniform(-0.8, 0.8, (N, J, 3)).astype(np.float32)
    center = (np.array([0., 0., 760.]) + rng.normal(0, 40, (N, 3))).astype(np.float32)
    cube = np.where(rng.rand(N, 1) < 0.5, 300.0, 250.0).repeat(3, 1).astype(np.float32)
    Ms = []
    for n in range(N):
        s = img_size / rng.uniform(150, 260)
        th = rng.uniform(-0.6, 0.6) if n % 2 else 0.0
        t1 = np.eye(3); t1[0, 2], t1[1, 2] = -rng.uniform(150, 330), -rng.uniform(80, 250)
        sc = np.diag([s, s, 1.0])
        rot = np.array([[np.cos(th), -np.sin(th), 0], [np.sin(th), np.cos(th), 0], [0, 0, 1.0]])
        t2 = np.eye(3); t2[0, 2], t2[1, 2] = rng.randint(0, 12), rng.randint(0, 12)
        Ms.append((t2 @ rot @ sc @ t1).astype(np.float32))
    # ground truth = the prediction's own camera-space position + N(0, 9 mm) per axis (a few joints 60 mm off), cube-normalised as the
    # loader does (nyu_loader.py:64): errors then spread over the 0-50 mm PCK thresholds
    gt = np.zeros_like(uvd)
    for n in range(N):
        xyz = eval_feed_np(uvd[n], np.zeros((J, 3), np.float32), center[n], Ms[n], cube[n], img_size, return_xyz=True)
        noise = rng.normal(0, 9.0, (J, 3)) + (rng.rand(J, 1) < 0.05) * 60.0
        gt[n] = ((xyz + noise - center[n]) / (cube[n] / 2.)).astype(np.float32)
    vis = rng.rand(N, J) > 0.15
    return uvd, gt, center, np.stack(Ms), cube, vis


# --------------------------------------------------------------------------------------
# depth preprocessing of the non-augmented path (dataloader/loader.py:19-51,88-101,181-240; nyu_loader.py:38-60)  -- SURVEY 8 f.2
# --------------------------------------------------------------------------------------
def center2bounds_np(center, csize, paras=NYU_PARAS):
    """loader.py:181-188 (float64 arithmetic on a float32 centre, int() truncation)."""
    import numpy as np
    center, csize, p2 = np.asarray(center), np.asarray(csize, dtype=np.float64), np.asarray(paras[:2])
    ustart, vstart = center[:2] - (csize[:2] / 2.) / center[2] * p2 + 0.5
    uend, vend = center[:2] + (csize[:2] / 2.) / center[2] * p2 + 0.5
    return int(ustart), int(uend), int(vstart), int(vend), center[2] - csize[2] / 2., center[2] + csize[2] / 2.


def center2transmat_np(center, csize, dsize, paras=NYU_PARAS):
    """loader.py:210-240: crop affine (translate, isotropic scale, centring pad)."""
    import numpy as np
    ustart, uend, vstart, vend, _, _ = center2bounds_np(center, csize, paras)
    trans1 = np.eye(3); trans1[0][2] = -ustart; trans1[1][2] = -vstart
    w, h = (uend - ustart), (vend - vstart)
    scale = min(dsize[0] / w, dsize[1] / h)
    size = (int(w * scale), int(h * scale))
    sc = scale * np.eye(3); sc[2][2] = 1
    trans2 = np.eye(3)
    trans2[0][2] = int(np.floor(dsize[0] / 2. - size[0] / 2.)); trans2[1][2] = int(np.floor(dsize[1] / 2. - size[1] / 2.))
    return np.dot(trans2, np.dot(sc, trans1)).astype(np.float32)


def crop_normalize_np(depth, center_uvd, center_z, cube, img_size, paras=NYU_PARAS):
    """Loader.crop (loader.py:19-51) + Loader.normalize (:88-101) of one frame, cv2.resize(INTER_NEAREST) restated as its index
    rule sx = min(floor(x * src_w / dst_w), src_w - 1) [verified against cv2 4.13 on random sizes].  depth (Hs, Ws) float32 mm (0 = invalid);
    center_uvd (3,) float32; center_z the z of center_xyz used by normalize (nyu_loader.py:60).  Returns (img (D,D) float32, M (3,3) float32)."""
    import numpy as np
    dsize = np.array([img_size, img_size])
    ustart, uend, vstart, vend, zstart, zend = center2bounds_np(center_uvd, cube, paras)
    Hs, Ws = depth.shape
    # bounds2crop (:190-207): slice + zero pad to the full box, then clamp depths to the cube
    crop = np.zeros((vend - vstart, uend - ustart), np.float32)
    v0, v1, u0, u1 = max(vstart, 0), min(vend, Hs), max(ustart, 0), min(uend, Ws)
    if v1 > v0 and u1 > u0:
        crop[v0 - vstart:v1 - vstart, u0 - ustart:u1 - ustart] = depth[v0:v1, u0:u1]
    m1 = np.logical_and(crop < zstart, crop != 0); m2 = np.logical_and(crop > zend, crop != 0)
    crop[m1] = zstart; crop[m2] = 0
    w, h = (uend - ustart), (vend - vstart)
    scale = min(dsize[0] / w, dsize[1] / h)
    size = (int(w * scale), int(h * scale))
    sx = np.minimum(np.floor(np.arange(size[0]) * (1.0 / (float(size[0]) / w))).astype(int), w - 1)
    sy = np.minimum(np.floor(np.arange(size[1]) * (1.0 / (float(size[1]) / h))).astype(int), h - 1)
    resized = crop[sy][:, sx]
    res = np.zeros((img_size, img_size), np.float32)
    us, vs = (dsize - size) / 2.
    res[int(vs):int(vs + size[1]), int(us):int(us + size[0])] = resized
    # normalize (:88-101)
    cz = np.float64(center_z) if not isinstance(center_z, np.floating) else center_z
    half = np.asarray(cube, dtype=np.float64)[2] / 2.
    depth_max = res.max()
    res[res == depth_max] = cz + half
    res[res == 0] = cz + half
    out = np.clip(res.astype(np.float64), cz - half, cz + half)
    out = (out - cz) / half
    return out.astype(np.float32), center2transmat_np(center_uvd, cube, dsize, paras)


def preprocess_case_inputs(N: int, seed: int, Hs: int = 480, Ws: int = 640):
    """Synthetic raw NYU-like frames: a hand-sized blob at 500-1000 mm over a far wall, invalid (0) speckles, a frame whose crop box
    leaves the image on two sides, one with a very near object inside the box."""
    import numpy as np
    rng = np.random.RandomState(seed)
    yy, xx = np.mgrid[0:Hs, 0:Ws]
    frames, centers, cubes = [], [], []
    for n in range(N):
        z = rng.uniform(500, 1000)
        cu, cv_ = (rng.uniform(60, Ws - 60), rng.uniform(60, Hs - 60)) if n % 3 else (rng.choice([15.0, Ws - 20.0]), rng.choice([12.0, Hs - 15.0]))
        r = 588.0 * 90.0 / z
        blob = ((xx - cu) ** 2 + (yy - cv_) ** 2) < r * r
        d = np.full((Hs, Ws), rng.uniform(1500, 2500), np.float32) + rng.normal(0, 5, (Hs, Ws)).astype(np.float32)
        d[blob] = (z + 40 * np.sin(xx / 9.0) + 30 * np.cos(yy / 7.0) + rng.normal(0, 2, (Hs, Ws)))[blob]
        d[rng.rand(Hs, Ws) < 0.01] = 0
        if n % 4 == 1:
            d[int(cv_) - 5:int(cv_) + 5, int(cu) + 10:int(cu) + 25] = z - 400          # nearer than the cube front
        d = np.round(d).astype(np.float32)                                             # 16-bit integer millimetres on the wire
        frames.append(d)
        centers.append(np.array([cu, cv_, z], np.float32))
        cubes.append(np.array([300.0, 300.0, 300.0]) * (1.0 if n % 2 else 5.0 / 6.0))
    return np.stack(frames), np.stack(centers), np.stack(cubes)


# --------------------------------------------------------------------------------------
# training-time augmentation (dataloader/loader.py:53-179; nyu_loader.py:38-66)  -- SURVEY 8 f.2, second half
# cv2 is NOT used here: cv2.warpAffine / cv2.warpPerspective (float32, INTER_LINEAR, BORDER_CONSTANT), cv2.invert (3x3) and
# cv2.getRotationMatrix2D are third-party code (opencv-python 4.13.0 in this image; the reference pins none), restated from OpenCV's
# published algorithm (imgproc/src/imgwarp.cpp: WarpAffineInvoker / WarpPerspectiveInvoker + remapBilinear; core/src/lapack.cpp invert)
# and pinned bit-exactly against the reference running the real cv2 by tests/golden/augment_cases.npz.
# --------------------------------------------------------------------------------------
def _np_uvd2xyz(pts, paras, flip):
    """util/util.py:13-20 (float32 result)."""
    import numpy as np
    q = np.array(pts, copy=True).reshape(-1, 3)
    q[:, :2] = (q[:, :2] - paras[2:]) * q[:, 2:] / paras[:2]
    q[:, 1] *= flip
    return q.reshape(np.shape(pts)).astype(np.float32)


def _np_xyz2uvd(pts, paras, flip):
    """util/util.py:3-10 (float32 result)."""
    import numpy as np
    q = np.array(pts, copy=True).reshape(-1, 3)
    q[:, 1] *= flip
    q[:, :2] = q[:, :2] * paras[:2] / q[:, 2:] + paras[2:]
    return q.reshape(np.shape(pts)).astype(np.float32)


def cv_invert3x3_np(M):
    """cv2.invert of a 3x3 float64 matrix (DECOMP_LU's closed form for n = 3: adjugate times 1/det, OpenCV core/src/lapack.cpp)."""
    import numpy as np
    S = [[float(M[i][j]) for j in range(3)] for i in range(3)]
    d = S[0][0] * (S[1][1] * S[2][2] - S[1][2] * S[2][1]) - S[0][1] * (S[1][0] * S[2][2] - S[1][2] * S[2][0]) + \
        S[0][2] * (S[1][0] * S[2][1] - S[1][1] * S[2][0])
    if d == 0.0:
        return np.zeros((3, 3))
    d = 1.0 / d
    return np.array([(S[1][1] * S[2][2] - S[1][2] * S[2][1]) * d, (S[0][2] * S[2][1] - S[0][1] * S[2][2]) * d, (S[0][1] * S[1][2] - S[0][2] * S[1][1]) * d,
                     (S[1][2] * S[2][0] - S[1][0] * S[2][2]) * d, (S[0][0] * S[2][2] - S[0][2] * S[2][0]) * d, (S[0][2] * S[1][0] - S[0][0] * S[1][2]) * d,
                     (S[1][0] * S[2][1] - S[1][1] * S[2][0]) * d, (S[0][1] * S[2][0] - S[0][0] * S[2][1]) * d, (S[0][0] * S[1][1] - S[0][1] * S[1][0]) * d]).reshape(3, 3)


def cv_rotation_matrix_2d_np(center, angle_deg, scale=1.0):
    """cv2.getRotationMatrix2D (imgproc/src/imgwarp.cpp): libm cos/sin of the angle in radians, float64 2x3."""
    import math
    import numpy as np
    a = angle_deg * math.pi / 180.0
    al, be = math.cos(a) * scale, math.sin(a) * scale
    return np.array([[al, be, (1 - al) * center[0] - be * center[1]], [-be, al, be * center[0] + (1 - al) * center[1]]])


def _bilinear_fixed_np(img, X, Y, border):
    """remapBilinear on OpenCV's fixed-point map: X, Y are source coordinates in 1/32 pixel (INTER_BITS = 5); weights come from the
    32x32 float table (1-fx, fx) x (1-fy, fy) (products rounded to float32); taps outside the image read `border`; the four products
    are summed left to right in float32."""
    import numpy as np
    H, W = img.shape
    sx, sy = np.clip(X >> 5, -32768, 32767), np.clip(Y >> 5, -32768, 32767)
    fx = (X & 31).astype(np.float32) * np.float32(1.0 / 32)
    fy = (Y & 31).astype(np.float32) * np.float32(1.0 / 32)
    one = np.float32(1)
    w = [(one - fy) * (one - fx), (one - fy) * fx, fy * (one - fx), fy * fx]

    def tap(yy, xx):
        ok = (yy >= 0) & (yy < H) & (xx >= 0) & (xx < W)
        return np.where(ok, img[np.clip(yy, 0, H - 1), np.clip(xx, 0, W - 1)], np.float32(border)).astype(np.float32)
    t = [tap(sy, sx), tap(sy, sx + 1), tap(sy + 1, sx), tap(sy + 1, sx + 1)]
    return (t[0] * w[0] + t[1] * w[1] + t[2] * w[2] + t[3] * w[3]).astype(np.float32)


def cv_warp_affine_linear_np(img, M, border=0.0):
    """cv2.warpAffine(img float32 (H,W), M 2x3, (W,H), INTER_LINEAR, BORDER_CONSTANT, border): the 2x3 map is inverted in closed form,
    coordinates are 10-bit fixed point (AB_SCALE = 1024) with the x and y terms rounded separately (round-half-even), + 16, >> 5."""
    import numpy as np
    m = np.asarray(M, dtype=np.float64).ravel().copy()
    D = m[0] * m[4] - m[1] * m[3]
    D = 1.0 / D if D != 0 else 0.0
    a11, a22 = m[4] * D, m[0] * D
    m[0], m[1], m[3], m[4] = a11, m[1] * -D, m[3] * -D, a22
    b1, b2 = -m[0] * m[2] - m[1] * m[5], -m[3] * m[2] - m[4] * m[5]
    m[2], m[5] = b1, b2
    H, W = img.shape
    x, y = np.arange(W), np.arange(H)
    ad, bd = np.rint(m[0] * x * 1024).astype(np.int64), np.rint(m[3] * x * 1024).astype(np.int64)
    X0 = np.rint((m[1] * y + m[2]) * 1024).astype(np.int64) + 16
    Y0 = np.rint((m[4] * y + m[5]) * 1024).astype(np.int64) + 16
    return _bilinear_fixed_np(img, (X0[:, None] + ad[None, :]) >> 5, (Y0[:, None] + bd[None, :]) >> 5, border)


def cv_warp_block_np(H, W):
    """Tile of WarpPerspectiveInvoker (BLOCK_SZ = 32): the row term of the homography is evaluated at the tile's first column."""
    bh = min(16, H)
    bw = min(1024 // bh, W)
    bh = min(1024 // bw, H)
    return bw, bh


def cv_warp_perspective_linear_np(img, M, border=0.0):
    """cv2.warpPerspective(img float32 (H,W), M 3x3, (W,H), INTER_LINEAR, BORDER_CONSTANT, border): M is inverted (cv2.invert), each pixel's
    source coordinate is ((X0 + m0*x1) * 32/W) rounded half-even to 1/32 pixel, X0 = m0*bx + m1*y + m2 per tile row."""
    import numpy as np
    m = cv_invert3x3_np(np.asarray(M, dtype=np.float64)).ravel()
    H, W = img.shape
    bw, _ = cv_warp_block_np(H, W)
    y = np.arange(H, dtype=np.float64)[:, None]
    xs = np.arange(W)
    bx = (xs // bw * bw).astype(np.float64)[None, :]
    x1 = (xs % bw).astype(np.float64)[None, :]
    X0 = m[0] * bx + m[1] * y + m[2]
    Y0 = m[3] * bx + m[4] * y + m[5]
    W0 = m[6] * bx + m[7] * y + m[8]
    Wv = W0 + m[6] * x1
    with np.errstate(divide="ignore"):
        Wv = np.where(Wv != 0, 32.0 / np.where(Wv != 0, Wv, 1.0), 0.0)
    lo, hi = -2147483648.0, 2147483647.0
    X = np.rint(np.maximum(lo, np.minimum(hi, (X0 + m[0] * x1) * Wv))).astype(np.int64)
    Y = np.rint(np.maximum(lo, np.minimum(hi, (Y0 + m[3] * x1) * Wv))).astype(np.int64)
    return _bilinear_fixed_np(img, X, Y, border)


def crop_np(depth, center_uvd, cube, img_size, paras=NYU_PARAS):
    """Loader.crop alone (loader.py:19-51): the un-normalised crop in millimetres (0 = background) and the crop affine."""
    import numpy as np
    dsize = np.array([img_size, img_size])
    ustart, uend, vstart, vend, zstart, zend = center2bounds_np(center_uvd, cube, paras)
    Hs, Ws = depth.shape
    box = np.zeros((vend - vstart, uend - ustart), np.float32)
    v0, v1, u0, u1 = max(vstart, 0), min(vend, Hs), max(ustart, 0), min(uend, Ws)
    if v1 > v0 and u1 > u0:
        box[v0 - vstart:v1 - vstart, u0 - ustart:u1 - ustart] = depth[v0:v1, u0:u1]
    m1 = np.logical_and(box < zstart, box != 0); m2 = np.logical_and(box > zend, box != 0)
    box[m1] = zstart; box[m2] = 0
    w, h = (uend - ustart), (vend - vstart)
    scale = min(dsize[0] / w, dsize[1] / h)
    size = (int(w * scale), int(h * scale))
    sx = np.minimum(np.floor(np.arange(size[0]) * (1.0 / (float(size[0]) / w))).astype(int), w - 1)
    sy = np.minimum(np.floor(np.arange(size[1]) * (1.0 / (float(size[1]) / h))).astype(int), h - 1)
    res = np.zeros((img_size, img_size), np.float32)
    us, vs = (dsize - size) / 2.
    res[int(vs):int(vs + size[1]), int(us):int(us + size[0])] = box[sy][:, sx]
    return res, center2transmat_np(center_uvd, cube, dsize, paras)


def _recrop_np(img, center, cube, M_new, M_old, nv_val, paras):
    """Loader.recrop (loader.py:125-139): perspective warp by M_new * inv(M_old) (float32 product of a float32 LAPACK inverse, as numpy
    computes it), pixels below nv_val -> 0, then the cube clamp of bounds2crop."""
    import numpy as np
    out = cv_warp_perspective_linear_np(img, np.dot(M_new, np.linalg.inv(M_old)), 0.0)
    out[out < nv_val] = 0.0
    _, _, _, _, zstart, zend = center2bounds_np(center, cube, paras)
    m1 = np.logical_and(out < zstart, out != 0); m2 = np.logical_and(out > zend, out != 0)
    out[m1] = zstart; out[m2] = 0.
    return out.astype(np.float32)


def _normalize_np(depth_max, img, center_z, half):
    """Loader.normalize (loader.py:88-101) with float64 bounds; returns float64 like the reference (the caller casts)."""
    import numpy as np
    img = img.copy()
    img[img == depth_max] = center_z + half
    img[img == 0] = center_z + half
    out = np.clip(img.astype(np.float64), center_z - half, center_z + half)
    return (out - center_z) / half


def rotate_pts_np(pt, center, angle):
    """loader.py:242-252."""
    import numpy as np
    alpha = angle * np.pi / 180.
    r = pt.copy()
    r[:, 0] = (pt[:, 0] - center[0]) * np.cos(alpha) - (pt[:, 1] - center[1]) * np.sin(alpha)
    r[:, 1] = (pt[:, 0] - center[0]) * np.sin(alpha) + (pt[:, 1] - center[1]) * np.cos(alpha)
    r[:, :2] += center[:2]
    return r.astype(np.float32)


def nyu_train_item_np(depth, jt_xyz, center_xyz, cube, img_size, aug_op, trans, scale, rot, paras=NYU_PARAS, flip=NYU_FLIP):
    """NYU.__getitem__ of the training phase (nyu_loader.py:38-66) for one frame with the augmentation already drawn
    (Loader.random_aug, loader.py:53-72): crop -> augment (translate | rotate | scale | none, loader.py:74-86,103-179) -> normalize ->
    labels.  depth (Hs,Ws) float32 mm; jt_xyz (J,3) float64 mm; center_xyz (3,) float64 mm; cube (3,) int64/float64 mm.
    Returns (img (1,D,D) f32, jt_xyz_norm (J,3) f32, jt_uvd_norm (J,3) f32, center_xyz (3,) f32, M (3,3) f32, cube (3,) f32)."""
    import numpy as np
    paras = np.asarray(paras, dtype=np.float64) if not isinstance(paras, tuple) else paras
    P = np.asarray(paras, dtype=np.float64)
    cube = np.asarray(cube)
    center_uvd = _np_xyz2uvd(np.asarray(center_xyz, dtype=np.float64), P, flip)
    jt = np.asarray(jt_xyz, dtype=np.float64) - center_xyz
    img, M = crop_np(depth, center_uvd, cube, img_size, paras)
    depth_max = img.max()
    center = center_uvd
    if aug_op == "trans" and not np.allclose(trans, 0.):
        new_center = _np_xyz2uvd(_np_uvd2xyz(center, P, flip) + trans, P, flip)
        if not np.allclose(center[2], 0.) or np.allclose(new_center[2], 0.):
            new_M = center2transmat_np(new_center, cube, np.array(img.shape), paras)
            img = _recrop_np(img, new_center, cube, new_M, M, np.min(img[img > 0]) - 1, paras)
        else:
            new_M = M
        jt = jt + _np_uvd2xyz(center, P, flip) - _np_uvd2xyz(new_center, P, flip)
        center, M = new_center, new_M
    elif aug_op == "rot":
        r = np.mod(rot, 360)
        img = cv_warp_affine_linear_np(img, cv_rotation_matrix_2d_np((img.shape[1] // 2, img.shape[0] // 2), -r, 1), 0.0)
        c_xyz = _np_uvd2xyz(center, P, flip)
        uvd = rotate_pts_np(_np_xyz2uvd(jt + c_xyz, P, flip), center, r)
        jt = _np_uvd2xyz(uvd, P, flip) - c_xyz
    elif aug_op == "scale" and not np.allclose(scale, 1.):
        new_cube = cube * scale
        if not np.allclose(center[2], 0.):
            new_M = center2transmat_np(center, new_cube, np.array(img.shape), paras)
            img = _recrop_np(img, center, new_cube, new_M, M, np.min(img[img > 0]) - 1, paras)
        else:
            new_M = M
        cube, M = new_cube, new_M
    img = _normalize_np(depth_max, img, center[2], cube[2] / 2.)
    c_xyz = _np_uvd2xyz(center, P, flip)
    q = _np_xyz2uvd(jt + c_xyz, P, flip)
    h = np.hstack([q[:, :2], np.ones((q.shape[0], 1))])
    h = np.dot(M, h.T).T
    h[:, :2] /= h[:, 2:]
    jt_uvd = np.hstack([h[:, :2], q[:, 2:]]).astype(np.float32)
    jt_uvd[:, :2] = jt_uvd[:, :2] / (img_size / 2.) - 1
    jt_uvd[:, 2] = (jt_uvd[:, 2] - c_xyz[2]) / (cube[2] / 2.0)
    jt_n = jt / (cube / 2.)
    return (img[np.newaxis, :].astype(np.float32), jt_n.astype(np.float32), jt_uvd.astype(np.float32), c_xyz.astype(np.float32),
            M.astype(np.float32), cube.astype(np.float32))


def augment_case_inputs(N: int, seed: int, J: int = 14):
    """Raw frames of preprocess_case_inputs plus float64 centre / joint labels in millimetres as nyu_loader.make_dataset would hold them."""
    import numpy as np
    frames, centers, _ = preprocess_case_inputs(N, seed)
    rng = np.random.RandomState(seed + 1000)
    P = np.asarray(NYU_PARAS)
    center_xyz = np.stack([_np_uvd2xyz(c.astype(np.float64), P, NYU_FLIP).astype(np.float64) for c in centers])
    jt_xyz = center_xyz[:, None, :] + rng.uniform(-90, 90, (N, J, 3))
    return frames, jt_xyz, center_xyz


# --------------------------------------------------------------------------------------
# synthetic inputs (SURVEY.md section 8 d)
# --------------------------------------------------------------------------------------
def synthetic_batch(B: int, H: int, J: int, seed: int) -> Tuple[Tensor, Tensor]:
    """Depth crop (B,1,H,H): background == +1.0, elliptical foreground blob with depth in [-1,0.98];
    jt_uvd_gt (B,J,3) ~ U(-0.5,0.5).  CPU-generator seeded so the same batch exists on every box."""
    g = torch.Generator().manual_seed(seed)
    ax = (torch.arange(H).float() + 0.5) / H * 2 - 1
    yy, xx = torch.meshgrid(ax, ax, indexing="ij")
    cx = (torch.rand(B, 1, 1, generator=g) - 0.5) * 0.4
    cy = (torch.rand(B, 1, 1, generator=g) - 0.5) * 0.4
    rx = 0.45 + 0.3 * torch.rand(B, 1, 1, generator=g)
    ry = 0.45 + 0.3 * torch.rand(B, 1, 1, generator=g)
    inside = ((xx - cx) / rx) ** 2 + ((yy - cy) / ry) ** 2 < 1.0
    ramp = 0.5 * (xx - cx) + 0.3 * (yy - cy)
    depth = (ramp + 0.05 * torch.randn(B, H, H, generator=g)).clamp(-1.0, 0.98)
    img = torch.where(inside, depth, torch.ones_like(depth)).unsqueeze(1).contiguous()
    jt = torch.rand(B, J, 3, generator=g) - 0.5
    return img.float(), jt.float()


def head_case_inputs(B: int, J: int, Fs: int, H: int, ks: float, seed: int):
    """Inputs of the head/loss golden cases (tests/golden/make_golden.py): depth crop, GT joints, a peaked
    prediction volume (GT volume + 0.05 N(0,1)) and an upstream gradient for the UVD output."""
    g = torch.Generator().manual_seed(seed)
    img, jt = synthetic_batch(B, H, J, seed)
    if seed == 13:                      # depth values straddling the 0.99 threshold
        img[0, 0, :8] = 0.9899
        img[0, 0, 8:16] = 0.99
        img[0, 0, 16:24] = 0.9901
    if seed == 15:
        img[1] = 1.0                     # all background
        img[0] = img[0].clamp(max=0.5)   # all foreground
    gt = joint2offset(jt, img, ks, Fs)
    pred = gt + 0.05 * torch.randn(gt.shape, generator=g)
    g_uvd = torch.randn(B, J, 3, generator=g)
    return img, jt, pred, g_uvd


HEAD_CASES = [(2, 14, 64, 128, 1.0, 11), (2, 14, 64, 128, 0.4, 12), (1, 21, 32, 128, 0.4, 13),
              (3, 16, 128, 128, 1.0, 14), (2, 14, 16, 128, 1.0, 15), (2, 14, 128, 256, 0.4, 16)]
