"""nn.Module drop-ins for the reference backbones, executed by engine.Plan on libawr_b200.so.

    get_deconv_net(layers, num_classes, downsample)   <- model/resnet_deconv.py:8-16
    PoseNet(net, joint_num, ...)                      <- model/hourglass.py:105-165

Same constructor signatures, forward() signatures / return types, parameter + buffer names, shapes and dtypes
(fp32) as the reference, so `results/*.pth` checkpoints load with strict=True and train.py / test.py run
unchanged.  The canonical nn.Parameters are (possibly permuted) views of one flat fp32 buffer laid out in the
order the kernels read, so there is no repack step between the optimizer and the kernels.
"""
import math
import os

import torch
import torch.nn as nn

from . import _lib as L
from . import engine as E


class ParamStore:
    """Flat device storage behind a module's parameters: params / grads (fp32), bf16 shadow, BN buffers."""

    def __init__(self, layout, device):
        self.layout, self.device = layout, device
        n = max(layout.total, 64)
        self.params = torch.zeros(n, dtype=torch.float32, device=device)
        self.grads = torch.zeros(n, dtype=torch.float32, device=device)
        self.shadow = torch.zeros(n, dtype=torch.bfloat16, device=device)
        # bit-reproducible build: the weight-gradient kernels accumulate into 16-byte order-independent slots (same element offsets as
        # `grads`), folded into `grads` at the end of backward
        self.grads_acc = torch.zeros(n, 2, dtype=torch.int64, device=device) if L.deterministic() else None
        self.buffers = {}
        self.shadow_version = -1

    def view(self, name):
        return self.layout.view(self.params, name)

    def grad_view(self, name):
        return self.layout.view(self.grads, name)

    def refresh_shadow(self):
        L.check(L.lib().awr_cast_f32_to_bf16(self.params.data_ptr(), self.shadow.data_ptr(), self.params.numel(), L.stream()),
                "awr_cast_f32_to_bf16")


def _get_or_make(root, path):
    m = root
    for part in path:
        if part not in m._modules:
            m.add_module(part, nn.Module())
        m = m._modules[part]
    return m


class AWRBackbone(nn.Module):
    """Common implementation; see get_deconv_net / PoseNet for the two public constructors."""

    def __init__(self, net, joint_num, downsample, layout, precision=None):
        super().__init__()
        self._net, self._J, self._ds = net, joint_num, downsample
        object.__setattr__(self, "_layout", layout)
        object.__setattr__(self, "_store", None)
        object.__setattr__(self, "_plans", {})
        object.__setattr__(self, "_pnames", list(layout.order))
        self.precision = precision or os.environ.get("AWR_B200_PRECISION", "fp32")
        # canonical parameters / buffers under the reference's hierarchical names
        for name in layout.order:
            s = layout.specs[name]
            *path, leaf = name.split(".")
            _get_or_make(self, path).register_parameter(leaf, nn.Parameter(torch.zeros(s.shape)))
        for name, (shape, dtype) in layout.buffers.items():
            *path, leaf = name.split(".")
            init = torch.ones(shape) if leaf == "running_var" else torch.zeros(shape, dtype=dtype)
            _get_or_make(self, path).register_buffer(leaf, init)
        # parameter order == module-tree traversal order == the reference's parameters() order
        object.__setattr__(self, "_pnames", [k for k, _ in self.named_parameters()])
        object.__setattr__(self, "_params_dirty", False)
        # a load_state_dict() after a FusedTrainer was built must reach the trainer's bf16 weight shadow: flag it, the trainer refreshes
        self.register_load_state_dict_post_hook(lambda module, incompatible: object.__setattr__(module, "_params_dirty", True))

    # ---- storage management ---------------------------------------------------------------------------
    def _params_by_name(self):
        return dict(self.named_parameters())

    def _adopt(self, device):
        """(Re)create the flat store on `device`, move every parameter's current value into its view and re-point
        param.data at the view.  Parameter objects are preserved (optimizers keep working)."""
        store = ParamStore(self._layout, device)
        pn = self._params_by_name()
        with torch.no_grad():
            for name in self._pnames:
                p = pn[name]
                v = store.view(name)
                v.copy_(p.data.to(device=device, dtype=torch.float32))
                p.data = v
        for name, b in self.named_buffers():
            if b.device != device:
                raise RuntimeError("module buffers and parameters live on different devices")
            store.buffers[name] = b
        object.__setattr__(self, "_store", store)
        object.__setattr__(self, "_plans", {})
        object.__setattr__(self, "_sig", self._signature())
        return store

    def _signature(self):
        first_p = next(self.parameters())
        bufs = tuple(b.data_ptr() for b in self.buffers())
        return (first_p.data_ptr(), first_p.device, bufs)

    def store(self):
        p0 = next(self.parameters())
        if not p0.is_cuda:
            raise RuntimeError("awr_b200 backbones run on CUDA only: call .cuda() first (sm_100a kernels; no CPU fallback)")
        st = self._store
        if st is None or self._sig != self._signature() or any(
                p.data_ptr() != st.view(n).data_ptr() for n, p in zip(self._pnames[:4], list(self.parameters())[:4])):
            st = self._adopt(p0.device)
        return st

    def plan(self, B, H, training):
        st = self.store()
        key = (B, H, bool(training), self.precision)
        pl = self._plans.get(key)
        if pl is None:
            pl = E.Plan(self._net, self._J, self._ds, B, H, self.precision, training, st, st.device)
            pl.version = 0
            self._plans[key] = pl
        return pl

    # ---- forward ----------------------------------------------------------------------------------------
    def _run(self, x):
        if x.dim() != 4 or x.shape[1] != 1 or x.shape[2] != x.shape[3]:
            raise ValueError("expected input of shape (B,1,H,H)")
        L.require_cuda(x)
        params = list(self.parameters())
        outs = _BackboneFn.apply(self, x, *params)
        return list(outs)


class _BackboneFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, x, *params):
        B, H = x.shape[0], x.shape[-1]
        pl = module.plan(B, H, module.training)
        st = module._store
        if pl.precision == "bf16":
            st.refresh_shadow()
        pl.img.copy_(x.detach().to(torch.float32))
        pl.arena_used().zero_()
        pl.run_forward()
        pl.version += 1
        ctx.module, ctx.pl, ctx.version = module, pl, pl.version
        ctx.n_params = len(params)
        return tuple(h.pred.clone() for h in pl.heads)

    @staticmethod
    def backward(ctx, *gouts):
        pl, module = ctx.pl, ctx.module
        if not pl.training:
            raise NotImplementedError("backward through an eval-mode (running-statistics) forward is not supported")
        if pl.version != ctx.version:
            raise RuntimeError("only the most recent forward of this module can be back-propagated (static activation plan)")
        st = module._store
        st.grads.zero_()
        for h, g in zip(pl.heads, gouts):
            if g is None:
                h.dpred.zero_()
            else:
                h.dpred.copy_(g)
        pl.run_backward()
        # Always hand autograd private copies: AccumulateGrad may install (steal) the returned tensor as p.grad, and a view of the shared
        # flat buffer would then be wiped by the next backward's zero fill / rewritten by its kernels (doubling every gradient under
        # zero_grad(set_to_none=False) and losing the first micro-batch under gradient accumulation).  One flat copy, then views of it.
        flat = st.grads.clone()
        lay = st.layout
        grads = [lay.view(flat, name).contiguous() for name in module._pnames]
        return (None, None) + tuple(grads)


# --------------------------------------------------------------------------------------------------------
# public constructors
# --------------------------------------------------------------------------------------------------------
class ResnetDeconv(AWRBackbone):
    def forward(self, x):
        return self._run(x)[0]                                   # (B,4J,H/ds,W/ds) fp32 NCHW  (resnet_deconv.py:136)


class PoseNetB200(AWRBackbone):
    def forward(self, imgs):
        return self._run(imgs)                                   # list of nstack tensors (hourglass.py:165)


def get_deconv_net(layers, num_classes, downsample, precision=None):
    """Same call as the reference (model/resnet_deconv.py:8): layers in {18,50,101,152}, num_classes = joints,
    downsample in {1,2,4}.  Weights initialised with the reference's distributions (resnet_deconv.py:93-115)."""
    if layers not in E.RESNET_SPEC:
        raise KeyError(layers)
    if downsample not in (1, 2, 4):
        raise ValueError("downsample must be 1, 2 or 4")
    lay = E.resnet_layout(layers, num_classes, downsample)
    m = ResnetDeconv(f"resnet_{layers}", num_classes, downsample, lay, precision)
    with torch.no_grad():
        for name, p in m.named_parameters():
            s = lay.specs[name]
            if s.kind == "conv":
                co, ci, k, _ = s.shape
                p.normal_(0, math.sqrt(2.0 / (k * k * co)))
            elif s.kind in ("deconv", "headw"):
                p.normal_(0, 0.001)
            elif name.endswith(".weight"):
                p.fill_(1.0)                                     # BN gamma
            else:
                p.zero_()                                        # BN beta, head biases
    return m


def PoseNet(net, joint_num, inp_dim=256, bn=False, increase=0, precision=None, **kwargs):
    """Same call as the reference (model/hourglass.py:106): net = 'hourglass_<nstack>'."""
    if inp_dim != 256 or increase != 0:
        raise NotImplementedError("only the reference's shipped configuration (inp_dim=256, increase=0) is built")
    nstack = int(net.split("_")[-1])
    lay = E.hourglass_layout(nstack, joint_num)
    m = PoseNetB200(f"hourglass_{nstack}", joint_num, 2, lay, precision)
    m.nstack, m.joint_num = nstack, joint_num
    with torch.no_grad():
        pn = dict(m.named_parameters())
        for name, p in pn.items():
            s = lay.specs[name]
            if s.kind in ("conv", "headw"):
                fan_in = s.shape[1] * s.shape[2] * s.shape[3]
                bound = 1.0 / math.sqrt(fan_in)                  # kaiming_uniform(a=sqrt(5)) == U(-1/sqrt(fan_in), ..)
                p.uniform_(-bound, bound)
                bname = name[: -len("weight")] + "bias"
                if bname in pn:
                    pn[bname].uniform_(-bound, bound)
            elif name.endswith(".weight"):
                p.fill_(1.0)
            elif ".bn" in name or name.endswith("bn.bias"):
                p.zero_()
    return m
