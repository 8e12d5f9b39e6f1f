"""Static execution plans for the AWR backbones on top of the C-ABI kernels.

A `Plan` is built once per (network, batch, image size, precision, train/eval): it owns every activation /
gradient buffer (NHWC, PyTorch-allocated), and two flat lists of kernel launches (forward, backward) with all
pointers bound, so a step is a pure launch sequence that can be CUDA-graph captured.  Python only orders
launches; every arithmetic op runs in libawr_b200.so.

Network structure follows the reference modules it replaces:
  ResNet-deconv : model/resnet_deconv.py:19-215      Hourglass : model/hourglass.py:6-165
"""
import math
import os

import torch

from . import _lib as L

BN_EPS = 1e-5
BN_MOMENTUM = 0.1
RESNET_SPEC = {18: ("basic", [2, 2, 2, 2]), 50: ("bottleneck", [3, 4, 6, 3]),
               101: ("bottleneck", [3, 4, 23, 3]), 152: ("bottleneck", [3, 8, 36, 3])}


def _align(n, a=64):
    return (n + a - 1) // a * a


# ======================================================================================================
# parameters: one flat fp32 buffer (+ grads, + bf16 shadow); canonical nn.Parameter tensors are VIEWS of it
# ======================================================================================================
class ParamSpec:
    __slots__ = ("name", "kind", "shape", "offset", "numel", "phys_shape")

    def __init__(self, name, kind, shape, offset, phys_shape):
        self.name, self.kind, self.shape, self.offset, self.phys_shape = name, kind, tuple(shape), offset, tuple(phys_shape)
        self.numel = int(math.prod(phys_shape))


class ParamLayout:
    """Physical order of every weight is what the kernels read:
         conv   (Co,Ci,kh,kw) canonical  -> physical [kh][kw][Co][Ci]      (canonical = phys.permute(2,3,0,1))
         deconv (Ci,Co,kh,kw) canonical  -> physical [kh][kw][Co][Ci]      (canonical = phys.permute(3,2,0,1))
         head   two 1x1 convs (3J / J rows) share one zero-padded [align64(4J)][Cin] block; biases share one [align64(4J)] block
         vec    (C,)"""

    def __init__(self):
        self.specs = {}
        self.order = []
        self.total = 0
        self.buffers = {}        # non-trainable: name -> (shape, dtype)
        self.groups = {}         # physical blocks shared by several canonical params: gname -> (offset, phys_shape)

    def _alloc(self, numel):
        off = self.total
        self.total = _align(off + numel)
        return off

    def conv(self, name, co, ci, k, ci_pad=None):
        """ci_pad: physical input-channel count (zero-padded) when the consumer activation is channel-padded."""
        self._add(name, "conv", (co, ci, k, k), (k, k, co, ci_pad or ci))

    def deconv(self, name, ci, co, k):
        self._add(name, "deconv", (ci, co, k, k), (k, k, co, ci))

    def vec(self, name, c):
        self._add(name, "vec", (c,), (c,))

    def _add(self, name, kind, shape, phys):
        spec = ParamSpec(name, kind, shape, self._alloc(int(math.prod(phys))), phys)
        self.specs[name] = spec
        self.order.append(name)

    def head(self, gname, names_rows, cin, bias_names):
        """names_rows: [(param name, rows)] stacked into one [align64(sum rows)][cin] block (rows beyond the sum stay zero):
        64 rows for the 14/16-joint datasets, 128 for the 21-joint ones (config.py:1-6)."""
        hc = _align(sum(r for _, r in names_rows))
        off = self._alloc(hc * cin)
        self.groups[gname + ".weight"] = (off, (hc, cin))
        r = 0
        for n, rows in names_rows:
            s = ParamSpec(n, "headw", (rows, cin, 1, 1), off + r * cin, (rows, cin))
            self.specs[n] = s
            self.order.append(n)
            r += rows
        boff = self._alloc(hc)
        self.groups[gname + ".bias"] = (boff, (hc,))
        r = 0
        for (n, rows) in zip(bias_names, [x[1] for x in names_rows]):
            s = ParamSpec(n, "vec", (rows,), boff + r, (rows,))
            self.specs[n] = s
            self.order.append(n)
            r += rows

    def bn(self, prefix, c):
        self.vec(prefix + ".weight", c)
        self.vec(prefix + ".bias", c)
        self.buffers[prefix + ".running_mean"] = ((c,), torch.float32)
        self.buffers[prefix + ".running_var"] = ((c,), torch.float32)
        self.buffers[prefix + ".num_batches_tracked"] = ((), torch.long)

    def view(self, flat, name):
        s = self.specs[name]
        t = flat[s.offset: s.offset + s.numel].view(s.phys_shape)
        if s.kind == "conv":
            return t[..., : s.shape[1]].permute(2, 3, 0, 1)
        if s.kind == "deconv":
            return t.permute(3, 2, 0, 1)
        if s.kind == "headw":
            return t.view(s.shape)
        return t

    def phys(self, flat, name):
        if name in self.groups:
            off, shp = self.groups[name]
            return flat[off: off + int(math.prod(shp))].view(shp)
        s = self.specs[name]
        return flat[s.offset: s.offset + s.numel].view(s.phys_shape)


def resnet_layout(layers, J, downsample):
    """Key schema of ResnetDeconv.state_dict() (resnet_deconv.py:31-53)."""
    kind, counts = RESNET_SPEC[layers]
    exp = 1 if kind == "basic" else 4
    lay = ParamLayout()
    lay.conv("pre.0.weight", 64, 1, 5)
    lay.bn("pre.1", 64)
    inpl = 64
    for li, (planes, nblk) in enumerate(zip([64, 128, 256, 512], counts), start=1):
        for bi in range(nblk):
            p = f"layer{li}.{bi}"
            stride = 2 if (li > 1 and bi == 0) else 1
            if kind == "basic":
                lay.conv(p + ".conv1.weight", planes, inpl, 3); lay.bn(p + ".bn1", planes)
                lay.conv(p + ".conv2.weight", planes, planes, 3); lay.bn(p + ".bn2", planes)
            else:
                lay.conv(p + ".conv1.weight", planes, inpl, 1); lay.bn(p + ".bn1", planes)
                lay.conv(p + ".conv2.weight", planes, planes, 3); lay.bn(p + ".bn2", planes)
                lay.conv(p + ".conv3.weight", planes * 4, planes, 1); lay.bn(p + ".bn3", planes * 4)
            if bi == 0 and (stride != 1 or inpl != planes * exp):
                lay.conv(p + ".downsample.0.weight", planes * exp, inpl, 1); lay.bn(p + ".downsample.1", planes * exp)
            inpl = planes * exp
    for i in range(4 - int(math.log(downsample, 2))):
        lay.deconv(f"deconv_layers.{3 * i}.weight", inpl, 256, 4)
        lay.bn(f"deconv_layers.{3 * i + 1}", 256)
        inpl = 256
    lay.head("final", [("final1.weight", 3 * J), ("final2.weight", J)], 256, ["final1.bias", "final2.bias"])
    return lay


def hourglass_layout(nstack, J):
    """Key schema of PoseNet.state_dict() (hourglass.py:105-142)."""
    lay = ParamLayout()

    def conv(p, ci, co, k):
        lay.conv(p + ".weight", co, ci, k)
        lay.vec(p + ".bias", co)

    def residual(p, ci, co):
        lay.bn(p + ".bn1", ci); conv(p + ".conv1.conv", ci, co // 2, 1)
        lay.bn(p + ".bn2", co // 2); conv(p + ".conv2.conv", co // 2, co // 2, 3)
        lay.bn(p + ".bn3", co // 2); conv(p + ".conv3.conv", co // 2, co, 1)
        conv(p + ".skip_layer.conv", ci, co, 1)

    def hourglass(p, n, f):
        residual(p + ".up1", f, f); residual(p + ".low1", f, f)
        if n > 1:
            hourglass(p + ".low2", n - 1, f)
        else:
            residual(p + ".low2", f, f)
        residual(p + ".low3", f, f)

    conv("pre.0.conv", 1, 64, 5); lay.bn("pre.0.bn", 64)
    residual("pre.1", 64, 128); residual("pre.3", 128, 256); residual("pre.4", 256, 256)
    for i in range(nstack):
        hourglass(f"hgs.{i}.0", 4, 256)
        residual(f"features.{i}.0", 256, 256)
        conv(f"features.{i}.1.conv", 256, 256, 1); lay.bn(f"features.{i}.1.bn", 256)
        lay.head(f"outs.{i}", [(f"outs_1.{i}.weight", 3 * J), (f"outs_2.{i}.weight", J)], 256, [f"outs_1.{i}.bias", f"outs_2.{i}.bias"])
    for i in range(nstack - 1):
        conv(f"merge_features.{i}.conv.conv", 256, 256, 1)
        # merge_preds consumes the 4J-channel prediction volume, held as a zero-padded NHWC activation of align64(4J) channels
        lay.conv(f"merge_preds.{i}.conv.conv.weight", 256, 4 * J, 1, ci_pad=_align(4 * J))
        lay.vec(f"merge_preds.{i}.conv.conv.bias", 256)
    return lay


# ======================================================================================================
# activations
# ======================================================================================================
class Act:
    """NHWC activation (N,H,W,C) + lazily allocated gradient buffer."""

    def __init__(self, plan, N, H, W, C, dtype=None):
        self.plan, self.N, self.H, self.W, self.C = plan, N, H, W, C
        self.t = torch.empty(N, H, W, C, dtype=dtype or plan.tdtype, device=plan.device)
        self.g = None
        self.gw = False          # gradient already written in the backward plan (next contribution accumulates)
        self.stats = None        # awr_acc_t[2C] per-channel sum / sum-of-squares filled by the producing conv's epilogue (bf16 mode)

    @property
    def M(self):
        return self.N * self.H * self.W

    def grad(self):
        if self.g is None:
            self.g = torch.empty_like(self.t)
        return self.g


ACC = 4          # floats per order-independent accumulator (awr_acc_t = 2 x int64, include/awr_b200.h)


class BNState:
    def __init__(self, plan, prefix, C, sums=None):
        self.prefix, self.C = prefix, C
        self.fused_stats = sums is not None
        self.sums = sums if sums is not None else plan.arena(ACC * 2 * C)
        self.dsums = plan.arena(ACC * 2 * C)
        self.ss = torch.empty(2 * C, dtype=torch.float32, device=plan.device)
        self.mi = torch.empty(2 * C, dtype=torch.float32, device=plan.device)


# ======================================================================================================
# plan
# ======================================================================================================
class Plan:
    def __init__(self, net, J, downsample, B, H, precision, training, store, device):
        """store: ParamStore (flat params/grads/buffers). precision: 'fp32' | 'bf16'."""
        self.net, self.J, self.ds, self.B, self.H = net, J, downsample, B, H
        self.precision, self.training, self.store, self.device = precision, training, store, device
        self.tdtype = torch.float32 if precision == "fp32" else torch.bfloat16
        self.dt = L.F32 if precision == "fp32" else L.BF16
        # bf16 mode: convolutions on tcgen05 tensor cores (csrc/conv_tc.cu, wgrad_tc.cu).  AWR_B200_DEBUG_SIMT=1 keeps the bf16
        # activations but runs the CUDA-core conv kernels with fp32 weights -- a debugging aid for isolating rounding effects.
        self.tc = precision == "bf16" and os.environ.get("AWR_B200_DEBUG_SIMT") != "1"
        self.lib = L.lib()
        self.fwd, self.bwd = [], []
        self.fwd_meta, self.bwd_meta = [], []     # per launch: (kernel entry point, algorithmic flops, algorithmic bytes)
        self.fwd_side, self._fwd_branch, self._fwd_join = [], False, False      # per forward launch: 0 main, 1 branch stream, 2 main after joining it
        self.bwd_side = []                        # per backward launch: may run on the side stream (weight gradients: nothing downstream
                                                  # in the backward pass consumes them, so they overlap the dgrad / BN chain)
        self._arena_total = 0
        self.arena_buf = torch.zeros(1 << 22, dtype=torch.float32, device=device)    # zeroed at the start of every step
        self.ops = []
        self.launches_fwd = 0
        self.launches_bwd = 0
        self.img = torch.empty(B, 1, H, H, dtype=torch.float32, device=device)       # network input (NCHW == NHWC for C=1)
        kind, n = net.split("_")
        if kind == "resnet":
            self._build_resnet(int(n))
        elif kind == "hourglass":
            self._build_hourglass(int(n))
        else:
            raise ValueError(net)
        self.bwd_split = None          # first entry of bwd_splits (kept for callers that use a single early bucket)
        self.bwd_splits = []           # [(launch index, flat gradient offset)]: gradients at/after the offset are final after that many launches
        if training:
            marks = {}
            for i in range(len(self.ops) - 1, -1, -1):
                self.ops[i].plan_bwd()
                marks[i] = len(self.bwd)
            self._find_bwd_split(marks)

    def _op_param_names(self, op):
        names = []
        for attr in ("wname", "bname"):
            n = getattr(op, attr, None)
            if n:
                names.append(n)
        for attr in ("prefix", "res_prefix"):
            n = getattr(op, attr, None)
            if n:
                names += [n + ".weight", n + ".bias"]
        g = getattr(op, "gname", None)
        if g:
            names += [g + ".weight", g + ".bias"]
        return names

    def trained_param_names(self):
        """Canonical parameter names some op of this plan reads (and, in training, writes a gradient for).  The rest -- Hourglass
        skip_layer convs of equal-width Residual blocks (hourglass.py:38,45-48) -- have grad None in the reference."""
        lay = self.store.layout
        spans = []
        for op in self.ops:
            for n in self._op_param_names(op):
                if n in lay.groups:
                    off, shp = lay.groups[n]
                    spans.append((off, off + int(math.prod(shp))))
                elif n in lay.specs:
                    spans.append((lay.specs[n].offset, lay.specs[n].offset + lay.specs[n].numel))
        return [n for n in lay.order if any(a <= lay.specs[n].offset < b for a, b in spans)]

    def _find_bwd_split(self, marks):
        """Early gradient bucket for overlapping the data-parallel all-reduce with the rest of backward: the latest ops (head, deconvs,
        last stage) hold most parameters and finish their gradients first.  Picks the op suffix whose parameters are a contiguous tail
        of the flat buffer covering ~80 % of it."""
        lay = self.store.layout

        def span(name):
            if name in lay.groups:
                off, shp = lay.groups[name]
                return off, off + int(math.prod(shp))
            sp = lay.specs[name]
            return sp.offset, sp.offset + sp.numel
        total = lay.total
        cands = []                     # (launch count, offset, fraction of the buffer that is final), in backward order
        lo = total
        for i in range(len(self.ops) - 1, 0, -1):
            for n in self._op_param_names(self.ops[i]):
                lo = min(lo, span(n)[0])
            # every parameter at or beyond `lo` must belong to ops >= i
            ok = True
            for j in range(i):
                for n in self._op_param_names(self.ops[j]):
                    if span(n)[1] > lo:
                        ok = False
                        break
                if not ok:
                    break
            frac = (total - lo) / total
            if ok and 0 < marks[i] < len(self.bwd) and frac >= 0.5:
                cands.append((marks[i], lo, frac))
        # gradient buckets for the data-parallel all-reduce: the first closes when ~80 % of the bytes are final (its all-reduce overlaps
        # the rest of backward), then ~95 % and ~99 %, so that what is reduced AFTER backward is a fraction of a megabyte
        picked = []
        for target in (0.80, 0.95, 0.99):
            pool = [c for c in cands if not picked or (c[0] > picked[-1][0] and c[1] < picked[-1][1])]
            if not pool:
                break
            best = min(pool, key=lambda c: abs(c[2] - target))
            if picked and best[2] - picked[-1][2] < 0.02:
                continue
            picked.append(best)
        self.bwd_splits = [(c[0], c[1]) for c in picked]
        self.bwd_split = self.bwd_splits[0] if self.bwd_splits else None

    def bucket_range(self, part):
        """[a, b) of the flat gradient buffer that is final once backward part `part` has run."""
        offs = [self.store.grads.numel()] + [c[1] for c in self.bwd_splits] + [0]
        return offs[part + 1], offs[part]

    # ---- infrastructure ---------------------------------------------------------------------------
    def arena(self, n):
        """fp32 scratch that must be zero at the start of every step (BN sums etc.): carved from one buffer."""
        off = self._arena_total
        self._arena_total = _align(off + n)
        if self._arena_total > self.arena_buf.numel():
            raise RuntimeError("plan arena exhausted")
        return self.arena_buf[off: off + n]

    def arena_used(self):
        """The carved part of the arena (what a step has to re-zero)."""
        return self.arena_buf[: max(self._arena_total, 64)]

    def call(self, lst, fn, *args, flops=0, nbytes=0, tag=None, detail="", side=False):
        """Bind a C-ABI launch. Tensor-like args are resolved to pointers now (buffers are static).
        flops / nbytes: algorithmic work of this launch (for the roofline report); tag: kernel class label."""
        cargs = [a.data_ptr() if isinstance(a, torch.Tensor) else a for a in args]
        name = fn
        f = getattr(self.lib, fn)

        def run(stream):
            rc = f(*cargs, stream)
            if rc != 0:
                L.check(rc, name)
        lst.append(run)
        (self.fwd_meta if lst is self.fwd else self.bwd_meta).append((tag or fn, flops, nbytes, detail))
        if lst is self.bwd:
            self.bwd_side.append(bool(side))
        else:
            # forward: launches planned while a branch is open (Hourglass up1) may run on a second stream; a launch flagged `join`
            # (the add that consumes the branch) waits for it
            self.fwd_side.append(1 if self._fwd_branch else (2 if self._fwd_join else 0))
            self._fwd_join = False

    def P(self, name):
        return self.store.layout.phys(self.store.params, name)

    def G(self, name):
        return self.store.layout.phys(self.store.grads, name)

    def GW(self, name):
        """Weight-gradient target of a wgrad kernel: the fp32 gradient view, or (bit-reproducible build) the address of the same
        element range in the accumulator-slot buffer."""
        acc = self.store.grads_acc
        if acc is None:
            return self.G(name)
        lay = self.store.layout
        off = lay.groups[name][0] if name in lay.groups else lay.specs[name].offset
        return acc.data_ptr() + 16 * off

    def W16(self, name):
        return self.store.layout.phys(self.store.shadow, name)

    def buf(self, name):
        return self.store.buffers[name]

    def run_forward(self, stream=None, side=None):
        """side: optional torch.cuda.Stream for independent branches (Hourglass: each level's full-resolution `up1` Residual runs there
        while the low-resolution spine -- a chain of small, latency-bound launches -- continues on the main stream; the level's
        upsample+add joins them).  Without it everything runs in program order on one stream."""
        s = L.stream() if stream is None else stream
        if side is None or not any(self.fwd_side):
            for f in self.fwd:
                f(s)
            return
        main = torch.cuda.current_stream()
        on_side = False
        for f, where in zip(self.fwd, self.fwd_side):
            if where == 1:
                if not on_side:                  # a branch opens: its input is whatever the main stream has produced so far
                    side.wait_stream(main)
                    on_side = True
                f(side.cuda_stream)
            else:
                on_side = False
                if where == 2:
                    main.wait_stream(side)
                f(s)
        main.wait_stream(side)

    def run_backward(self, stream=None, side=None, part=None, join=True):
        """side: optional torch.cuda.Stream.  Launches flagged `side` (weight gradients) are issued there, each after everything
        enqueued so far on the main stream (its inputs), and the main stream joins the side stream at the end (join=False: the caller
        orders whatever consumes this part's weight gradients after `side` itself, so the dgrad chain does not stall on them).  Works
        eagerly and under CUDA-graph capture (the fork/join events become graph edges)."""
        s = L.stream() if stream is None else stream
        lo, hi = 0, len(self.bwd)
        if part is not None and self.bwd_splits:                   # part i: launches between split i-1 and split i (0 .. len(splits))
            cuts = [0] + [c[0] for c in self.bwd_splits] + [len(self.bwd)]
            lo, hi = cuts[part], cuts[part + 1]
        if side is None:
            for f in self.bwd[lo:hi]:
                f(s)
        else:
            main = torch.cuda.current_stream()
            for f, on_side in list(zip(self.bwd, self.bwd_side))[lo:hi]:
                if on_side:
                    side.wait_stream(main)
                    f(side.cuda_stream)
                else:
                    f(s)
            if join or self.store.grads_acc is not None:
                main.wait_stream(side)
        acc = self.store.grads_acc
        if acc is not None:
            # bit-reproducible build: fold the accumulator slots of the gradients that are final now into the fp32 buffer
            n = self.store.grads.numel()
            a, b = 0, n
            if part is not None and self.bwd_splits:
                a, b = self.bucket_range(part)
            L.check(self.lib.awr_grad_acc_finalize(acc.data_ptr() + 16 * a, self.store.grads.data_ptr() + 4 * a, b - a, s), "awr_grad_acc_finalize")

    # ---- ops ------------------------------------------------------------------------------------------
    def stem(self, wname, bname, Cout, k):
        op = _Stem(self, wname, bname, Cout, k)
        self.ops.append(op)
        return op.y

    def conv(self, x, wname, bname, Cout, k, stride, pad, bn_next=True):
        """bn_next: the output feeds a BatchNorm, so (tensor-core path, training) the conv epilogue also produces its batch statistics."""
        op = _Conv(self, x, wname, bname, Cout, k, stride, pad, transposed=False, want_stats=bn_next)
        self.ops.append(op)
        return op.y

    def deconv(self, x, wname, Cout, k=4, stride=2, pad=1):
        op = _Conv(self, x, wname, None, Cout, k, stride, pad, transposed=True, want_stats=True)
        self.ops.append(op)
        return op.y

    def bn_act(self, y, prefix, relu, res=None, res_y=None, res_prefix=None):
        op = _BNAct(self, y, prefix, relu, res, res_y, res_prefix)
        self.ops.append(op)
        return op.out

    def bn_pool(self, y, prefix, k, s, p):
        op = _BNPool(self, y, prefix, k, s, p)
        self.ops.append(op)
        return op.out

    def add(self, a, b, c=None):
        op = _Add(self, a, b, c)
        self.ops.append(op)
        return op.out

    def maxpool(self, x, k, s, p):
        op = _MaxPool(self, x, k, s, p)
        self.ops.append(op)
        return op.out

    def upsample_add(self, up, low):
        op = _UpAdd(self, up, low)
        self.ops.append(op)
        return op.out

    def head(self, x, gname):
        op = _Head(self, x, gname)
        self.ops.append(op)
        return op

    # ---- networks -------------------------------------------------------------------------------------
    def _build_resnet(self, layers):
        kind, counts = RESNET_SPEC[layers]
        exp = 1 if kind == "basic" else 4
        y0 = self.stem("pre.0.weight", None, 64, 5)
        c = self.bn_pool(y0, "pre.1", 3, 2, 1)           # BN + ReLU + MaxPool(3,2,1) fused: the 128x128x64 normalised tensor is never stored
        inpl = 64
        for li, (planes, nblk) in enumerate(zip([64, 128, 256, 512], counts), start=1):
            for bi in range(nblk):
                p = f"layer{li}.{bi}"
                stride = 2 if (li > 1 and bi == 0) else 1
                has_ds = bi == 0 and (stride != 1 or inpl != planes * exp)
                if kind == "basic":
                    o = self.bn_act(self.conv(c, p + ".conv1.weight", None, planes, 3, stride, 1), p + ".bn1", True)
                    y_last = self.conv(o, p + ".conv2.weight", None, planes, 3, 1, 1)
                    last_bn = p + ".bn2"
                else:
                    o = self.bn_act(self.conv(c, p + ".conv1.weight", None, planes, 1, 1, 0), p + ".bn1", True)
                    o = self.bn_act(self.conv(o, p + ".conv2.weight", None, planes, 3, stride, 1), p + ".bn2", True)
                    y_last = self.conv(o, p + ".conv3.weight", None, planes * 4, 1, 1, 0)
                    last_bn = p + ".bn3"
                if has_ds:
                    yd = self.conv(c, p + ".downsample.0.weight", None, planes * exp, 1, stride, 0)
                    c = self.bn_act(y_last, last_bn, True, res_y=yd, res_prefix=p + ".downsample.1")
                else:
                    c = self.bn_act(y_last, last_bn, True, res=c)
                inpl = planes * exp
        for i in range(4 - int(math.log(self.ds, 2))):
            c = self.bn_act(self.deconv(c, f"deconv_layers.{3 * i}.weight", 256), f"deconv_layers.{3 * i + 1}", True)
        self.heads = [self.head(c, "final")]

    def _hg_conv(self, x, p, co, k, bn_next=True):
        return self.conv(x, p + ".conv.weight", p + ".conv.bias", co, k, 1, (k - 1) // 2, bn_next=bn_next)

    def _residual(self, x, p, co):
        ci = x.C
        o = self._hg_conv(self.bn_act(x, p + ".bn1", True), p + ".conv1", co // 2, 1)
        o = self._hg_conv(self.bn_act(o, p + ".bn2", True), p + ".conv2", co // 2, 3)
        o = self._hg_conv(self.bn_act(o, p + ".bn3", True), p + ".conv3", co, 1, bn_next=False)
        res = self._hg_conv(x, p + ".skip_layer", co, 1, bn_next=False) if ci != co else x
        return self.add(o, res)

    def _hourglass(self, x, p, n):
        self._fwd_branch = True                  # hourglass.py:79-88: up1 depends only on x, the pool -> low1 -> low2 -> low3 spine likewise
        up1 = self._residual(x, p + ".up1", x.C)
        self._fwd_branch = False
        low1 = self._residual(self.maxpool(x, 2, 2, 0), p + ".low1", x.C)
        low2 = self._hourglass(low1, p + ".low2", n - 1) if n > 1 else self._residual(low1, p + ".low2", x.C)
        low3 = self._residual(low2, p + ".low3", x.C)
        self._fwd_join = True
        return self.upsample_add(up1, low3)

    def _build_hourglass(self, nstack):
        y0 = self.stem("pre.0.conv.weight", "pre.0.conv.bias", 64, 5)
        c = self.bn_act(y0, "pre.0.bn", True)
        c = self._residual(c, "pre.1", 128)
        c = self.maxpool(c, 2, 2, 0)
        c = self._residual(c, "pre.3", 256)
        c = self._residual(c, "pre.4", 256)
        self.heads = []
        for i in range(nstack):
            hg = self._hourglass(c, f"hgs.{i}.0", 4)
            f = self._residual(hg, f"features.{i}.0", 256)
            f = self.bn_act(self._hg_conv(f, f"features.{i}.1", 256, 1), f"features.{i}.1.bn", True)
            hd = self.head(f, f"outs.{i}")
            self.heads.append(hd)
            if i < nstack - 1:
                mp = self.conv(hd.pred_nhwc(), f"merge_preds.{i}.conv.conv.weight", f"merge_preds.{i}.conv.conv.bias", 256, 1, 1, 0, bn_next=False)
                mf = self._hg_conv(f, f"merge_features.{i}.conv", 256, 1, bn_next=False)
                c = self.add(c, mp, mf)


def _contribute(plan, act, emit):
    """Route one gradient contribution into act.grad(): emit(dst_tensor, accumulate: bool)."""
    g = act.grad()
    emit(g, act.gw)
    act.gw = True


class _Op:
    def plan_bwd(self):
        raise NotImplementedError


class _Stem(_Op):
    """1-channel k x k stem conv (resnet_deconv.py:32 / hourglass.py:112)."""

    def __init__(self, plan, wname, bname, Cout, k):
        self.plan, self.wname, self.bname, self.k = plan, wname, bname, k
        B, H = plan.B, plan.H
        self.y = Act(plan, B, H, H, Cout)
        w = plan.P(wname)        # physical [k][k][Cout][1] == [k*k][Cout]
        b = plan.P(bname) if bname else None
        if plan.training and k == 5:
            self.y.stats = plan.arena(ACC * 2 * Cout)    # BatchNorm statistics come out of the conv kernel itself
        plan.call(plan.fwd, "awr_stem_conv", plan.img, w, b, self.y.t, self.y.stats, plan.dt, B, H, H, Cout, k)

    def plan_bwd(self):
        pl = self.plan
        if not self.y.gw:
            return
        pl.call(pl.bwd, "awr_stem_wgrad", pl.img, self.y.grad(), pl.GW(self.wname), pl.GW(self.bname) if self.bname else None,
                pl.dt, pl.B, pl.H, pl.H, self.y.C, self.k)


class _Conv(_Op):
    """Conv2d / ConvTranspose2d on NHWC activations; output is the raw (pre-BN) tensor."""

    def __init__(self, plan, x, wname, bname, Cout, k, stride, pad, transposed, want_stats=False):
        self.plan, self.x, self.wname, self.bname = plan, x, wname, bname
        self.k, self.stride, self.pad, self.transposed, self.Cout = k, stride, pad, transposed, Cout
        self.Cin = x.C
        if transposed:
            Ho, Wo = (x.H - 1) * stride - 2 * pad + k, (x.W - 1) * stride - 2 * pad + k
        else:
            Ho, Wo = (x.H + 2 * pad - k) // stride + 1, (x.W + 2 * pad - k) // stride + 1
        self.y = Act(plan, x.N, Ho, Wo, Cout)
        self.want_stats = None                   # BNState set by the following bn_act (stats over this output)
        coarse = x.M if transposed else self.y.M   # algorithmic GEMM work: 2 * Cin*Cout*k*k per pixel of the coarse side
        self.flops = 2 * coarse * self.Cin * Cout * k * k
        self.detail = f"{'deconv' if transposed else 'conv'}{k}x{k}s{stride} {self.Cin}->{Cout} @{x.H}->{Ho} N{x.N}"
        if plan.tc:
            self._check_tc_geometry()
        if plan.tc and plan.training and want_stats:
            self.y.stats = plan.arena(ACC * 2 * Cout)
        self._emit_fwd()

    def _check_tc_geometry(self):
        """The tcgen05 kernels' envelope (csrc/conv_tc.cu, conv_halo_tc.cu, wgrad_tc.cu host checks), verified here so that an unsupported
        configuration fails at plan construction with the layer named instead of as `invalid argument (-1)` from a launch."""
        x, y = self.x, self.y
        c = x if self.transposed else y                      # the coarse side tiles the GEMM's pixel dimension
        pow2 = lambda v: v > 0 and (v & (v - 1)) == 0
        why = None
        if self.Cin % 64 or self.Cout % 64:
            why = f"channel counts must be multiples of 64 (got {self.Cin} -> {self.Cout})"
        elif not (pow2(c.H) and pow2(c.W) and c.H <= 256 and c.W <= 256):
            why = f"feature maps must have power-of-two sides <= 256 (got {c.H}x{c.W}: img_size must be a power of two <= 512)"
        elif self.stride not in (1, 2):
            why = f"stride {self.stride}"
        else:
            cg = self.Cout if self.transposed else self.Cin   # channels of the gathered operand of the weight-gradient GEMM
            if self.plan.training and not (cg == 64 or cg % 128 == 0):
                why = f"weight-gradient kernel needs 64 or a multiple of 128 gathered channels (got {cg})"
        if why:
            raise ValueError(f"precision='bf16' (tensor-core path) does not support layer `{self.wname}` [{self.detail}]: {why}. "
                             "Use precision='fp32' (CUDA-core kernels, any geometry) for this configuration.")

    def _emit_fwd(self):
        pl, x, y = self.plan, self.x, self.y
        b = pl.P(self.bname) if self.bname else None
        if pl.tc:
            pl.call(pl.fwd, "awr_conv_tc", x.t, pl.W16(self.wname), b, y.t, y.stats, x.N, x.H, x.W, self.Cin, y.H, y.W, self.Cout, self.k, self.k,
                    self.stride, self.pad, int(self.transposed), 1, self.Cin, self.Cout * self.Cin, 0, 0, 0, flops=self.flops, tag="conv_fprop", detail=self.detail)
            return
        pl.call(pl.fwd, "awr_conv_simt", x.t, pl.P(self.wname), b, y.t, pl.dt, x.N, x.H, x.W, self.Cin, y.H, y.W, self.Cout, self.k, self.k,
                self.stride, self.pad, int(self.transposed), 1, self.Cin, self.Cout * self.Cin, 0, 0, 0, flops=self.flops, tag="conv_fprop", detail=self.detail)

    def plan_bwd(self):
        pl, x, y = self.plan, self.x, self.y
        if not y.gw:
            return
        dy = y.grad()
        gW = pl.GW(self.wname)
        if pl.tc:
            if not self.transposed:
                pl.call(pl.bwd, "awr_conv_wgrad_tc", dy, x.t, gW, x.N, y.H, y.W, self.Cout, x.H, x.W, self.Cin, self.k, self.k, self.stride,
                        self.pad, self.Cin, 1, self.Cout * self.Cin, flops=self.flops, tag="conv_wgrad", detail=self.detail, side=True)
            else:
                pl.call(pl.bwd, "awr_conv_wgrad_tc", x.t, dy, gW, x.N, x.H, x.W, self.Cin, y.H, y.W, self.Cout, self.k, self.k, self.stride,
                        self.pad, 1, self.Cin, self.Cout * self.Cin, flops=self.flops, tag="conv_wgrad", detail=self.detail, side=True)
            if self.bname:
                pl.call(pl.bwd, "awr_channel_stats", dy, pl.dt, y.M, self.Cout, pl.arena(ACC * self.Cout), 0, pl.G(self.bname), pl.arena(1),
                        side=True)          # bias gradient: like the weight gradient, nothing downstream in backward consumes it
            w16 = pl.W16(self.wname)

            def emit_tc(dst, acc):
                pl.call(pl.bwd, "awr_conv_tc", dy, w16, None, dst, None, y.N, y.H, y.W, self.Cout, x.H, x.W, self.Cin, self.k, self.k, self.stride,
                        self.pad, int(not self.transposed), self.Cin, 1, self.Cout * self.Cin, 0, 0, int(acc), flops=self.flops, tag="conv_dgrad", detail=self.detail)
            _contribute(pl, x, emit_tc)
            return
        # weight gradient
        if not self.transposed:
            pl.call(pl.bwd, "awr_conv_wgrad_simt", dy, x.t, gW, pl.dt, x.N, y.H, y.W, self.Cout, x.H, x.W, self.Cin, self.k, self.k,
                    self.stride, self.pad, self.Cin, 1, self.Cout * self.Cin, flops=self.flops, tag="conv_wgrad", detail=self.detail, side=True)
        else:
            pl.call(pl.bwd, "awr_conv_wgrad_simt", x.t, dy, gW, pl.dt, x.N, x.H, x.W, self.Cin, y.H, y.W, self.Cout, self.k, self.k,
                    self.stride, self.pad, 1, self.Cin, self.Cout * self.Cin, flops=self.flops, tag="conv_wgrad", detail=self.detail, side=True)
        if self.bname:
            pl.call(pl.bwd, "awr_channel_stats", dy, pl.dt, y.M, self.Cout, pl.arena(ACC * self.Cout), 0, pl.G(self.bname), pl.arena(1),
                        side=True)          # bias gradient: like the weight gradient, nothing downstream in backward consumes it
        # data gradient
        if True:
            w = pl.P(self.wname)

            def emit(dst, acc):
                pl.call(pl.bwd, "awr_conv_simt", dy, w, None, dst, pl.dt, y.N, y.H, y.W, self.Cout, x.H, x.W, self.Cin, self.k, self.k,
                        self.stride, self.pad, int(not self.transposed), self.Cin, 1, self.Cout * self.Cin, 0, 0, int(acc),
                        flops=self.flops, tag="conv_dgrad", detail=self.detail)
            _contribute(pl, x, emit)


class _BNAct(_Op):
    """out = act( BN(y) [+ res | + BN_res(res_y)] )   -- BatchNorm2d(+ReLU)(+residual) of both backbones."""

    def __init__(self, plan, y, prefix, relu, res, res_y, res_prefix):
        self.plan, self.y, self.prefix, self.relu, self.res, self.res_y, self.res_prefix = plan, y, prefix, relu, res, res_y, res_prefix
        self.bn = BNState(plan, prefix, y.C, y.stats)
        self.bn_res = BNState(plan, res_prefix, y.C, res_y.stats) if res_y is not None else None
        self.out = Act(plan, y.N, y.H, y.W, y.C)
        pl = plan
        tr = pl.training
        for t, bn in ((y, self.bn), (res_y, self.bn_res)):
            if bn is not None and tr and not bn.fused_stats:
                pl.call(pl.fwd, "awr_channel_stats", t.t, pl.dt, t.M, t.C, bn.sums, 1, None, None)

        def bnset(bn):
            if bn is None:
                return [None] * 7
            p = bn.prefix
            return [bn.sums if tr else None, pl.P(p + ".weight"), pl.P(p + ".bias"), pl.buf(p + ".running_mean"), pl.buf(p + ".running_var"),
                    pl.buf(p + ".num_batches_tracked") if tr else None, bn.mi]
        a, b = bnset(self.bn), bnset(self.bn_res)
        rt = res.t if res is not None else (res_y.t if res_y is not None else None)
        self.detail = f"bn {y.C}ch @{y.H} N{y.N}{' +res' if res is not None else ''}{' +resbn' if res_y is not None else ''}"
        pl.call(pl.fwd, "awr_bn_act", y.t, *a, rt, *b, self.out.t, pl.dt, y.M, y.C, BN_MOMENTUM, BN_EPS, int(tr), int(relu), detail=self.detail)

    def plan_bwd(self):
        pl, y, out = self.plan, self.y, self.out
        if not out.gw:
            return
        dout = out.grad()
        act = out.t if self.relu else None
        p = self.prefix
        # ReLU(BN(y)) without a residual: the mask is a function of y, so the backward never reads the activation tensor
        remask = self.relu and self.res is None and self.res_y is None
        mg, mb = (pl.P(p + ".weight"), pl.P(p + ".bias")) if remask else (None, None)
        if remask:
            act = None
        dy = y.grad()
        dy_add = dy if y.gw else None
        dres = dres_add = None
        if self.res is not None:
            dres = self.res.grad()
            dres_add = dres if self.res.gw else None
            self.res.gw = True
        # small tensors (operands fit in one CTA per SM): ONE launch -- bulk-TMA staging in shared memory, grid barrier, dy from the same copy
        if pl.lib.awr_bn_bwd_fused_ok(y.M, y.C, pl.dt, int(act is not None)):
            pl.call(pl.bwd, "awr_bn_bwd_fused", dout, act, y.t, self.bn.mi, pl.P(p + ".weight"), mb, self.bn.dsums, pl.arena(1), dy, dy_add,
                    dres, dres_add, pl.G(p + ".weight"), pl.G(p + ".bias"), pl.dt, y.M, y.C, 1, tag="awr_bn_bwd_fused", detail=self.detail)
        else:
            pl.call(pl.bwd, "awr_bn_bwd_reduce", dout, act, y.t, self.bn.mi, mg, mb, pl.dt, y.M, y.C, self.bn.dsums, detail=self.detail)
            pl.call(pl.bwd, "awr_bn_bwd_apply", dout, act, y.t, self.bn.mi, self.bn.dsums, pl.P(p + ".weight"), dy, dy_add, dres, dres_add,
                    pl.G(p + ".weight"), pl.G(p + ".bias"), mb, pl.dt, y.M, y.C, 1, detail=self.detail)
        y.gw = True
        if self.res_y is not None:
            ry, rp = self.res_y, self.res_prefix
            dry = ry.grad()
            if pl.lib.awr_bn_bwd_fused_ok(ry.M, ry.C, pl.dt, int(act is not None)):
                pl.call(pl.bwd, "awr_bn_bwd_fused", dout, act, ry.t, self.bn_res.mi, pl.P(rp + ".weight"), None, self.bn_res.dsums, pl.arena(1), dry,
                        dry if ry.gw else None, None, None, pl.G(rp + ".weight"), pl.G(rp + ".bias"), pl.dt, ry.M, ry.C, 1,
                        tag="awr_bn_bwd_fused")
            else:
                pl.call(pl.bwd, "awr_bn_bwd_reduce", dout, act, ry.t, self.bn_res.mi, None, None, pl.dt, ry.M, ry.C, self.bn_res.dsums)
                pl.call(pl.bwd, "awr_bn_bwd_apply", dout, act, ry.t, self.bn_res.mi, self.bn_res.dsums, pl.P(rp + ".weight"), dry,
                        dry if ry.gw else None, None, None, pl.G(rp + ".weight"), pl.G(rp + ".bias"), None, pl.dt, ry.M, ry.C, 1)
            ry.gw = True


class _BNPool(_Op):
    """out = MaxPool(ReLU(BN(y)))  (resnet_deconv.py:33-35) without materialising the normalised tensor or its gradient."""

    def __init__(self, plan, y, prefix, k, s, p):
        self.plan, self.y, self.prefix, self.k, self.s, self.p = plan, y, prefix, k, s, p
        pl, tr = plan, plan.training
        self.bn = BNState(plan, prefix, y.C, y.stats)
        Ho, Wo = (y.H + 2 * p - k) // s + 1, (y.W + 2 * p - k) // s + 1
        self.out = Act(plan, y.N, Ho, Wo, y.C)
        self.idx = torch.empty(y.N, Ho, Wo, y.C, dtype=torch.uint8, device=plan.device) if tr else None
        if tr and not self.bn.fused_stats:
            pl.call(pl.fwd, "awr_channel_stats", y.t, pl.dt, y.M, y.C, self.bn.sums, 1, None, None)
        pl.call(pl.fwd, "awr_bn_relu_maxpool_fwd", y.t, self.bn.sums if tr else None, pl.P(prefix + ".weight"), pl.P(prefix + ".bias"),
                pl.buf(prefix + ".running_mean"), pl.buf(prefix + ".running_var"), pl.buf(prefix + ".num_batches_tracked") if tr else None,
                self.bn.mi, self.out.t, self.idx, pl.dt, y.N, y.H, y.W, y.C, k, s, p, BN_MOMENTUM, BN_EPS, int(tr))

    def plan_bwd(self):
        pl, y, out, pf = self.plan, self.y, self.out, self.prefix
        if not out.gw:
            return
        assert not y.gw
        # pass 0 (sum dz, sum dz*yhat) from the two POOLED tensors alone -- see pool_bn_bwd_reduce_kernel; AWR_B200_POOL_PASS0=full keeps the
        # full-resolution reduction
        if os.environ.get("AWR_B200_POOL_PASS0") == "full":
            pl.call(pl.bwd, "awr_maxpool_bn_bwd", out.grad(), self.idx, y.t, self.bn.mi, pl.P(pf + ".weight"), pl.P(pf + ".bias"), self.bn.dsums,
                    y.grad(), pl.G(pf + ".weight"), pl.G(pf + ".bias"), pl.dt, y.N, y.H, y.W, y.C, self.k, self.s, self.p, 0, 1)
        else:
            pl.call(pl.bwd, "awr_pool_bn_bwd_reduce", out.grad(), out.t, pl.P(pf + ".weight"), pl.P(pf + ".bias"), self.bn.dsums, pl.dt, out.M, out.C)
        pl.call(pl.bwd, "awr_maxpool_bn_bwd", out.grad(), self.idx, y.t, self.bn.mi, pl.P(pf + ".weight"), pl.P(pf + ".bias"), self.bn.dsums,
                y.grad(), pl.G(pf + ".weight"), pl.G(pf + ".bias"), pl.dt, y.N, y.H, y.W, y.C, self.k, self.s, self.p, 1, 1)
        y.gw = True


class _Add(_Op):
    """out = a + b (+ c)   (hourglass.py:58, :163)"""

    def __init__(self, plan, a, b, c=None):
        self.plan, self.ins = plan, [a, b] + ([c] if c is not None else [])
        self.out = Act(plan, a.N, a.H, a.W, a.C)
        pl = plan
        pl.call(pl.fwd, "awr_affine_act", a.t, None, b.t, None, self.out.t, pl.dt, a.M, a.C, 0)
        if c is not None:
            pl.call(pl.fwd, "awr_affine_act", self.out.t, None, c.t, None, self.out.t, pl.dt, a.M, a.C, 0)

    def plan_bwd(self):
        pl, out = self.plan, self.out
        if not out.gw:
            return
        dout = out.grad()
        n = out.M * out.C
        for x in self.ins:
            def emit(dst, acc, x=x):
                pl.call(pl.bwd, "awr_relu_bwd", dout, None, dst if acc else None, dst, pl.dt, n)
            _contribute(pl, x, emit)


class _MaxPool(_Op):
    def __init__(self, plan, x, k, s, p):
        self.plan, self.x, self.k, self.s, self.p = plan, x, k, s, p
        Ho, Wo = (x.H + 2 * p - k) // s + 1, (x.W + 2 * p - k) // s + 1
        self.out = Act(plan, x.N, Ho, Wo, x.C)
        self.idx = torch.empty(x.N, Ho, Wo, x.C, dtype=torch.uint8, device=plan.device) if plan.training else None
        plan.call(plan.fwd, "awr_maxpool_fwd", x.t, self.out.t, self.idx, plan.dt, x.N, x.H, x.W, x.C, k, s, p)

    def plan_bwd(self):
        pl, x, out = self.plan, self.x, self.out
        if not out.gw:
            return

        def emit(dst, acc):
            pl.call(pl.bwd, "awr_maxpool_bwd", out.grad(), self.idx, dst, pl.dt, x.N, x.H, x.W, x.C, self.k, self.s, self.p, int(acc))
        _contribute(pl, x, emit)


class _UpAdd(_Op):
    """out = up + nearest_x2(low)   (hourglass.py:87-88)"""

    def __init__(self, plan, up, low):
        self.plan, self.up, self.low = plan, up, low
        self.out = Act(plan, up.N, up.H, up.W, up.C)
        plan.call(plan.fwd, "awr_upsample2_add", up.t, low.t, self.out.t, plan.dt, up.N, up.H, up.W, up.C)

    def plan_bwd(self):
        pl, out = self.plan, self.out
        if not out.gw:
            return
        dout = out.grad()
        n = out.M * out.C

        def emit_up(dst, acc):
            pl.call(pl.bwd, "awr_relu_bwd", dout, None, dst if acc else None, dst, pl.dt, n)
        _contribute(pl, self.up, emit_up)

        def emit_low(dst, acc):
            pl.call(pl.bwd, "awr_upsample2_bwd", dout, dst, pl.dt, out.N, out.H, out.W, out.C, int(acc))
        _contribute(pl, self.low, emit_low)


class _Head(_Op):
    """final1 || final2 (resnet_deconv.py:52-53,133-136) / outs_1 || outs_2 (hourglass.py:135-136,153-157):
    one 1x1 GEMM with N = 4J (padded to a multiple of 64) writing the (B,4J,F,F) fp32 NCHW volume the AWR head consumes."""

    def __init__(self, plan, x, gname):
        self.plan, self.x, self.gname = plan, x, gname
        J = plan.J
        HC = self.HC = _align(4 * J)
        self.pred = torch.empty(x.N, 4 * J, x.H, x.W, dtype=torch.float32, device=plan.device)
        self.dpred = torch.zeros_like(self.pred) if plan.training else None   # written by the head/loss backward (or autograd)
        self._nhwc = None
        pl = plan
        if pl.tc:
            pl.call(pl.fwd, "awr_conv_tc", x.t, pl.W16(gname + ".weight"), pl.P(gname + ".bias"), self.pred, None, x.N, x.H, x.W, x.C,
                    x.H, x.W, HC, 1, 1, 1, 0, 0, 1, x.C, HC * x.C, 1, 4 * J, 0, flops=2 * x.M * x.C * 4 * J, tag="conv_fprop")
        else:
            pl.call(pl.fwd, "awr_conv_simt", x.t, pl.P(gname + ".weight"), pl.P(gname + ".bias"), self.pred, pl.dt, x.N, x.H, x.W, x.C,
                    x.H, x.W, HC, 1, 1, 1, 0, 0, 1, x.C, HC * x.C, 1, 4 * J, 0, flops=2 * x.M * x.C * 4 * J, tag="conv_fprop")

    def pred_nhwc(self):
        """Prediction volume as a zero-padded NHWC activation of align64(4J) channels (input of merge_preds in stacked hourglasses)."""
        if self._nhwc is None:
            pl, x = self.plan, self.x
            self._nhwc = Act(pl, x.N, x.H, x.W, self.HC)
            pl.call(pl.fwd, "awr_nchw_to_nhwc", self.pred, self._nhwc.t, pl.dt, x.N, 4 * pl.J, self.HC, x.H * x.W)
        return self._nhwc

    def plan_bwd(self):
        pl, x = self.plan, self.x
        J, HC = pl.J, self.HC
        P = x.H * x.W
        has_direct = self.dpred is not None
        has_nhwc = self._nhwc is not None and self._nhwc.gw
        if not (has_direct or has_nhwc):
            return
        if self._nhwc is None:
            self._nhwc = Act(pl, x.N, x.H, x.W, HC)
        d = self._nhwc.grad()
        if has_direct:
            if has_nhwc:
                tmp = torch.empty_like(d)
                pl.call(pl.bwd, "awr_nchw_to_nhwc", self.dpred, tmp, pl.dt, x.N, 4 * J, HC, P)
                pl.call(pl.bwd, "awr_relu_bwd", tmp, None, d, d, pl.dt, d.numel())
            else:
                pl.call(pl.bwd, "awr_nchw_to_nhwc", self.dpred, d, pl.dt, x.N, 4 * J, HC, P)
        g = self.gname
        if pl.tc:
            pl.call(pl.bwd, "awr_conv_wgrad_tc", d, x.t, pl.GW(g + ".weight"), x.N, x.H, x.W, HC, x.H, x.W, x.C, 1, 1, 1, 0, x.C, 1, HC * x.C,
                    flops=2 * x.M * x.C * 4 * J, tag="conv_wgrad")
            pl.call(pl.bwd, "awr_channel_stats", d, pl.dt, x.M, HC, pl.arena(ACC * HC), 0, pl.G(g + ".bias"), pl.arena(1))

            def emit_tc(dst, acc):
                pl.call(pl.bwd, "awr_conv_tc", d, pl.W16(g + ".weight"), None, dst, None, x.N, x.H, x.W, HC, x.H, x.W, x.C, 1, 1, 1, 0, 0,
                        x.C, 1, HC * x.C, 0, 0, int(acc), flops=2 * x.M * x.C * 4 * J, tag="conv_dgrad")
            _contribute(pl, x, emit_tc)
            return
        pl.call(pl.bwd, "awr_conv_wgrad_simt", d, x.t, pl.GW(g + ".weight"), pl.dt, x.N, x.H, x.W, HC, x.H, x.W, x.C, 1, 1, 1, 0, x.C, 1,
                HC * x.C, flops=2 * x.M * x.C * 4 * J, tag="conv_wgrad")
        pl.call(pl.bwd, "awr_channel_stats", d, pl.dt, x.M, HC, pl.arena(ACC * HC), 0, pl.G(g + ".bias"), pl.arena(1))

        def emit(dst, acc):
            pl.call(pl.bwd, "awr_conv_simt", d, pl.P(g + ".weight"), None, dst, pl.dt, x.N, x.H, x.W, HC, x.H, x.W, x.C, 1, 1, 1, 0, 0,
                    x.C, 1, HC * x.C, 0, 0, int(acc), flops=2 * x.M * x.C * 4 * J, tag="conv_dgrad")
        _contribute(pl, x, emit)
