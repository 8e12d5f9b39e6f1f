"""Fused training step of the AWR hot path (the loop body of the reference's train.py:107-131) on libawr_b200.so.

    H2D(img, jt_uvd_gt) -> backbone fwd -> fused head+loss fwd -> fused head+loss bwd -> backbone bwd
        -> [NCCL all-reduce of the flat gradient buffer] -> fused Adam

The step is a fixed launch sequence over static buffers (engine.Plan), captured once into CUDA graphs:
`graph_fb` (zero scratch, forward, head, backward) and `graph_opt` (Adam); the data-parallel all-reduce
runs between them on the same stream through torch.distributed (NCCL over NVLink).  GT volumes, the
coordinate grid and the loss temporaries of the reference are never materialised (csrc/head.cu).

Semantics kept from train.py: loss = coord_weight*SmoothL1(uvd, jt) + dense_weight*SmoothL1(pred, joint2offset(jt))
(:119-120); for 'hourglass_N' only the last stack is supervised (:116-121 overwrite `loss`); Adam(lr, betas
(0.9,0.999), eps 1e-8, weight_decay) (:67) or SGD(momentum 0.9) (:69); parameters whose gradient is None in the reference
(Hourglass skip_layer convs that forward never calls) are not stepped, as torch.optim skips them; BN statistics are per
replica (the reference has no SyncBN).  The learning rate lives in device memory, so StepLR / ReduceLROnPlateau
(awr_b200.optim, train.py:89-92,157-160) drive the captured graph without re-capture.
"""
import os

import torch

from . import _lib as L
from . import dp
from .modules import AWRBackbone


def map_optimizer_state(sd, pnames):
    """torch.optim state_dict -> {canonical parameter name: per-parameter state}.  State keys are whatever `param_groups[*]['params']`
    lists, in parameter order: positions in current torch, `id(param)` values in the torch<=1.1 pickles the reference ships
    (results/hourglass_1.pth: 250 parameters, 220 state entries -- the skip_layer convs forward never calls have none)."""
    keys = [k for g in sd["param_groups"] for k in g["params"]]
    if len(keys) != len(pnames):
        raise ValueError(f"optimizer state lists {len(keys)} parameters, the module has {len(pnames)}")
    return {name: sd["state"][key] for key, name in zip(keys, pnames) if key in sd["state"]}


class FusedTrainer:
    def __init__(self, module: AWRBackbone, batch_size, img_size, kernel_size, coord_weight=1.0, dense_weight=1.0, lr=1e-3,
                 betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, world_size=1, use_graph=True, process_group=None, all_stacks=False,
                 optimizer="adam", momentum=0.9, keep_grads=False):
        """all_stacks (hourglass_N, N > 1): supervise every stack and sum the per-stack losses (test.py:74-80); the default keeps
        train.py:116-121's behaviour, where only the last stack's loss survives the loop.
        optimizer: 'adam' (train.py:67) | 'sgd' (momentum `momentum`, train.py:69).
        keep_grads: leave the step's gradients in store.grads (tests, inspection); by default the optimizer kernel zero-fills the
        buffer once it has consumed it, which replaces a separate 4*n-byte fill per step."""
        if not isinstance(module, AWRBackbone):
            raise TypeError("FusedTrainer drives awr_b200 backbones (get_deconv_net / PoseNet)")
        self.module = module
        module.train()
        self.B, self.H = int(batch_size), int(img_size)
        self.ks, self.cw, self.dw = float(kernel_size), float(coord_weight), float(dense_weight)
        self.lr, self.betas, self.eps, self.wd = float(lr), betas, float(eps), float(weight_decay)
        if optimizer not in ("adam", "sgd"):
            raise ValueError("optimizer must be 'adam' or 'sgd' (train.py:66-69)")
        self.optimizer, self.momentum, self.keep_grads = optimizer, float(momentum), bool(keep_grads)
        self.world, self.pg = int(world_size), process_group
        self.store = module.store()
        self.plan = module.plan(self.B, self.H, True)
        dev = self.store.device
        self.device = dev
        self.lib = L.lib()
        J = module._J
        self.J = J
        self.sup_heads = list(self.plan.heads) if all_stacks else [self.plan.heads[-1]]
        self.head = self.plan.heads[-1]
        self.F = self.head.pred.shape[-1]
        n = self.store.params.numel()
        self.m = torch.zeros(n, dtype=torch.float32, device=dev)            # Adam exp_avg / SGD momentum buffer
        self.v = torch.zeros(n, dtype=torch.float32, device=dev) if optimizer == "adam" else None
        self.hyper = torch.tensor([0.0, self.lr], dtype=torch.float32, device=dev)      # [step count, learning rate]: read by the optimizer kernel
        self.step_dev = self.hyper[:1]
        # parameters no launch of the backward plan writes a gradient for (reference: grad None -> torch.optim skips them)
        lay = self.store.layout
        trained = set(self.plan.trained_param_names())
        self.unused_params = [nm for nm in lay.order if nm not in trained]
        merged = []
        for a, b in sorted((lay.specs[nm].offset, lay.specs[nm].offset + (lay.specs[nm].numel + 63) // 64 * 64) for nm in self.unused_params):
            if merged and merged[-1][1] == a:
                merged[-1][1] = b
            else:
                merged.append([a, b])
        if len(merged) > 256:
            raise RuntimeError("more than 256 never-trained parameter spans")
        self.skip_spans = torch.tensor(merged, dtype=torch.int64, device=dev).reshape(-1) if merged else None
        self.n_skip = len(merged)
        self._skip_cache = {}
        # optimizer per gradient bucket, overlapped with the rest of backward (AWR_B200_OPT_OVERLAP=0: one launch after backward)
        self.opt_overlap = os.environ.get("AWR_B200_OPT_OVERLAP", "1") == "1"
        # Hourglass: full-resolution up1 branches on a second stream beside the low-resolution spine (AWR_B200_FWD_OVERLAP=0: one stream)
        self.fwd_overlap = os.environ.get("AWR_B200_FWD_OVERLAP", "1") == "1"
        self.jt = torch.empty(self.B, J, 3, dtype=torch.float32, device=dev)
        ns = len(self.sup_heads)
        self.uvd_all = [torch.empty(self.B, J, 3, dtype=torch.float32, device=dev) for _ in range(ns)]
        self.loss_all = torch.zeros(ns, 2, dtype=torch.float32, device=dev)       # per supervised stack: (SmoothL1 joints, SmoothL1 dense)
        self.ws_all = [torch.zeros(4 * self.B * J + 4, dtype=torch.float32, device=dev) for _ in range(ns)]
        self.uvd, self.loss, self.ws = self.uvd_all[-1], self.loss_all[-1], self.ws_all[-1]        # the last stack (what test.py evaluates)
        self.loss_host = torch.zeros(ns, 2, dtype=torch.float32).pin_memory()
        self.steps_done = 0
        self.use_graph = use_graph
        self.graph_fb = self.graph_opt = self.graph_step = None
        self._pipe = None
        self.side = torch.cuda.Stream(device=dev)          # weight-gradient GEMMs overlap the dgrad / BatchNorm backward chain
        if self.plan.precision == "bf16":
            self.store.refresh_shadow()
        # launches of OUR kernels per step (memsets / NCCL not counted)
        n_opt = len(self.plan.bwd_splits) + 1 if (use_graph and self.opt_overlap and self.plan.bwd_splits) else 1
        self.launches_per_step = len(self.plan.fwd) + len(self.plan.bwd) + 2 * ns + 1 + n_opt

    # ---- launch sequences -------------------------------------------------------------------------------
    def _fwd_bwd(self, part=None, join=True):
        """part None: whole forward+backward; 0: forward + head + backward up to the first gradient-bucket split; i > 0: backward
        between splits i-1 and i (engine.Plan.bwd_splits)."""
        pl, st, s = self.plan, self.store, L.stream()
        if part is not None and part > 0:
            pl.run_backward(s, side=self.side, part=part, join=join)
            return
        au = pl.arena_used()
        L.check(self.lib.awr_memset_zero(au.data_ptr(), au.numel() * 4, s), "awr_memset_zero")        # BN accumulators: a memset node, no fill kernel
        if self.keep_grads:                      # otherwise the optimizer kernel of the previous step left the buffer zeroed
            L.check(self.lib.awr_memset_zero(st.grads.data_ptr(), st.grads.numel() * 4, s), "awr_memset_zero")
        for h in pl.heads:
            if h not in self.sup_heads:
                L.check(self.lib.awr_memset_zero(h.dpred.data_ptr(), h.dpred.numel() * 4, s), "awr_memset_zero")
        pl.run_forward(s, side=self.side if self.fwd_overlap else None)
        for i, hd in enumerate(self.sup_heads):
            uvd, ws, loss = self.uvd_all[i], self.ws_all[i], self.loss_all[i]
            L.check(self.lib.awr_head_fwd(hd.pred.data_ptr(), L.F32, pl.img.data_ptr(), self.jt.data_ptr(), uvd.data_ptr(),
                                          loss.data_ptr(), ws.data_ptr(), self.B, self.J, self.F, self.H, self.ks, s), "awr_head_fwd")
            L.check(self.lib.awr_head_bwd(hd.pred.data_ptr(), L.F32, pl.img.data_ptr(), self.jt.data_ptr(), uvd.data_ptr(),
                                          ws.data_ptr(), None, None, hd.dpred.data_ptr(), self.B, self.J, self.F, self.H, self.ks,
                                          self.cw, self.dw, s), "awr_head_bwd")
        pl.run_backward(s, side=self.side, part=part, join=join)

    def _tick(self):
        L.check(self.lib.awr_adam_tick(self.step_dev.data_ptr(), L.stream()), "awr_adam_tick")

    def _opt(self, a=None, b=None, tick=True):
        """The optimizer over the flat buffers, or (a, b given) over elements [a, b) only: the bucket whose gradients have just become
        final (and, data parallel, summed), so that its update overlaps the rest of backward."""
        st, s = self.store, L.stream()
        if tick:
            self._tick()
        n = st.params.numel()
        a, b = (0, n) if a is None else (int(a), int(b))
        if b <= a:
            return
        if a % 4:
            raise RuntimeError("optimizer bucket not 16-byte aligned")
        skip, n_skip = self._skip_for(a, b)
        off = 4 * a
        shadow = st.shadow.data_ptr() + 2 * a if self.plan.precision == "bf16" else None
        zero = 0 if self.keep_grads else 1
        if self.optimizer == "adam":
            L.check(self.lib.awr_optim_adam(st.params.data_ptr() + off, st.grads.data_ptr() + off, self.m.data_ptr() + off, self.v.data_ptr() + off,
                                            shadow, b - a, self.hyper.data_ptr(), self.betas[0], self.betas[1], self.eps, self.wd,
                                            1.0 / self.world, skip, n_skip, zero, s), "awr_optim_adam")
        else:
            L.check(self.lib.awr_optim_sgd(st.params.data_ptr() + off, st.grads.data_ptr() + off, self.m.data_ptr() + off, shadow, b - a,
                                           self.hyper.data_ptr(), self.momentum, self.wd, 1.0 / self.world, skip, n_skip, zero, s),
                    "awr_optim_sgd")

    def _skip_for(self, a, b):
        """Never-trained spans clipped to [a, b) and rebased to a (device int64 pairs, cached per bucket)."""
        if not self.n_skip:
            return None, 0
        key = (a, b)
        if key not in self._skip_cache:
            sp = self.skip_spans.view(-1, 2).tolist()
            cl = [[max(x, a) - a, min(y, b) - a] for x, y in sp if min(y, b) > max(x, a)]
            self._skip_cache[key] = (torch.tensor(cl, dtype=torch.int64, device=self.device).reshape(-1) if cl else None, len(cl))
        t, k = self._skip_cache[key]
        return (t.data_ptr() if t is not None else None), k

    def _allreduce(self):
        if self.world > 1:
            dp.allreduce_sum_(self.store.grads, self.pg)        # NCCL over NVLink; 1/world is applied inside the optimizer kernel

    def _overlapped(self):
        """Fallback data-parallel step (AWR_B200_DP_GRAPH=0): graph A -> async NCCL on bucket 1 -> graph B -> async NCCL on bucket 0 ->
        wait both -> optimizer graph, all driven from the host."""
        import torch.distributed as dist
        g, pl = self.store.grads, self.plan
        works = []
        for part, gr in enumerate(self.graph_parts):
            gr.replay()
            a, b = pl.bucket_range(part)
            works.append(dist.all_reduce(g[a:b], op=dist.ReduceOp.SUM, group=self.pg, async_op=True))
        for w in works:
            w.wait()
        self.graph_opt.replay()

    def _dp_step_body(self):
        """The whole step as ONE capturable sequence.  Backward runs in parts that end where a gradient bucket becomes final
        (engine.Plan.bwd_splits: ~80 %, 95 %, 99 % of the bytes, last layers first).  As each bucket closes, a second stream takes it:
        data parallel, its NCCL all-reduce (captured into the graph; `nccl_sms` SMs are kept free of persistent conv CTAs while one is in
        flight -- round 1 measured 96 % of the all-reduce exposed without that); then the optimizer over exactly that range.  Both overlap
        the rest of backward, which no longer reads those layers' weights.  What is left after backward is the sub-megabyte last bucket."""
        import torch.distributed as dist
        main = torch.cuda.current_stream()
        g = self.store.grads
        pl = self.plan
        dpar = self.world > 1
        self._tick()
        if not pl.bwd_splits:
            self._fwd_bwd()
            if dpar:
                dist.all_reduce(g, op=dist.ReduceOp.SUM, group=self.pg)
            self._opt(tick=False)
            return
        nparts = len(pl.bwd_splits) + 1
        prev = None
        for part in range(nparts):
            last = part == nparts - 1
            self._fwd_bwd(part=part, join=last or not self.opt_overlap)
            a, b = pl.bucket_range(part)
            if part == 0 and dpar:                            # from here on a collective is in flight: leave it its SMs
                prev = self.lib.awr_set_sm_budget(max(8, 148 - self.nccl_sms))
            if not last:
                self.comm.wait_stream(main)
                if self.opt_overlap:
                    self.side.wait_stream(main)               # (keeps `side` inside the capture even when this part launched nothing on it)
                    self.comm.wait_stream(self.side)          # the bucket's weight gradients (main did not join them)
                with torch.cuda.stream(self.comm):
                    if dpar:
                        dist.all_reduce(g[a:b], op=dist.ReduceOp.SUM, group=self.pg)
                    if self.opt_overlap:
                        self._opt(a, b, tick=False)
            else:
                if dpar:
                    self.lib.awr_set_sm_budget(prev)
                    dist.all_reduce(g[a:b], op=dist.ReduceOp.SUM, group=self.pg)      # the last, sub-megabyte bucket: nothing left to hide it behind
                main.wait_stream(self.comm)
                if self.opt_overlap:
                    self._opt(a, b, tick=False)
                else:
                    self._opt(tick=False)

    def _capture(self):
        # warm-up outside capture (lazy module loads, first-touch).  It is a real forward/backward on the batch in the static buffers:
        # the BatchNorm running statistics / num_batches_tracked it advanced are put back, and its gradients are discarded, so the first
        # replayed step starts from exactly the state the reference would be in.
        saved = {k: b.clone() for k, b in self.store.buffers.items()}
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            self._fwd_bwd()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        for k, b in self.store.buffers.items():
            b.copy_(saved[k])
        self.store.grads.zero_()
        self.split = self.world > 1 and bool(self.plan.bwd_splits) and os.environ.get("AWR_B200_NO_OVERLAP") != "1"
        self.graph_step = None
        one_graph = (self.world > 1 and os.environ.get("AWR_B200_DP_GRAPH", "1") == "1") or \
            (self.world == 1 and self.opt_overlap and bool(self.plan.bwd_splits))
        if one_graph:
            if self.world > 1:
                import torch.distributed as dist
                warm = torch.zeros(8, device=self.device)
                dist.all_reduce(warm, group=self.pg)             # communicator / channels exist before capture
                torch.cuda.synchronize()
            self.comm = torch.cuda.Stream(device=self.device)
            for part in range(len(self.plan.bwd_splits) + 1):       # per-bucket skip tables are built (H2D) before the capture
                self._skip_for(*self.plan.bucket_range(part))
            # SMs kept free of persistent conv CTAs while a bucket is in flight (AWR_B200_SM_RESERVE; default = NCCL's CTA cap)
            self.nccl_sms = int(os.environ.get("AWR_B200_SM_RESERVE", os.environ.get("AWR_B200_NCCL_SMS", "16")))
            self.graph_step = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph_step):
                self._dp_step_body()
            self.graph_fb = self.graph_step
            return
        self.graph_fb = torch.cuda.CUDAGraph()
        self.graph_fb2 = None
        if self.split:
            self.graph_parts = []
            for part in range(len(self.plan.bwd_splits) + 1):
                gr = self.graph_fb if part == 0 else torch.cuda.CUDAGraph()
                with torch.cuda.graph(gr):
                    self._fwd_bwd(part=part)
                self.graph_parts.append(gr)
        else:
            with torch.cuda.graph(self.graph_fb):
                self._fwd_bwd()
        self.graph_opt = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph_opt):
            self._opt()

    # ---- public API -------------------------------------------------------------------------------------
    def load_batch(self, img, jt_uvd_gt):
        """Copy one batch (host pinned or device tensors) into the static input buffers (async on the current stream)."""
        self.plan.img.copy_(img.view(self.B, 1, self.H, self.H), non_blocking=True)
        self.jt.copy_(jt_uvd_gt, non_blocking=True)

    def run_step(self):
        """One optimisation step on the batch currently in the static buffers; no host sync."""
        if self.plan.precision == "bf16" and getattr(self.module, "_params_dirty", False):
            self.store.refresh_shadow()          # load_state_dict() after construction: the bf16 weight shadow follows the new parameters
            self.module._params_dirty = False
        if self.use_graph:
            if self.graph_fb is None:
                self._capture()
            if self.graph_step is not None:
                self.graph_step.replay()
            elif self.split:
                self._overlapped()
            else:
                self.graph_fb.replay()
                self._allreduce()
                self.graph_opt.replay()
        else:
            if not self.keep_grads and self.steps_done == 0:
                self.store.grads.zero_()
            self._fwd_bwd()
            self._allreduce()
            self._opt()
        self.steps_done += 1

    def train_step(self, img, jt_uvd_gt):
        """The call a training loop makes: host (pinned) or device batch in, python floats (loss_coord, loss_dense) out.
        Includes the H2D copies and the D2H loss read-back (train.py:109-133)."""
        self.load_batch(img, jt_uvd_gt)
        self.run_step()
        self.loss_host.copy_(self.loss_all, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return float(self.loss_host[:, 0].sum()), float(self.loss_host[:, 1].sum())       # sums over the supervised stacks

    # ---- pipelined form of train_step: the host never waits for the step it has just enqueued -------------------------------
    def submit(self, img, jt_uvd_gt):
        """Enqueue one optimisation step on a host (pinned) or device batch and return immediately.  The H2D copy runs on a copy
        stream into one of two staging slots, so it overlaps the previous step's kernels; the step's losses are copied D2H behind it
        and are returned by the matching collect().  At most two steps may be outstanding (submit, submit, collect, submit, ...)."""
        if self._pipe is None:
            dev, B, H, J = self.device, self.B, self.H, self.J
            self._pipe = {
                "copy": torch.cuda.Stream(device=dev),
                "img": [torch.empty(B, 1, H, H, dtype=torch.float32, device=dev) for _ in range(2)],
                "jt": [torch.empty(B, J, 3, dtype=torch.float32, device=dev) for _ in range(2)],
                "ready": [torch.cuda.Event() for _ in range(2)], "free": [torch.cuda.Event() for _ in range(2)],
                "done": [torch.cuda.Event() for _ in range(2)],
                "loss": [torch.zeros(len(self.sup_heads), 2, dtype=torch.float32).pin_memory() for _ in range(2)],
                "submitted": 0, "collected": 0}
        pp = self._pipe
        if pp["submitted"] - pp["collected"] >= 2:
            raise RuntimeError("FusedTrainer.submit: two steps already outstanding; call collect() first")
        slot = pp["submitted"] % 2
        main, cs = torch.cuda.current_stream(), pp["copy"]
        cs.wait_event(pp["free"][slot])                     # the step that last used this slot has copied it out (no-op the first time)
        with torch.cuda.stream(cs):
            pp["img"][slot].copy_(img.view(self.B, 1, self.H, self.H), non_blocking=True)
            pp["jt"][slot].copy_(jt_uvd_gt, non_blocking=True)
            pp["ready"][slot].record(cs)
        main.wait_event(pp["ready"][slot])
        self.plan.img.copy_(pp["img"][slot], non_blocking=True)
        self.jt.copy_(pp["jt"][slot], non_blocking=True)
        pp["free"][slot].record(main)
        self.run_step()
        pp["loss"][slot].copy_(self.loss_all, non_blocking=True)
        pp["done"][slot].record(main)
        pp["submitted"] += 1

    def collect(self):
        """(loss_coord, loss_dense) of the oldest submitted step not collected yet; blocks until that step has finished."""
        pp = self._pipe
        if pp is None or pp["collected"] >= pp["submitted"]:
            raise RuntimeError("FusedTrainer.collect: nothing outstanding")
        slot = pp["collected"] % 2
        pp["done"][slot].synchronize()
        pp["collected"] += 1
        return float(pp["loss"][slot][:, 0].sum()), float(pp["loss"][slot][:, 1].sum())

    def train_step_lagged(self, img, jt_uvd_gt):
        """train_step with one step of lag on the logged losses: enqueues this batch and returns the losses of the PREVIOUS call (None on
        the first).  Same work per call as train_step -- H2D of this batch, the step, D2H of its losses -- but the host read never
        drains the queue, so the copy of batch k+1 and the launch of step k+1 overlap step k.  Finish with collect()."""
        self.submit(img, jt_uvd_gt)
        pp = self._pipe
        return self.collect() if pp["submitted"] - pp["collected"] == 2 else None

    def feed_eval(self, eval_tool, jt_xyz_gt, center_xyz, M, cube):
        """train.py:141-148 without the per-frame `.cpu()` copies: hands the step's predicted UVD (still on the device) and the batch's
        ground truth / crop geometry to awr_b200.EvalUtil.feed_batch.  Enqueue it right after the step; no host synchronisation."""
        eval_tool.feed_batch(self.uvd, jt_xyz_gt, center_xyz, M, cube)

    def broadcast_parameters(self, src=0):
        """DDP-style start: every replica takes rank `src`'s parameters and BN buffers."""
        if self.world > 1:
            dp.broadcast_([self.store.params] + list(self.store.buffers.values()), src, self.pg)
            if self.plan.precision == "bf16":
                self.store.refresh_shadow()

    def release(self):
        """Drop the captured graphs (they hold the NCCL communicator busy: do this before tearing the process group down)."""
        torch.cuda.synchronize()
        self.graph_fb = self.graph_opt = self.graph_step = None
        self.graph_parts = []

    def set_lr(self, lr):
        """Learning-rate schedules (train.py:89-96,157-160: StepLR / ReduceLROnPlateau drive `optimizer.param_groups[0]['lr']`): the rate is
        a device scalar the optimizer kernel reads, so the captured graphs are untouched (awr_b200.optim wraps this)."""
        self.lr = float(lr)
        self.hyper[1:2].fill_(self.lr)

    @property
    def param_groups(self):
        """torch.optim-style view for schedulers and logging (train.py:155 prints optimizer.param_groups[0]['lr'])."""
        return [_LRGroup(self)]

    def load_optimizer_state_dict(self, sd):
        """Inverse of optimizer_state_dict(): resume from the `optimizer` entry of a checkpoint (train.py:82-84), i.e. a torch.optim
        Adam / SGD state over the canonical parameter order.  State keys are whatever `param_groups[0]['params']` lists -- positions in
        current torch, `id(param)` values in the torch<=1.1 checkpoints the reference ships (results/hourglass_1.pth: 220 entries for
        250 parameters: the never-trained ones are absent)."""
        lay = self.store.layout
        self.m.zero_()
        if self.v is not None:
            self.v.zero_()
        steps = 0
        for name, st in map_optimizer_state(sd, self.module._pnames).items():
            if self.optimizer == "adam":
                lay.view(self.m, name).copy_(st["exp_avg"].to(self.device))
                lay.view(self.v, name).copy_(st["exp_avg_sq"].to(self.device))
                steps = max(steps, int(float(st["step"])))
            elif st.get("momentum_buffer") is not None:
                lay.view(self.m, name).copy_(st["momentum_buffer"].to(self.device))
        if self.optimizer == "adam":
            self.steps_done = steps
            self.hyper[0:1].fill_(float(steps))
        self.set_lr(sd["param_groups"][0]["lr"])

    def optimizer_state_dict(self):
        """torch.optim-shaped state (train.py:165-172 saves optimizer.state_dict()) over the canonical parameters; like torch, parameters
        that never received a gradient have no state entry."""
        lay = self.store.layout
        state = {}
        unused = set(self.unused_params)
        for i, name in enumerate(self.module._pnames):
            if name in unused or self.steps_done == 0:
                continue
            if self.optimizer == "adam":
                state[i] = {"step": torch.tensor(float(self.steps_done)), "exp_avg": lay.view(self.m, name).clone(),
                            "exp_avg_sq": lay.view(self.v, name).clone()}
            else:
                state[i] = {"momentum_buffer": lay.view(self.m, name).clone()}
        if self.optimizer == "adam":
            group = {"lr": self.lr, "betas": tuple(self.betas), "eps": self.eps, "weight_decay": self.wd, "amsgrad": False}
        else:
            group = {"lr": self.lr, "momentum": self.momentum, "dampening": 0, "weight_decay": self.wd, "nesterov": False}
        group["params"] = list(range(len(self.module._pnames)))
        return {"state": state, "param_groups": [group]}

    # ---- checkpoints in the reference's layout (train.py:165-172 / :80-86; test.py:45-49) -------------------------------------------
    def save_checkpoint(self, path, best_records=None, rank=0):
        """{'model', 'optimizer', 'best_records'} exactly as train.py:165-172 writes it: `module.`-free keys, fp32 CPU tensors.  Under data
        parallelism only `rank` 0 writes (every replica holds the same parameters; BatchNorm running statistics are rank 0's, DDP-style)."""
        if rank != 0:
            return False
        sd = {k: v.detach().cpu().clone() for k, v in self.module.state_dict().items()}
        osd = self.optimizer_state_dict()
        for st in osd["state"].values():
            for k, v in list(st.items()):
                if torch.is_tensor(v):
                    st[k] = v.cpu()
        torch.save({"model": sd, "optimizer": osd, "best_records": dict(best_records or {"epoch": 0, "MPE": 1e10, "AUC": 0})}, path)
        return True

    def load_checkpoint(self, path_or_dict, load_optimizer=True):
        """train.py:80-86: model + optimizer (+ best_records, returned).  Accepts the reference's legacy pickles (numpy scalars in
        best_records), hence weights_only=False: only load checkpoints you trust."""
        ck = torch.load(path_or_dict, map_location="cpu", weights_only=False) if isinstance(path_or_dict, (str, bytes, os.PathLike)) else path_or_dict
        self.module.load_state_dict(ck["model"], strict=True)
        if self.plan.precision == "bf16":
            self.store.refresh_shadow()
            self.module._params_dirty = False
        if load_optimizer and "optimizer" in ck:
            self.load_optimizer_state_dict(ck["optimizer"])
        return ck.get("best_records")


class _LRGroup(dict):
    """param_groups[0] of a FusedTrainer: reading/writing ['lr'] reads/sets the device-side learning rate."""

    def __init__(self, trainer):
        super().__init__(lr=trainer.lr)
        self._tr = trainer

    def __getitem__(self, k):
        return self._tr.lr if k == "lr" else super().__getitem__(k)

    def __setitem__(self, k, v):
        if k == "lr":
            self._tr.set_lr(v)
        super().__setitem__(k, v)
