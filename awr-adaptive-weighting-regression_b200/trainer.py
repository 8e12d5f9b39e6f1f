"""Fused training step of the AWR hot path (the loop body of the reference's train.py:107-131) on libawr_b200.so.

    H2D(img, jt_uvd_gt) -> backbone fwd -> fused head+loss fwd -> fused head+loss bwd -> backbone bwd
        -> [NCCL all-reduce of the flat gradient buffer] -> fused Adam

The step is a fixed launch sequence over static buffers (engine.Plan), captured once into CUDA graphs:
`graph_fb` (zero scratch, forward, head, backward) and `graph_opt` (Adam); the data-parallel all-reduce
runs between them on the same stream through torch.distributed (NCCL over NVLink).  GT volumes, the
coordinate grid and the loss temporaries of the reference are never materialised (csrc/head.cu).

Semantics kept from train.py: loss = coord_weight*SmoothL1(uvd, jt) + dense_weight*SmoothL1(pred, joint2offset(jt))
(:119-120); for 'hourglass_N' only the last stack is supervised (:116-121 overwrite `loss`); Adam(lr, betas
(0.9,0.999), eps 1e-8, weight_decay) (:67); BN statistics are per replica (the reference has no SyncBN).
"""
import os

import torch

from . import _lib as L
from . import dp
from .modules import AWRBackbone


class FusedTrainer:
    def __init__(self, module: AWRBackbone, batch_size, img_size, kernel_size, coord_weight=1.0, dense_weight=1.0, lr=1e-3,
                 betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, world_size=1, use_graph=True, process_group=None, all_stacks=False):
        """all_stacks (hourglass_N, N > 1): supervise every stack and sum the per-stack losses (test.py:74-80); the default keeps
        train.py:116-121's behaviour, where only the last stack's loss survives the loop."""
        if not isinstance(module, AWRBackbone):
            raise TypeError("FusedTrainer drives awr_b200 backbones (get_deconv_net / PoseNet)")
        self.module = module
        module.train()
        self.B, self.H = int(batch_size), int(img_size)
        self.ks, self.cw, self.dw = float(kernel_size), float(coord_weight), float(dense_weight)
        self.lr, self.betas, self.eps, self.wd = float(lr), betas, float(eps), float(weight_decay)
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
        self.m = torch.zeros(n, dtype=torch.float32, device=dev)
        self.v = torch.zeros(n, dtype=torch.float32, device=dev)
        self.step_dev = torch.zeros(1, dtype=torch.float32, device=dev)
        self.jt = torch.empty(self.B, J, 3, dtype=torch.float32, device=dev)
        ns = len(self.sup_heads)
        self.uvd_all = [torch.empty(self.B, J, 3, dtype=torch.float32, device=dev) for _ in range(ns)]
        self.loss_all = torch.zeros(ns, 2, dtype=torch.float32, device=dev)       # per supervised stack: (SmoothL1 joints, SmoothL1 dense)
        self.ws_all = [torch.zeros(4 * self.B * J + 4, dtype=torch.float32, device=dev) for _ in range(ns)]
        self.uvd, self.loss, self.ws = self.uvd_all[-1], self.loss_all[-1], self.ws_all[-1]        # the last stack (what test.py evaluates)
        self.loss_host = torch.zeros(ns, 2, dtype=torch.float32).pin_memory()
        self.steps_done = 0
        self.use_graph = use_graph
        self.graph_fb = self.graph_opt = None
        self._pipe = None
        self.side = torch.cuda.Stream(device=dev)          # weight-gradient GEMMs overlap the dgrad / BatchNorm backward chain
        if self.plan.precision == "bf16":
            self.store.refresh_shadow()
        # launches of OUR kernels per step (memsets / NCCL not counted)
        self.launches_per_step = len(self.plan.fwd) + len(self.plan.bwd) + 2 * ns + 2

    # ---- launch sequences -------------------------------------------------------------------------------
    def _fwd_bwd(self, part=None):
        """part None: whole forward+backward; 0: forward + head + backward up to the gradient-bucket split; 1: rest of backward."""
        pl, st, s = self.plan, self.store, L.stream()
        if part == 1:
            pl.run_backward(s, side=self.side, part=1)
            return
        pl.arena_used().zero_()
        st.grads.zero_()
        for h in pl.heads:
            if h not in self.sup_heads:
                h.dpred.zero_()
        pl.run_forward(s)
        for i, hd in enumerate(self.sup_heads):
            uvd, ws, loss = self.uvd_all[i], self.ws_all[i], self.loss_all[i]
            L.check(self.lib.awr_head_fwd(hd.pred.data_ptr(), L.F32, pl.img.data_ptr(), self.jt.data_ptr(), uvd.data_ptr(),
                                          loss.data_ptr(), ws.data_ptr(), self.B, self.J, self.F, self.H, self.ks, s), "awr_head_fwd")
            L.check(self.lib.awr_head_bwd(hd.pred.data_ptr(), L.F32, pl.img.data_ptr(), self.jt.data_ptr(), uvd.data_ptr(),
                                          ws.data_ptr(), None, None, hd.dpred.data_ptr(), self.B, self.J, self.F, self.H, self.ks,
                                          self.cw, self.dw, s), "awr_head_bwd")
        pl.run_backward(s, side=self.side, part=part)

    def _opt(self):
        st, s = self.store, L.stream()
        L.check(self.lib.awr_adam_tick(self.step_dev.data_ptr(), s), "awr_adam_tick")
        shadow = st.shadow.data_ptr() if self.plan.precision == "bf16" else None
        L.check(self.lib.awr_adam_flat(st.params.data_ptr(), st.grads.data_ptr(), self.m.data_ptr(), self.v.data_ptr(), shadow,
                                       st.params.numel(), self.step_dev.data_ptr(), self.lr, self.betas[0], self.betas[1], self.eps,
                                       self.wd, 1.0 / self.world, s), "awr_adam_flat")

    def _allreduce(self):
        if self.world > 1:
            dp.allreduce_sum_(self.store.grads, self.pg)        # NCCL over NVLink; 1/world is applied inside awr_adam_flat

    def _overlapped(self):
        """Data-parallel step with the all-reduce of the early gradient bucket (last layers: most of the bytes) overlapping the rest of
        backward: graph A -> async NCCL on bucket 1 -> graph B -> async NCCL on bucket 0 -> wait both -> Adam graph."""
        import torch.distributed as dist
        off = self.plan.bwd_split[1]
        g = self.store.grads
        self.graph_fb.replay()
        w1 = dist.all_reduce(g[off:], op=dist.ReduceOp.SUM, group=self.pg, async_op=True)
        self.graph_fb2.replay()
        w0 = dist.all_reduce(g[:off], op=dist.ReduceOp.SUM, group=self.pg, async_op=True)
        w1.wait(); w0.wait()
        self.graph_opt.replay()

    def _capture(self):
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):           # warm-up outside capture (lazy module loads, first-touch)
            self._fwd_bwd()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.split = self.world > 1 and self.plan.bwd_split is not None and os.environ.get("AWR_B200_NO_OVERLAP") != "1"
        self.graph_fb = torch.cuda.CUDAGraph()
        self.graph_fb2 = None
        if self.split:
            with torch.cuda.graph(self.graph_fb):
                self._fwd_bwd(part=0)
            self.graph_fb2 = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph_fb2):
                self._fwd_bwd(part=1)
        else:
            with torch.cuda.graph(self.graph_fb):
                self._fwd_bwd()
        self.graph_opt = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph_opt):
            self._opt()
        # the warm-up / capture passes advanced BN running statistics but not the parameters or the Adam state

    # ---- public API -------------------------------------------------------------------------------------
    def load_batch(self, img, jt_uvd_gt):
        """Copy one batch (host pinned or device tensors) into the static input buffers (async on the current stream)."""
        self.plan.img.copy_(img.view(self.B, 1, self.H, self.H), non_blocking=True)
        self.jt.copy_(jt_uvd_gt, non_blocking=True)

    def run_step(self):
        """One optimisation step on the batch currently in the static buffers; no host sync."""
        if self.use_graph:
            if self.graph_fb is None:
                self._capture()
            if self.split:
                self._overlapped()
            else:
                self.graph_fb.replay()
                self._allreduce()
                self.graph_opt.replay()
        else:
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

    def set_lr(self, lr):
        """Learning-rate schedules (train.py:68-69 StepLR / ReduceLROnPlateau drive `optimizer.param_groups[0]['lr']`): the rate is a
        launch argument of the captured Adam graph, so a change re-captures that two-kernel graph; the forward/backward graph is untouched."""
        lr = float(lr)
        if lr == self.lr:
            return
        self.lr = lr
        if self.graph_opt is not None:
            self.graph_opt = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph_opt):
                self._opt()

    def load_optimizer_state_dict(self, sd):
        """Inverse of optimizer_state_dict(): resume from the `optimizer` entry of a checkpoint (train.py:89-96), i.e. a torch.optim.Adam
        state over the canonical parameter order.  Parameters the reference never steps (unused Hourglass skip layers) may be absent."""
        lay = self.store.layout
        self.m.zero_(); self.v.zero_()
        steps = 0
        for i, name in enumerate(self.module._pnames):
            st = sd["state"].get(i)
            if st is None:
                continue
            lay.view(self.m, name).copy_(st["exp_avg"].to(self.device))
            lay.view(self.v, name).copy_(st["exp_avg_sq"].to(self.device))
            steps = max(steps, int(float(st["step"])))
        self.steps_done = steps
        self.step_dev.fill_(float(steps))
        self.set_lr(sd["param_groups"][0]["lr"])

    def optimizer_state_dict(self):
        """torch.optim.Adam-shaped state (train.py:165-172 saves optimizer.state_dict()) over the canonical parameters."""
        lay, st = self.store.layout, self.store
        state = {}
        for i, name in enumerate(self.module._pnames):
            state[i] = {"step": torch.tensor(float(self.steps_done)), "exp_avg": lay.view(self.m, name).clone(),
                        "exp_avg_sq": lay.view(self.v, name).clone()}
        group = {"lr": self.lr, "betas": tuple(self.betas), "eps": self.eps, "weight_decay": self.wd, "amsgrad": False,
                 "params": list(range(len(self.module._pnames)))}
        return {"state": state, "param_groups": [group]}
