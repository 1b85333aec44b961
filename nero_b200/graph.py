"""The training render of a fixed ray count as two CUDA graphs (forward / backward).

NeROShapeRenderer.train_step (network/renderer.py:167-187 -> render :445-463 -> render_core :550-606) launches ~230 kernels
per step.  At the reference's default `train_ray_num` = 512 they are a few microseconds each, so the step is bound by the host
launching them and by the one device->host read of the sample counts in the middle of the forward pass.  With
cfg['cuda_graph'] the renderer replays the same kernel sequence from two captured graphs instead:

  forward graph : ray construction (_process_ray_batch) -> weight fold -> hierarchical sampling -> render_core forward ->
                  eikonal mean -> occlusion loss
  backward graph: render_core backward (all parameter gradients, accumulated into the engine's persistent flat buffer)

Nothing in either graph depends on a host-side value: row counts are read by every kernel from device memory, the occlusion
subset is drawn on the device (engine._occ_forward_static), cos_anneal_ratio sits in `engine.car_dev`, and the random draws
are filled into static buffers before each replay.  The loss between the two graphs stays ordinary eager autograd, so the
trainer's loss objects and optimizer work unchanged.

One difference is visible to callers: `gradient_error` is returned already averaged over the inner samples (one element)
instead of per sample -- the eikonal loss takes its mean (network/loss.py:62-65), which is unchanged by that.
Graphs are keyed by the flags that change the kernel sequence (occlusion loss on/off, inv_s frozen or not); the
SDF-initialisation phase (step < 1000, which adds the regularisation pass) is not graphed and runs eagerly.
"""
import torch


class _ReplayFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, g, *params):
        g.fwd.replay()
        ctx.g = g
        ctx.nparams = len(params)
        return g.rgb.clone(), g.gmean.clone(), g.locc.clone()

    @staticmethod
    def backward(ctx, d_rgb, d_g, d_occ):
        g = ctx.g
        g.d_rgb.copy_(d_rgb)
        g.d_g.copy_(d_g.reshape(1))
        g.d_occ.copy_(d_occ.reshape(1))
        g.engine.grads.ensure(strict=True)
        g.bwd.replay()
        return (None,) * (1 + ctx.nparams)


class TrainStepGraphs:
    def __init__(self, net, R, step):
        """net: NeROShapeRenderer; R: rays per step; step: any step of the regime to capture (decides occ / freeze flags)."""
        e = net.engine
        cfg = net.cfg
        dev = e.dev
        self.net, self.engine, self.R, self.step = net, e, R, int(step)
        assert step >= 1000
        self.occ_on = bool(cfg['apply_occ_loss'] and step >= cfg['occ_loss_step'])
        S = cfg['n_samples'] + cfg['n_importance'] + cfg['n_bg_samples']
        e._alloc(R, S)
        R_cap, S_cap = e.cap
        cap = R_cap * S_cap
        z = lambda *s, dt=torch.float32: torch.zeros(*s, dtype=dt, device=dev)
        # static inputs (filled before every replay)
        self.dirs, self.idxs = z(R, 3), z(R, 1, dt=torch.int64)
        self.perturb = cfg['perturb'] > 0
        self.rand_inner, self.rand_bg = z(R, 1), z(R, cfg['n_bg_samples'])
        self.occ_keys = z(cap)
        self.ar = torch.arange(cap, device=dev, dtype=torch.int32)
        self.d_rgb, self.d_g, self.d_occ = z(R, 3), z(1), z(1)
        self.P_dev = None
        self.poses = None
        self.fwd = self.bwd = None

    @staticmethod
    def key(net, R, step):
        cfg = net.cfg
        net.engine._alloc(R, cfg['n_samples'] + cfg['n_importance'] + cfg['n_bg_samples'])
        occ = bool(cfg['apply_occ_loss'] and step >= cfg['occ_loss_step'])
        frozen = cfg['freeze_inv_s_step'] is not None and step < cfg['freeze_inv_s_step']
        return (R, occ, frozen, net.engine.cap)

    # ------------------------------------------------------------------ the two captured bodies
    def _forward_body(self):
        net, e = self.net, self.engine
        rays_o, rays_d, near, far, hp = net._process_ray_batch({'dirs': self.dirs, 'idxs': self.idxs}, self.poses)
        e.prepare_weights()
        with torch.no_grad():
            z_vals = e.sample_ray(rays_o.contiguous(), rays_d.contiguous(), near.contiguous(), far.contiguous(),
                                  self.rand_inner if self.perturb else None, self.rand_bg if self.perturb else None)
            self.rgb = e.render_core_forward(rays_o.contiguous(), rays_d.contiguous(), z_vals, hp.contiguous(), None, self.step,
                                             static=self)
            w = e.w
            n_in = w['n_in']
            cap = w['GERR'].shape[0]
            self.mask = self.ar[:cap] < n_in
            self.nf = n_in.clamp(min=1).float()
            self.gmean = torch.where(self.mask, w['GERR'], 0.0).sum().reshape(1) / self.nf
            if self.occ_on:
                self.Pf = self.P_dev.clamp(min=1).float()
                self.locc = w['OCC_LOSS'] / self.Pf
            else:
                self.locc = torch.zeros(1, device=e.dev)

    def _backward_body(self):
        e = self.engine
        with torch.no_grad():
            dg = torch.where(self.mask, self.d_g / self.nf, 0.0)
            dscale = (self.d_occ / self.Pf).reshape(()) if self.occ_on else None
            e.render_core_backward(self.d_rgb, dg, dscale, static=self)

    def capture(self, poses):
        """Called with this step's inputs already in the static buffers, so the eager warm-up passes run on real rays."""
        self.poses = poses
        e = self.engine
        e.grads.ensure(strict=True)
        kept = e.grads.flat.clone()              # the warm-up passes must not leave gradients behind
        side = torch.cuda.Stream(device=e.dev)
        side.wait_stream(torch.cuda.current_stream(e.dev))
        with torch.cuda.stream(side):            # warm-up: sizes workspaces, builds tensor maps and job tables
            for _ in range(2):
                self._forward_body()
                e._alloc_backward()
                self._backward_body()
        torch.cuda.current_stream(e.dev).wait_stream(side)
        torch.cuda.synchronize(e.dev)
        e.grads.flat.copy_(kept)
        self.fwd = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.fwd):
            self._forward_body()
        self.bwd = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.bwd):
            self._backward_body()

    # ------------------------------------------------------------------ one step
    def run(self, dirs, idxs, poses, cos_anneal_ratio, params):
        """dirs [R,3] / idxs [R,1] may live in pinned host memory: they are copied straight into the graph's input buffers.
        poses [N,3,4]: the device-resident camera table (the same tensor every step)."""
        e = self.engine
        self.dirs.copy_(dirs, non_blocking=True)
        self.idxs.copy_(idxs, non_blocking=True)
        e.car_dev.fill_(float(cos_anneal_ratio))
        if self.perturb:
            self.rand_inner.uniform_()
            self.rand_bg.uniform_()
        if self.occ_on:
            self.occ_keys.uniform_()
        if self.fwd is None or poses.data_ptr() != self.poses.data_ptr():
            self.capture(poses)
        return _ReplayFn.apply(self, *params)
