"""TrainEngine: one NeRF optimisation step as a single CUDA graph (SURVEY.md section 7 step 6, rows f1 + f2).

Work of one step = training_step_nerf of the reference (joint_train_lightning_net.py:473-513): render the ray batch
with perturb=True, the three losses, backward, Adam on both parameter groups.  Here it is a fixed chain of
libucsa_nerf.so kernels on a static workspace -- no autograd engine, no eager tensor math, no host synchronisation:

    step counter += 1 -> zero gradients -> pipeline.forward_chain -> ucsa_nerf_loss -> pipeline.backward_chain
    -> world == 1 : ucsa_grad_check -> ucsa_adam_step (ONE launch over the flat parameter space)
       world  > 1 : ucsa_grad_check -> barrier -> ucsa_adam_exchange -> barrier   (exchange = "peer", one NVLink domain)
                    NCCL all-reduce of the flat gradients -> ucsa_grad_check -> ucsa_adam_step   (exchange = "nccl")

ucsa_grad_check is GradScaler's overflow test (joint_train_lightning_net.py:46,509-513): the backward runs in fp16
with a fixed loss scale, so an overflow shows up as inf / NaN in the flat gradient; such a step is skipped on every
rank (moments, masters and fp16 copies untouched) and does not advance Adam's bias-correction count.

captured once with torch.cuda.graph and replayed every step.  Random numbers and Adam's bias correction read the
step counter from device memory, so every replay draws fresh samples.

exchange = "peer": gradients, fp32 masters and the fp16 working copies live in symmetric memory; each rank sums the
gradients of, updates and re-broadcasts 1/world of the flat parameter space in ONE kernel (multimem.ld_reduce /
multimem.st through the NVSwitch when multicast is available, plain peer loads / stores otherwise), so the whole step
is one graph and the Adam work per rank shrinks with the world size.  exchange = "nccl" keeps the all-reduce outside
the graphs (forward/backward graph -> all-reduce -> optimizer graph)."""
from __future__ import annotations

import torch

from . import ops, parallel, pipeline


class TrainEngine:

    def __init__(self, net, n_rays, num_steps=256, upsample_steps=256, lr=1e-2, betas=(0.9, 0.99), eps=1e-15,
                 weight_decay_net=1e-6, weight_depth=0.1, weight_semantics=0.04, one_m_to_scene_uom=1.0, seed=0x5EED,
                 use_graph=True, exchange=None):
        self.net = net
        dev = net.encoder.params.device
        self.device = dev
        self.rank, self.world = parallel.world()
        self.n = n_rays
        self.c = net.num_semantic_classes
        self.lr, self.betas, self.eps = lr, betas, eps
        self.w_depth, self.w_sem, self.uom = weight_depth, weight_semantics, one_m_to_scene_uom
        self.seed = seed
        self.use_graph = use_graph
        # test hooks: injected random draws ([n, num_steps] stratified jitter, [n, upsample_steps] inverse-CDF samples)
        # instead of the counter-based generator, so that a CPU oracle can see the same numbers
        self.t_rand = self.u = None
        self.ws = pipeline.RenderWorkspace(n_rays, num_steps, upsample_steps, self.c, dev, need_grad=True)
        f32 = dict(dtype=torch.float32, device=dev)
        # static inputs (the step's batch is copied in before the replay)
        self.rays_o = torch.zeros(n_rays, 3, **f32)
        self.rays_d = torch.zeros(n_rays, 3, **f32)
        self.dnorm = torch.ones(n_rays, **f32)
        self.gt_rgb = torch.zeros(n_rays, 3, dtype=torch.float16, device=dev)
        self.labels = torch.zeros(n_rays, dtype=torch.int64, device=dev)
        self.gt_depth = torch.zeros(n_rays, **f32)
        self.ray_base = self.rank * n_rays
        # loss + its gradients
        self.loss = torch.zeros(4, **f32)
        self.g_image = torch.empty(n_rays, 3, **f32)
        self.g_depth = torch.empty(n_rays, **f32)
        self.g_sem = torch.empty(n_rays, self.c, **f32)
        # parameters: one flat gradient buffer (the all-reduce payload), Adam moments, device step counter
        self.groups = [(net.encoder, 0.0), (net.sigma_net, weight_decay_net), (net.color_net, weight_decay_net),
                       (net.semantics_net, weight_decay_net)]
        sizes = [m.params.numel() for m, _ in self.groups]
        if exchange is None:
            exchange = "peer" if parallel.peer_exchange_available() else "nccl"
        if exchange not in ("peer", "nccl"):
            raise ValueError("exchange must be 'peer', 'nccl' or None")
        self.exchange = exchange if self.world > 1 else "none"
        self.wd_net = weight_decay_net
        if self.world > 1:  # replicas must start identical (both exchange modes rely on it)
            for m, _ in self.groups:
                torch.distributed.broadcast(m.params.data, src=0)
        self.peer = None
        if self.exchange == "peer":
            # symmetric memory needs peer access between all ranks; agree on the outcome before relying on it
            try:
                self.peer = parallel.PeerExchange(sum(sizes), dev)
                ok = 1.0
            except Exception as exc:  # noqa: BLE001 - any failure means "use NCCL", never a silent CPU path
                import sys

                sys.stderr.write(f"[ucsa] peer-memory gradient exchange unavailable ({exc!r}); using NCCL all-reduce\n")
                ok = 0.0
            flag = torch.tensor([ok], device=dev)
            torch.distributed.all_reduce(flag, op=torch.distributed.ReduceOp.MIN)
            if float(flag) == 0.0:
                self.peer, self.exchange = None, "nccl"
        if self.exchange == "peer":
            self.flat_grad = self.peer.grad
            off = 0
            with torch.no_grad():
                for (m, _), s in zip(self.groups, sizes):  # re-home masters and fp16 copies in symmetric memory
                    view = self.peer.param[off:off + s]
                    view.copy_(m.params.data.reshape(-1))
                    m.params.data = view.view(m.params.shape)
                    m._half = self.peer.param_h[off:off + s].view(m.params.shape)
                    m._half_key = None
                    off += s
            n_own = self.peer.end - self.peer.begin
            self.exp_avg = [torch.zeros(n_own, **f32)]
            self.exp_avg_sq = [torch.zeros(n_own, **f32)]
        else:
            # one flat parameter space here too (masters, fp16 copies, moments, gradients): the optimizer is ONE launch
            # over it, weight decay from the first MLP parameter on
            total = sum(sizes)
            self.flat_grad = torch.zeros(total, **f32)
            self.flat_param = torch.empty(total, **f32)
            self.flat_half = torch.empty(total, dtype=torch.float16, device=dev)
            off = 0
            with torch.no_grad():
                for (m, _), s in zip(self.groups, sizes):
                    view = self.flat_param[off:off + s]
                    view.copy_(m.params.data.reshape(-1))
                    m.params.data = view.view(m.params.shape)
                    m._half = self.flat_half[off:off + s].view(m.params.shape)
                    m._half_key = None
                    off += s
            self.exp_avg = [torch.zeros(total, **f32)]
            self.exp_avg_sq = [torch.zeros(total, **f32)]
        self.grads, off = [], 0
        for s in sizes:
            self.grads.append(self.flat_grad[off:off + s])
            off += s
        self.step_dev = torch.zeros(1, dtype=torch.int32, device=dev)
        # GradScaler-style overflow handling, all on the device: flag of this step, number of skipped steps so far
        self.found_inf = self.peer.flag[:1] if self.peer is not None else torch.zeros(1, **f32)
        self.skipped_dev = torch.zeros(1, dtype=torch.int32, device=dev)
        self._check_scratch = torch.zeros(2, dtype=torch.int32, device=dev)
        self._loss_scratch = ops.loss_scratch(dev)
        for m, _ in self.groups:
            m.half_params()
        self._graph_fb = self._graph_opt = None

    # ------------------------------------------------------------------ the two halves of a step
    def _forward_backward(self):
        net, ws = self.net, self.ws
        self.step_dev.add_(1)
        self.flat_grad.zero_()
        aabb = net.aabb_train
        pipeline.forward_chain(net, ws, self.rays_o, self.rays_d, self.dnorm, aabb, perturb=True, t_rand=self.t_rand,
                               u=self.u, seed=self.seed, ray_base=self.ray_base, step_dev=self.step_dev)
        ops.nerf_loss(ws.image, ws.depth, ws.semantics, self.gt_rgb, self.labels, self.gt_depth, self.uom, self.w_sem,
                      self.w_depth, 1.0 / self.world, self.loss, self.g_image, self.g_depth, self.g_sem,
                      self._loss_scratch)
        pipeline.backward_chain(net, ws, self.rays_o, self.rays_d, self.dnorm, aabb, self.g_image, self.g_depth,
                                self.g_sem, *self.grads)

    def _optimizer(self):
        if self.exchange == "peer":
            ops.grad_check(self.flat_grad, self.found_inf, self._check_scratch)  # this rank's flag, read by all ranks
            self.peer.barrier(0)  # every rank's gradients (and flag) are complete
            ops.adam_exchange(self.peer, self.exp_avg[0], self.exp_avg_sq[0], wd_begin=self.groups[0][0].params.numel(),
                              lr=self.lr, beta1=self.betas[0], beta2=self.betas[1], eps=self.eps,
                              weight_decay=self.wd_net, step=1, step_dev=self.step_dev,
                              found_inf_ptrs=self.peer.flag_ptrs, skipped_dev=self.skipped_dev,
                              broadcast_masters=self.peer.broadcast_masters)
            self.peer.barrier(1)  # every rank's parameters are written; gradients may be zeroed again
            return
        # (exchange == "nccl": the all-reduce already spread any inf / NaN to every rank's copy)
        ops.grad_check(self.flat_grad, self.found_inf, self._check_scratch, skipped_dev=self.skipped_dev)
        for m, _ in self.groups:  # (masters changed from outside: version check + recast, no launch otherwise)
            m.half_params()
        ops.adam_step(self.flat_param, self.flat_grad, self.exp_avg[0], self.exp_avg_sq[0], self.flat_half, lr=self.lr,
                      beta1=self.betas[0], beta2=self.betas[1], eps=self.eps, weight_decay=self.wd_net,
                      wd_begin=self.groups[0][0].params.numel(), grad_scale_inv=1.0, found_inf=self.found_inf, step=1,
                      step_dev=self.step_dev, skipped_dev=self.skipped_dev)

    def _capture(self):
        side = torch.cuda.Stream(device=self.device)
        state = self._snapshot()  # (before the side stream is forked: the clones must not race the warm-up steps)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):  # warm-up outside capture (lazy attribute set-up, allocator)
            for _ in range(2):
                self._forward_backward()
                if self.exchange == "nccl":
                    parallel.all_reduce_gradients(self.flat_grad)
                self._optimizer()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self._restore(state)
        self._graph_fb = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph_fb):
            self._forward_backward()
            if self.exchange != "nccl":
                self._optimizer()
        if self.exchange == "nccl":
            self._graph_opt = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._graph_opt):
                self._optimizer()
        torch.cuda.synchronize()
        self._restore(state)

    def _snapshot(self):
        return ([m.params.detach().clone() for m, _ in self.groups], [t.clone() for t in self.exp_avg],
                [t.clone() for t in self.exp_avg_sq], self.step_dev.clone(), self.skipped_dev.clone())

    def _restore(self, state):
        params, ea, eas, step, skipped = state
        with torch.no_grad():
            for (m, _), p in zip(self.groups, params):
                m.params.copy_(p)
                m.half_params()
            for dst, src in zip(self.exp_avg, ea):
                dst.copy_(src)
            for dst, src in zip(self.exp_avg_sq, eas):
                dst.copy_(src)
            self.step_dev.copy_(step)
            self.skipped_dev.copy_(skipped)

    # ------------------------------------------------------------------ public
    def load_batch(self, rays_o, rays_d, direction_norms, gt_rgb, labels, gt_depth, non_blocking=True):
        """Copy one step's rays + ground truth (host or device tensors, any leading 1-dims) into the static inputs."""
        self.rays_o.copy_(rays_o.reshape(-1, 3), non_blocking=non_blocking)
        self.rays_d.copy_(rays_d.reshape(-1, 3), non_blocking=non_blocking)
        self.dnorm.copy_(direction_norms.reshape(-1), non_blocking=non_blocking)
        self.gt_rgb.copy_(gt_rgb.reshape(-1, 3), non_blocking=non_blocking)
        self.labels.copy_(labels.reshape(-1), non_blocking=non_blocking)
        self.gt_depth.copy_(gt_depth.reshape(-1), non_blocking=non_blocking)

    def step(self):
        """One optimisation step on the loaded batch; returns the device tensor (total, colour, semantic, depth)."""
        if not self.use_graph:
            self._forward_backward()
            if self.exchange == "nccl":
                parallel.all_reduce_gradients(self.flat_grad)
            self._optimizer()
            return self.loss
        if self._graph_fb is None:
            if self.peer is not None and not self.peer.broadcast_masters:
                self.peer.gather_masters()  # the capture snapshots / restores the full masters
            self._capture()
        for m, _ in self.groups:  # masters changed outside the graph (load_state_dict, copy_): recast before replay
            m.half_params()
        self._graph_fb.replay()
        if self.exchange == "nccl":
            parallel.all_reduce_gradients(self.flat_grad)
            self._graph_opt.replay()
        return self.loss

    def gather_masters(self):
        """Collective.  With the peer exchange the fp32 masters of a parameter slice live on its owner only; call this
        on every rank before reading parameters on the host side (net.state_dict(), checkpoints, evaluation of
        p.data).  No-op on one GPU, with exchange="nccl" or with UCSA_PEER_BROADCAST_MASTERS=1."""
        if self.peer is not None:
            torch.cuda.current_stream().synchronize()
            self.peer.gather_masters()

    def state_dict(self):
        """Collective: complete masters, then the network's own state_dict."""
        self.gather_masters()
        return self.net.state_dict()

    def reload_params(self, src=0):
        """Collective.  After the masters were changed outside the engine on rank `src` (load_state_dict, a restored
        checkpoint): broadcast them, recast the fp16 working copies, keep the replicas identical."""
        with torch.no_grad():
            for m, _ in self.groups:
                if self.world > 1:
                    torch.distributed.broadcast(m.params.data, src=src)
                m._half_key = None
                m.half_params()

    @torch.no_grad()
    def exchange_check(self):
        """Collective (every rank calls it).  Evidence that the fused peer-memory exchange is the optimisation the
        NCCL formulation defines: ONE extra step on the loaded batch is taken through ucsa_adam_exchange, and the same
        step is replayed on private copies as NCCL all-reduce(sum) + ucsa_grad_check + ucsa_adam_step.  Returns
        {replicas_identical, half_consistent, max_rel_diff_vs_nccl, frac_within_1e-3, ...} (None without a peer
        exchange).  The step is a real one: parameters and moments advance."""
        if self.peer is None:
            return None
        import torch.distributed as dist

        peer, dev, w = self.peer, self.device, self.world
        n = peer.n
        torch.cuda.synchronize()
        self.gather_masters()
        per = parallel.owner_slice(n, 0, w)[1]

        def gather_full(own):
            pad = torch.zeros(per, dtype=own.dtype, device=dev)
            pad[:own.numel()] = own
            out = torch.empty(per * w, dtype=own.dtype, device=dev)
            dist.all_gather_into_tensor(out, pad)
            return out[:n].clone()

        ref_p = peer.param.clone()
        ref_m, ref_v = gather_full(self.exp_avg[0]), gather_full(self.exp_avg_sq[0])
        self._forward_backward()
        g_sum = self.flat_grad.clone()
        dist.all_reduce(g_sum, op=dist.ReduceOp.SUM)
        found = torch.zeros(1, dtype=torch.float32, device=dev)
        scratch = torch.zeros(2, dtype=torch.int32, device=dev)
        ops.grad_check(g_sum, found, scratch)
        ref_h = torch.empty(n, dtype=torch.float16, device=dev)
        step = int(self.step_dev) - int(self.skipped_dev)
        n_table = self.groups[0][0].params.numel()
        for lo, hi, wd in ((0, n_table, 0.0), (n_table, n, self.wd_net)):
            ops.adam_step(ref_p[lo:hi], g_sum[lo:hi], ref_m[lo:hi], ref_v[lo:hi], ref_h[lo:hi], lr=self.lr,
                          beta1=self.betas[0], beta2=self.betas[1], eps=self.eps, weight_decay=wd, grad_scale_inv=1.0,
                          found_inf=found, step=step)
        self._optimizer()  # the fused exchange (with its two barriers)
        torch.cuda.synchronize()
        self.gather_masters()

        def all_ranks(ok):
            flag = torch.tensor([1.0 if ok else 0.0], device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            return bool(flag.item() == 1.0)

        h0 = peer.param_h.clone()
        dist.broadcast(h0, src=0)
        p0 = peer.param.clone()
        dist.broadcast(p0, src=0)
        identical = all_ranks(torch.equal(h0, peer.param_h) and torch.equal(p0, peer.param))
        b, e = peer.begin, peer.end
        half_ok = all_ranks(torch.equal(peer.param_h[b:e], peer.param[b:e].half()))
        diff = (peer.param - ref_p).abs()
        scale = float(ref_p.abs().max())
        within = float((diff <= 1e-3 * ref_p.abs() + 1e-6).float().mean())
        return {"replicas_identical": identical, "half_consistent": half_ok,
                "max_rel_diff_vs_nccl": float(diff.max()) / scale, "frac_within_1e-3": within,
                "mean_abs_diff": float(diff.mean()), "half_equal_vs_nccl": float((peer.param_h == ref_h).float().mean()),
                "found_inf": float(found), "multicast": bool(peer.multicast),
                "masters": "broadcast" if peer.broadcast_masters else "owner-only + gather", "world": w}

    @property
    def skipped_steps(self):
        """steps skipped because of an fp16 overflow so far (reads the device counter: synchronises)"""
        return int(self.skipped_dev)

    def train_step(self, rays_o, rays_d, direction_norms, gt_rgb, labels, gt_depth):
        self.load_batch(rays_o, rays_d, direction_norms, gt_rgb, labels, gt_depth)
        return self.step()
