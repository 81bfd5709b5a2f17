"""TrainEngine (graph-captured training step) and the fused loss kernel against the autograd path / torch.  GPU only."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"
MIX = 0x9E3779B97F4A7C15


def _batch(n, c, seed):
    g = torch.Generator().manual_seed(seed)
    o = ((torch.rand(1, n, 3, generator=g) - 0.5)).to(DEV)
    d = torch.nn.functional.normalize(torch.randn(1, n, 3, generator=g), dim=-1).to(DEV)
    dn = (1 + 0.2 * torch.rand(1, n, 1, generator=g)).to(DEV)
    rgb = torch.rand(1, n, 3, generator=g).half().to(DEV)
    labels = torch.randint(-1, c, (1, n), generator=g).to(DEV)
    depth = (torch.rand(1, n, generator=g) * 3).to(DEV)
    depth[0, ::7] = 0  # invalid depth pixels
    return o, d, dn, rgb, labels, depth


def _net(seed_amp=0.3):
    from ucsa_neural_rendering_b200 import build
    from ucsa_neural_rendering_b200.nerf import SemanticNeRFNetwork

    build.build_library()
    net = SemanticNeRFNetwork(encoding="hashgrid", bound=4, cuda_ray=False, density_scale=1, num_semantic_classes=40)
    with torch.no_grad():
        g = torch.Generator().manual_seed(1)
        net.encoder.params.copy_((torch.rand(net.encoder.params.numel(), generator=g) * 2 - 1) * seed_amp)
    return net.to(DEV).train()


def test_loss_kernel_matches_reference_losses():
    from ucsa_neural_rendering_b200 import ops
    from ucsa_neural_rendering_b200.trainer import nerf_losses

    n, c = 1000, 40
    g = torch.Generator().manual_seed(0)
    image = torch.rand(1, n, 3, generator=g).to(DEV).requires_grad_()
    depth = (torch.rand(1, n, generator=g) * 4).to(DEV).requires_grad_()
    sem = torch.rand(1, n, c, generator=g)
    sem[0, :5] = 0  # rays without any semantic mass: "invalid" -> ignored
    sem = sem.to(DEV).requires_grad_()
    _, _, _, rgb, labels, gt_depth = _batch(n, c, 3)
    from oracle.losses import nerf_losses as oracle_losses  # the reference's torch expression

    total, (lc, ls, ld) = oracle_losses({"image": image, "semantics": sem, "depth": depth}, rgb, labels, gt_depth, 0.6,
                                        global_scale=0.5)
    total.backward()
    loss4 = torch.zeros(4, device=DEV)
    gi, gd, gs = torch.empty(n, 3, device=DEV), torch.empty(n, device=DEV), torch.empty(n, c, device=DEV)
    ops.nerf_loss(image.detach().view(n, 3), depth.detach().view(n), sem.detach().view(n, c), rgb.view(n, 3),
                  labels.view(n), gt_depth.view(n), 0.6, 0.04, 0.1, 0.5, loss4, gi, gd, gs, ops.loss_scratch(DEV))
    torch.testing.assert_close(loss4, torch.stack([total.detach(), lc.detach(), ls.detach(), ld.detach()]), rtol=1e-5,
                               atol=1e-6)
    torch.testing.assert_close(gi, image.grad.view(n, 3), rtol=1e-5, atol=1e-9)
    torch.testing.assert_close(gd, depth.grad.view(n), rtol=1e-5, atol=1e-9)
    torch.testing.assert_close(gs, sem.grad.view(n, c), rtol=1e-4, atol=1e-9)
    # the autograd wrapper nerf_losses uses on CUDA tensors, with an upstream factor (GradScaler)
    img2, dep2, sem2 = (x.detach().clone().requires_grad_() for x in (image, depth, sem))
    total2, parts2 = nerf_losses({"image": img2, "semantics": sem2, "depth": dep2}, rgb, labels, gt_depth, 0.6,
                                 global_scale=0.5)
    (total2 * 8.0).backward()
    torch.testing.assert_close(total2.detach(), total.detach(), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(torch.stack(list(parts2)), torch.stack([lc, ls, ld]).detach(), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(img2.grad, 8.0 * image.grad, rtol=1e-5, atol=1e-9)
    torch.testing.assert_close(dep2.grad, 8.0 * depth.grad, rtol=1e-5, atol=1e-9)
    torch.testing.assert_close(sem2.grad, 8.0 * sem.grad, rtol=1e-4, atol=1e-9)


def test_engine_step_matches_autograd_path():
    from ucsa_neural_rendering_b200.engine import TrainEngine
    from ucsa_neural_rendering_b200.trainer import nerf_losses

    n, c = 512, 40
    net = _net()
    batch = _batch(n, c, 5)
    eng = TrainEngine(net, n, num_steps=32, upsample_steps=32, one_m_to_scene_uom=0.6, seed=77, use_graph=False)
    eng.load_batch(*batch)
    eng._forward_backward()  # step counter 0 -> 1
    torch.cuda.synchronize()
    flat = eng.flat_grad.clone()
    loss_engine = eng.loss.clone()
    for m, _ in eng.groups:
        m.params.grad = None
    o, d, dn, rgb, labels, depth = batch
    out = net.render(o, d, direction_norms=dn, perturb=True, num_steps=32, upsample_steps=32,
                     seed=(77 + MIX * 1) % (1 << 64))
    total, parts = nerf_losses(out, rgb, labels, depth, 0.6)
    total.backward()
    torch.testing.assert_close(loss_engine[0], total.detach(), rtol=1e-4, atol=1e-6)
    ref = torch.cat([m.params.grad.view(-1) for m, _ in eng.groups])
    torch.testing.assert_close(flat, ref, rtol=2e-2, atol=2e-3 * float(ref.abs().max()))


def test_graph_replay_equals_eager_and_trains():
    from ucsa_neural_rendering_b200.engine import TrainEngine

    n, c = 512, 40
    batches = [_batch(n, c, 10 + i) for i in range(4)]
    finals, losses = [], []
    for use_graph in (False, True):
        net = _net()
        eng = TrainEngine(net, n, num_steps=32, upsample_steps=32, one_m_to_scene_uom=0.6, seed=5, use_graph=use_graph)
        hist = []
        for i in range(8):
            hist.append(float(eng.train_step(*batches[i % 4])[0]))
        torch.cuda.synchronize()
        finals.append([m.params.detach().clone() for m, _ in eng.groups])
        losses.append(hist)
        assert int(eng.step_dev) == 8
    # Same kernels and seeds; only the atomic summation order differs between the two runs.  Adam with eps = 1e-15
    # turns that noise into +-lr steps on entries whose gradient is itself noise, so compare what training sees:
    # the loss trajectory (and the bulk of the parameters), not every table entry.
    assert abs(losses[0][0] - losses[1][0]) < 1e-4 * abs(losses[0][0])
    for a, b in zip(losses[0], losses[1]):
        assert abs(a - b) < 3e-2 * abs(a), (losses[0], losses[1])
    for a, b in zip(*finals):
        close = ((a - b).abs() <= 1e-3 + 1e-2 * b.abs()).float().mean()
        assert float(close) > 0.9, float(close)
    assert losses[1][-1] < losses[1][0], losses[1]


def test_adam_exchange_kernel_equals_adam_step_on_one_rank():
    """ucsa_adam_exchange with world = 1 (peer loads / stores on the rank's own buffers) must reproduce
    ucsa_adam_step on each parameter group: same arithmetic, weight decay only past wd_begin, owned slice only."""
    from types import SimpleNamespace

    from ucsa_neural_rendering_b200 import ops

    g = torch.Generator().manual_seed(4)
    n, wd_begin = 40000, 30000
    p0 = torch.randn(n, generator=g).to(DEV)
    grad = torch.randn(n, generator=g).to(DEV)
    step_dev = torch.tensor([3], dtype=torch.int32, device=DEV)
    hyper = dict(lr=1e-2, beta1=0.9, beta2=0.99, eps=1e-15)
    # reference: two groups through ucsa_adam_step
    ref_p, ref_h = p0.clone(), torch.empty(n, dtype=torch.float16, device=DEV)
    m_ref, v_ref = torch.rand(n, generator=g).to(DEV) * 0.1, torch.rand(n, generator=g).to(DEV) * 0.01
    m0, v0 = m_ref.clone(), v_ref.clone()
    for lo, hi, wd in ((0, wd_begin, 0.0), (wd_begin, n, 1e-3)):
        ops.adam_step(ref_p[lo:hi], grad[lo:hi], m_ref[lo:hi], v_ref[lo:hi], ref_h[lo:hi], weight_decay=wd,
                      grad_scale_inv=1.0, found_inf=None, step=1, step_dev=step_dev, **hyper)
    # the same two groups as ONE launch over the flat buffer (weight decay from wd_begin on): what TrainEngine issues
    one_p, one_h = p0.clone(), torch.empty(n, dtype=torch.float16, device=DEV)
    one_m, one_v = m0.clone(), v0.clone()
    ops.adam_step(one_p, grad, one_m, one_v, one_h, weight_decay=1e-3, wd_begin=wd_begin, grad_scale_inv=1.0,
                  found_inf=None, step=1, step_dev=step_dev, **hyper)
    assert torch.equal(one_p, ref_p) and torch.equal(one_h, ref_h) and torch.equal(one_m, m_ref) and torch.equal(one_v, v_ref)
    # fused exchange on the slice [8000, 36000) of a one-rank "world"
    begin, end = 8000, 36000
    p, h = p0.clone(), torch.zeros(n, dtype=torch.float16, device=DEV)
    peer = SimpleNamespace(world=1, rank=0, begin=begin, end=end, multicast=False, grad_ptrs=[grad.data_ptr()],
                           param_ptrs=[p.data_ptr()], param_h_ptrs=[h.data_ptr()], mc_grad=0, mc_param=0, mc_param_h=0)
    m, v = m0[begin:end].clone(), v0[begin:end].clone()
    ops.adam_exchange(peer, m, v, wd_begin=wd_begin, weight_decay=1e-3, step=1, step_dev=step_dev, **hyper)
    torch.cuda.synchronize()
    assert torch.equal(p[begin:end], ref_p[begin:end]) and torch.equal(h[begin:end], ref_h[begin:end])
    assert torch.equal(m, m_ref[begin:end]) and torch.equal(v, v_ref[begin:end])
    assert torch.equal(p[:begin], p0[:begin]) and torch.equal(p[end:], p0[end:])  # outside the owned slice: untouched


def test_adam_step_matches_torch_adam_under_gradscaler():
    """ucsa_grad_check + ucsa_adam_step against torch.optim.Adam driven by torch.amp.GradScaler, the optimizer of
    joint_train_lightning_net.py:46,509-513,897-919: lr 1e-2, betas (0.9, 0.99), eps 1e-15, two groups with weight
    decay 1e-6 on the MLP group only; 10 steps, one of which carries an inf gradient and must be skipped (no moment /
    parameter change, no bias-correction advance), after which GradScaler halves its scale."""
    from ucsa_neural_rendering_b200 import ops

    g = torch.Generator().manual_seed(11)
    sizes = (100_003, 14_336)  # "encoding" (odd length: exercises the scalar tail), "net"
    wds = (0.0, 1e-6)
    p0 = [(torch.rand(sizes[0], generator=g) * 2e-4 - 1e-4), torch.randn(sizes[1], generator=g) * 0.3]
    ref_p = [torch.nn.Parameter(p.clone().to(DEV)) for p in p0]
    opt = torch.optim.Adam([{"params": [ref_p[0]]}, {"params": [ref_p[1]], "weight_decay": wds[1]}], lr=1e-2,
                           betas=(0.9, 0.99), eps=1e-15)
    scaler = torch.amp.GradScaler("cuda", init_scale=128.0, growth_interval=10_000)
    our_p = [p.clone().to(DEV) for p in p0]
    our_h = [torch.empty(n, dtype=torch.float16, device=DEV) for n in sizes]
    our_m = [torch.zeros(n, device=DEV) for n in sizes]
    our_v = [torch.zeros(n, device=DEV) for n in sizes]
    # one flat gradient buffer like the engine's; views must stay 16-byte aligned -> pad the first group
    pad0 = (sizes[0] + 3) // 4 * 4
    flat = torch.zeros(pad0 + sizes[1], device=DEV)
    views = [flat[:sizes[0]], flat[pad0:pad0 + sizes[1]]]
    step_dev = torch.zeros(1, dtype=torch.int32, device=DEV)
    skipped = torch.zeros(1, dtype=torch.int32, device=DEV)
    found = torch.zeros(1, device=DEV)
    scratch = torch.zeros(2, dtype=torch.int32, device=DEV)
    inf_step = 4
    for it in range(10):
        scale = float(scaler.get_scale())
        true_g = [torch.randn(n, generator=g) * (10.0 ** float(torch.randint(-6, 1, (1,), generator=g))) for n in sizes]
        true_g[0][::7] = 0.0  # untouched hash entries
        scaled = [(t * scale).to(DEV) for t in true_g]
        if it == inf_step:
            scaled[0][12345] = float("inf")
        # reference: GradScaler.unscale_ + step + update
        scaler.scale(torch.zeros(1, device=DEV))  # initialises the scaler's state like a scaled backward would
        for p, sg in zip(ref_p, scaled):
            p.grad = sg.clone()
        scaler.step(opt)
        scaler.update()
        # ours
        flat.zero_()
        for v, sg in zip(views, scaled):
            v.copy_(sg)
        step_dev.add_(1)
        before = [t.clone() for t in our_p + our_m + our_v]
        ops.grad_check(flat, found, scratch, skipped_dev=skipped)
        for i in range(2):
            ops.adam_step(our_p[i], views[i], our_m[i], our_v[i], our_h[i], lr=1e-2, beta1=0.9, beta2=0.99, eps=1e-15,
                          weight_decay=wds[i], grad_scale_inv=1.0 / scale, found_inf=found, step=1, step_dev=step_dev,
                          skipped_dev=skipped)
        torch.cuda.synchronize()
        if it == inf_step:
            assert float(found) == 1.0 and int(skipped) == 1
            for a, b in zip(before, our_p + our_m + our_v):
                assert torch.equal(a, b), "a skipped step must not touch parameters or moments"
            assert float(scaler.get_scale()) == scale * 0.5  # the reference backed off
        else:
            assert float(found) == 0.0
        assert int(scratch[0]) == 0 and int(scratch[1]) == 0, "the check re-arms its scratch"
    assert int(step_dev) - int(skipped) == 9
    for i in range(2):
        ref = ref_p[i].detach()
        st = opt.state[ref_p[i]]
        assert int(st["step"]) == 9
        rel = ((our_p[i] - ref).abs() / ref.abs().clamp_min(1e-12))
        # the same float operations in the same order; what is left is fused-multiply-add contraction inside torch's
        # kernels.  atol = 1e-6 x the step size (lr = 1e-2): parameters that cancelled to ~0 over the 9 updates carry
        # the absolute error of the updates, not a relative one
        torch.testing.assert_close(our_p[i], ref, rtol=1e-6, atol=1e-8, msg=lambda m: f"group {i}: {m}; max rel {float(rel.max()):.3e}")
        torch.testing.assert_close(our_m[i], st["exp_avg"], rtol=1e-6, atol=1e-12)
        torch.testing.assert_close(our_v[i], st["exp_avg_sq"], rtol=1e-6, atol=1e-20)
        assert torch.equal(our_h[i], our_p[i].half()), "fp16 working copy = rounded master"


def test_fused_adam_optimizer_is_torch_adam_under_gradscaler():
    """ucsa_neural_rendering_b200.optim.FusedAdam (torch.optim interface, `_step_supports_amp_scaling`) against
    torch.optim.Adam, both driven by torch.amp.GradScaler exactly like training_step_nerf does
    (joint_train_lightning_net.py:509-513): scale(loss).backward(); scaler.step(opt); scaler.update().  Eight steps on
    the real network, one of them with a poisoned target (inf loss -> skipped by both)."""
    import copy

    from ucsa_neural_rendering_b200.optim import FusedAdam

    net_a = _net()
    net_b = copy.deepcopy(net_a)

    def groups(net):
        return [{"name": "encoding", "params": list(net.encoder.parameters())},
                {"name": "net", "params": list(net.sigma_net.parameters()) + list(net.color_net.parameters())
                 + list(net.semantics_net.parameters()), "weight_decay": 1e-6}]

    opt_a = torch.optim.Adam(groups(net_a), lr=1e-2, betas=(0.9, 0.99), eps=1e-15)
    opt_b = FusedAdam(groups(net_b), lr=1e-2, betas=(0.9, 0.99), eps=1e-15, network=net_b)
    scal_a, scal_b = torch.amp.GradScaler("cuda", init_scale=1024.0), torch.amp.GradScaler("cuda", init_scale=1024.0)
    n = 128
    o, d, dn, rgb, labels, depth = _batch(n, 40, 21)
    for it in range(8):
        target = rgb.float().clone()
        if it == 3:
            target[0, 0, 0] = float("inf")
        # one backward pass (network A); network B receives the very same scaled gradients, so that only the optimizers
        # differ (two backward passes would differ in the order of their atomics, which Adam with eps = 1e-15 amplifies)
        opt_a.zero_grad()
        with torch.autocast("cuda", enabled=True):
            out = net_a.render(o, d, direction_norms=dn, perturb=True, num_steps=32, upsample_steps=32, seed=100 + it)
            loss = ((out["image"] - target) ** 2).mean() + 0.1 * out["depth"].mean()
        scal_a.scale(loss).backward()
        scal_b.scale(torch.zeros(1, device=DEV))  # initialises scaler B like a scaled backward would
        for pa, pb in zip(net_a.parameters(), net_b.parameters()):
            pb.grad = pa.grad.clone()
        scal_a.step(opt_a)
        scal_a.update()
        scal_b.step(opt_b)
        scal_b.update()
    assert float(scal_a.get_scale()) == float(scal_b.get_scale()) == 512.0  # both backed off once
    assert int(opt_b.state[net_b.encoder.params]["step"]) == 7 == int(opt_a.state[net_a.encoder.params]["step"])
    for (name, pa), (_, pb) in zip(net_a.named_parameters(), net_b.named_parameters()):
        # same gradients, same operations in the same order: torch's own kernels may contract differently (fma)
        torch.testing.assert_close(pb.detach(), pa.detach(), rtol=1e-6, atol=1e-8, msg=lambda m, name=name: f"{name}: {m}")
    # the fp16 working copies follow the masters without a separate cast
    for m in (net_b.encoder, net_b.sigma_net, net_b.color_net, net_b.semantics_net):
        assert torch.equal(m.half_params(), m.params.detach().half())
