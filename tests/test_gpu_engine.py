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
    total, (lc, ls, ld) = nerf_losses({"image": image, "semantics": sem, "depth": depth}, rgb, labels, gt_depth, 0.6,
                                      global_scale=0.5, fused=False)  # the reference's torch expression
    total.backward()
    loss4 = torch.zeros(4, device=DEV)
    gi, gd, gs = torch.empty(n, 3, device=DEV), torch.empty(n, device=DEV), torch.empty(n, c, device=DEV)
    ops.nerf_loss(image.detach().view(n, 3), depth.detach().view(n), sem.detach().view(n, c), rgb.view(n, 3),
                  labels.view(n), gt_depth.view(n), 0.6, 0.04, 0.1, 0.5, loss4, gi, gd, gs)
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
