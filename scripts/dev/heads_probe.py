"""Bring-up probe: where does ucsa_heads_fwd spend its time?  Times the call with / without the saved activations and
with / without the fused compositing at config-2 sizes.   python scripts/heads_probe.py"""
import sys, torch
sys.path.insert(0, ".")
from ucsa_neural_rendering_b200 import ops
from ucsa_neural_rendering_b200.nerf import SemanticNeRFNetwork
dev = torch.device("cuda")
n, t, c = 4096, 512, 40
net = SemanticNeRFNetwork(encoding="hashgrid", bound=4, cuda_ray=False, density_scale=1, num_semantic_classes=c).to(dev)
g = torch.Generator(device=dev).manual_seed(0)
k_max = n * t
k = int(0.93 * k_max)
cnt = torch.full((n,), k // n, dtype=torch.int32, device=dev)
off = torch.zeros(n + 1, dtype=torch.int32, device=dev); ops.scan_counts(cnt, off)
k = int(off[-1])
per = k // n
sel = (torch.arange(n, device=dev).view(n, 1) * t + torch.arange(per, device=dev).view(1, per)).reshape(-1).int().contiguous()
sel = torch.cat([sel, torch.zeros(k_max - k, dtype=torch.int32, device=dev)])
w_sel = torch.rand(k_max, device=dev, generator=g) * 0.01
d = torch.nn.functional.normalize(torch.randn(n, 3, device=dev, generator=g), dim=-1)
h = (torch.randn(n, t, 16, device=dev, generator=g) * 0.3).half()
w_col, w_sem = net.color_net.half_params(), net.semantics_net.half_params()
rgb = torch.empty(k_max, 3, device=dev)
rows = ops.tile_rows(k_max)
hc1, hc2, hs = (torch.empty(rows, 64, dtype=torch.float16, device=dev) for _ in range(3))
image, sem = torch.zeros(n, 3, device=dev), torch.zeros(n, c, device=dev)
flush = torch.empty(64 * 1024 * 1024, device=dev)
def timeit(fn):
    for _ in range(2): fn()
    ts = []
    for _ in range(7):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]
for name, kw in (("save+composite", dict(hc1=hc1, hc2=hc2, hs=hs, w_sel=w_sel, image=image, semantics=sem)),
                 ("save only", dict(hc1=hc1, hc2=hc2, hs=hs)),
                 ("composite only", dict(w_sel=w_sel, image=image, semantics=sem)),
                 ("neither", dict())):
    ms = timeit(lambda: ops.heads_fwd(sel, off, n, t, k_max, d, h, w_col, w_sem, c, rgb, None, **kw))
    print(f"heads_fwd {name:16s} {ms:.3f} ms  (K = {k})")
