"""bring-up helper: run the tcgen05 MLP forward for one (shape, n) in isolation"""
import sys, torch
sys.path.insert(0, ".")
from ucsa_neural_rendering_b200 import ops
shape = {"sigma": [32, 64, 16], "color": [32, 64, 64, 16], "sem": [16, 64, 48]}[sys.argv[1]]
n = int(sys.argv[2]); pre = sys.argv[3] if len(sys.argv) > 3 else ""
dev = "cuda"
if pre == "density":
    grid = ops.make_grid_desc(4)
    tab = torch.zeros(grid.total_entries * 2, dtype=torch.float16, device=dev)
    w = torch.zeros(3072, dtype=torch.float16, device=dev)
    xyz = torch.rand(5000, 3, device=dev)
    ops.density_fwd(grid, tab, w, 4.0, xyz=xyz, sigma=torch.empty(5000, device=dev), h=torch.empty(5000, 16, dtype=torch.float16, device=dev))
    torch.cuda.synchronize(); print("density ok", flush=True)
x = torch.randn(n, shape[0], device=dev).half()
w = (torch.randn(sum(a * b for a, b in zip(shape[:-1], shape[1:])), device=dev) * 0.2).half()
y = torch.empty(n, shape[-1], dtype=torch.float16, device=dev)
acts = torch.empty(n, sum(shape[1:-1]), dtype=torch.float16, device=dev)
ops.mlp_fwd(x, w, shape, y, acts)
torch.cuda.synchronize()
y2 = torch.empty_like(y); a2 = torch.empty_like(acts)
ops.mlp_fwd(x, w, shape, y2, a2, simt=True)
torch.cuda.synchronize()
print(sys.argv[1:], "ok maxdiff", float((y.float() - y2.float()).abs().max()), flush=True)
