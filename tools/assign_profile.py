"""One nearest-centroid assignment (1M x 768 rows, 16384 centroids) for ncu launch lists."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
pkg = importlib.import_module("cuda-acceleratedvectordatabaseengine_b200")
n, dim, nlist = 1_000_000, 768, 16384
g = torch.Generator(device="cuda").manual_seed(3)
x = torch.randn(n, dim, generator=g, device="cuda")
ix = pkg.IVFFlatIndex(pkg.Config(dimension=dim, nlist=nlist))
ix.centroids = torch.randn(nlist, dim, generator=g, device="cuda").cpu().numpy()
for _ in range(2):
    ix.assign_device(x)
torch.cuda.synchronize()
