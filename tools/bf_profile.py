"""One configs[1] brute-force call (1M x 768, 1024 queries, k = 100) for ncu launch lists."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
pkg = importlib.import_module("cuda-acceleratedvectordatabaseengine_b200")
n, dim, nq, k = 1_000_000, 768, 1024, 100
g = torch.Generator(device="cuda").manual_seed(7)
db = torch.randn(n, dim, generator=g, device="cuda")
q = torch.randn(nq, dim, generator=g, device="cuda")
for _ in range(2):
    pkg.bruteforce_search(db, q, k)
torch.cuda.synchronize()
