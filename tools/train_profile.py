"""k-means++ seeding + Lloyd on 262144 x 768 rows, nlist 64 (for ncu launch lists of the training kernels)."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
pkg = importlib.import_module("cuda-acceleratedvectordatabaseengine_b200")
g = torch.Generator(device="cuda").manual_seed(12345)
x = torch.randn(262144, 768, generator=g, device="cuda")
ix = pkg.IVFFlatIndex(pkg.Config(dimension=768, nlist=64))
ix.train(x)
torch.cuda.synchronize()
