"""How exactly do the tensor cores accumulate bf16 x bf16 products?  The bf16 screen of the list scan (csrc/screen.cuh)
budgets ACC_SLACK = 2^-11 |q||v| for it.  Measured here through cuBLAS (same tcgen05 datapath, fp32 accumulate):
max |tc - exact| over |q||v| and over sum|q_i v_i| against a float64 product of the same bf16 values, K = 768 / 1024,
for Gaussian values and for values spread over many binades."""
import torch

torch.manual_seed(0)
for K in (768, 1024):
    for name, gen in (("gaussian", lambda *s: torch.randn(*s, device="cuda")),
                      ("wide", lambda *s: torch.randn(*s, device="cuda") * torch.exp2(torch.randint(-8, 9, s, device="cuda").float())),
                      ("positive", lambda *s: torch.rand(*s, device="cuda") + 0.5)):
        a = gen(4096, K).bfloat16()
        b = gen(512, K).bfloat16()
        tc = torch.mm(a, b.T, out_dtype=torch.float32) if "out_dtype" in (torch.mm.__doc__ or "") else (a @ b.T).float()
        a64, b64 = a.double(), b.double()
        ex = a64 @ b64.T
        err = (tc.double() - ex).abs()
        cs = a64.norm(dim=1)[:, None] * b64.norm(dim=1)[None, :]
        l1 = a64.abs() @ b64.abs().T
        print(f"K={K} {name}: out={tc.dtype} max err/|a||b| = {float((err / cs).max()):.3e}  max err/sum|ab| = {float((err / l1).max()):.3e}"
              f"  (2^-11 = {2**-11:.3e}, 2^-24 = {2**-24:.3e})", flush=True)
