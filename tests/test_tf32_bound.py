"""CPU check of the error bound that makes the tensor-core paths exact (coarse.cu, assign_tc.cu, bruteforce_tc.cu).

Those kernels feed fp32 rows to tcgen05 `kind::tf32` (the MMA reads the top 19 bits of every operand: 10 explicit
mantissa bits) and decide from the approximate dot product which candidates need an exact fp32 re-check.  The decision
is sound iff |true dot - tf32 dot| <= E with E = 1.05 * 2^-8 * |x| * |c| (plus (dim + 16) * 2^-24 of the magnitudes
for the fp32 accumulation).  Here TF32 operands are emulated both ways the hardware may convert (truncation and
round-to-nearest) and the bound is checked on random and adversarial vectors.
"""
import numpy as np

E_COEF = 1.05 * 2.0 ** -8


def tf32(a, mode):
    b = np.ascontiguousarray(a, np.float32).view(np.uint32)
    if mode == "truncate":
        return (b & np.uint32(0xFFFFE000)).view(np.float32)
    return ((b + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)  # round half up on the dropped bits


def cases():
    rng = np.random.default_rng(3)
    for dim in (16, 100, 768, 2048):
        yield rng.standard_normal((64, dim)).astype(np.float32), rng.standard_normal((64, dim)).astype(np.float32)
        # same sign everywhere, mantissas just below the truncation boundary: the worst case for truncation
        worst = np.float32(1.0) + np.float32(2.0 ** -10) - np.float32(2.0 ** -23)
        yield np.full((4, dim), worst, np.float32), np.full((4, dim), worst, np.float32)
        # wide dynamic range
        yield ((rng.standard_normal((64, dim)) * 10.0 ** rng.integers(-3, 4, (64, dim))).astype(np.float32),
               (rng.standard_normal((64, dim)) * 10.0 ** rng.integers(-3, 4, (64, dim))).astype(np.float32))


def test_tf32_dot_product_stays_inside_the_bound():
    for x, c in cases():
        dim = x.shape[1]
        true = np.einsum("nd,md->nm", x.astype(np.float64), c.astype(np.float64))
        xn = np.linalg.norm(x.astype(np.float64), axis=1)[:, None]
        cn = np.linalg.norm(c.astype(np.float64), axis=1)[None, :]
        bound = (E_COEF + (dim + 16) * 2.0 ** -24) * xn * cn
        for mode in ("truncate", "nearest"):
            approx = np.einsum("nd,md->nm", tf32(x, mode), tf32(c, mode), dtype=np.float32).astype(np.float64)
            err = np.abs(true - approx)
            assert (err <= bound).all(), (dim, mode, float((err / bound).max()))
            assert (err / bound).max() < 0.6 or dim <= 16  # and it is not tight by accident: real slack remains


def test_l2_score_bound_covers_the_exact_fp32_kernel():
    """the L2 admission tests compare |v|^2 + |q|^2 - 2 dot against distances the exact kernels computed in fp32"""
    rng = np.random.default_rng(4)
    dim = 768
    v = rng.standard_normal((256, dim)).astype(np.float32)
    q = rng.standard_normal((32, dim)).astype(np.float32)
    exact32 = np.zeros((32, 256), np.float32)
    for d in range(dim):  # ascending-dimension fp32 sum, like the reference / the exact kernels
        diff = (q[:, d][:, None] - v[:, d][None, :]).astype(np.float32)
        exact32 = (exact32 + diff * diff).astype(np.float32)
    v2 = (v.astype(np.float64) ** 2).sum(1)[None, :]
    q2 = (q.astype(np.float64) ** 2).sum(1)[:, None]
    dot = np.einsum("nd,md->nm", tf32(q, "truncate"), tf32(v, "truncate"), dtype=np.float32).astype(np.float64)
    score = v2 + q2 - 2.0 * dot
    eps = (dim + 16) * 2.0 ** -24
    e = (E_COEF + 2 * eps) * np.sqrt(v2) * np.sqrt(q2) + eps * (v2 + q2)
    assert (np.abs(score - exact32.astype(np.float64)) <= e).all()
