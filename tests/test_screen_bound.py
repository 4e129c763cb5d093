"""The error bound behind the tensor-core screen of the list scan (csrc/screen.cuh), checked in numpy on the CPU.

The screen streams a low-precision shadow v~ of every row (bf16, or int8 with one scale per row) and multiplies it with a
shadow q~ of the query (bf16, or TWO int8 terms).  It may drop a (row, query) pair only if a LOWER BOUND of the exact
distance already exceeds the query's running k-th distance, and the lower bound rests on

    q.v - q~.v~ = (q - q~).v~ + q.(v - v~)   =>   |q.v - q~.v~| <= e_q (|v| + e_v) + |q| e_v,

with e_q = |q - q~| and e_v = |v - v~| stored per query / per row (computed exactly when the shadow is written).
These tests restate the shadow construction of kmeans.cu (mirror_write_row) and screen.cuh (query_image_kernel) and
check the inequality in float64 on Gaussian, heavy-tailed, offset and adversarial inputs -- adversarial meaning the
row's rounding error is ALIGNED with the query, which is when Cauchy-Schwarz is tight."""
import zlib

import numpy as np
import pytest


def bf16(x):
    """round-to-nearest-even to bfloat16, returned as float32 (what __floats2bfloat162_rn does)"""
    b = np.ascontiguousarray(x, np.float32).view(np.uint32).astype(np.uint64)
    b = (b + 0x7FFF + ((b >> 16) & 1)) & 0xFFFF0000
    return b.astype(np.uint32).view(np.float32)


def int8_row(v):
    """mirror_write_row, int8: scale = max|v| / 127, n = clamp(rint(v / scale)); returns (n, scale)"""
    mx = np.abs(v).max(axis=-1, keepdims=True).astype(np.float32)
    scale = (mx * np.float32(1.0 / 127.0)).astype(np.float32)
    inv = np.where(scale > 0, np.float32(1.0) / np.where(scale > 0, scale, 1), 0).astype(np.float32)
    n = np.clip(np.rint(v.astype(np.float32) * inv), -127, 127)
    return n, scale


def int8_query(q):
    """query_image_kernel, int8: q ~ sa a + sb b with b the quantised residual of the first term"""
    a, sa = int8_row(q)
    r = (q.astype(np.float64) - sa.astype(np.float64) * a).astype(np.float32)  # fmaf(-sa, a, q): one rounding
    b, sb = int8_row(r)
    return a, sa, b, sb


CASES = {
    "gaussian768": lambda rng, n, d: rng.standard_normal((n, d)),
    "heavy_tails": lambda rng, n, d: rng.standard_t(2.0, (n, d)),
    "offset_cloud": lambda rng, n, d: rng.standard_normal((n, d)) * 0.01 + 100.0,
    "many_binades": lambda rng, n, d: rng.standard_normal((n, d)) * np.exp2(rng.integers(-12, 13, (n, d))),
    "sparse": lambda rng, n, d: rng.standard_normal((n, d)) * (rng.random((n, d)) < 0.05),
}


@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("kind", ["bf16", "int8"])
def test_dot_product_error_is_within_the_stored_bound(name, kind):
    rng = np.random.default_rng(zlib.crc32(f"{name}/{kind}".encode()))
    d = 768
    v = CASES[name](rng, 400, d).astype(np.float32)
    q = CASES[name](rng, 64, d).astype(np.float32)
    # adversarial queries: aligned with the rounding error of a row, and with the row itself
    if kind == "bf16":
        vs = bf16(v).astype(np.float64)
        qs = bf16(q)
    else:
        n, s = int8_row(v)
        vs = s.astype(np.float64) * n
        qs = None
    err_rows = v.astype(np.float64) - vs
    q = np.concatenate([q, (err_rows[:32] * 1e3).astype(np.float32), v[:32]])
    if kind == "bf16":
        qs = bf16(q).astype(np.float64)
    else:
        a, sa, b, sb = int8_query(q)
        qs = sa.astype(np.float64) * a + sb.astype(np.float64) * b
    q64, v64 = q.astype(np.float64), v.astype(np.float64)
    e_q = np.linalg.norm(q64 - qs, axis=1)
    e_v = np.linalg.norm(err_rows, axis=1)
    nq, nv = np.linalg.norm(q64, axis=1), np.linalg.norm(v64, axis=1)
    exact = q64 @ v64.T
    shadow = qs @ vs.T   # what the tensor cores compute (bf16: up to accumulation; int8: exactly, in integers)
    bound = e_q[:, None] * (nv + e_v)[None, :] + nq[:, None] * e_v[None, :]
    slack = bound - np.abs(exact - shadow)
    assert (slack >= -1e-9 * (nq[:, None] * nv[None, :] + 1e-30)).all(), float(slack.min())
    # and the bound is not vacuous: it is a small fraction of |q||v| (2^-8 for bf16, about 1 % for int8 on Gaussian rows)
    rel = bound / (nq[:, None] * nv[None, :] + 1e-300)
    if name == "gaussian768":
        assert np.median(rel) < (0.004 if kind == "bf16" else 0.012), float(np.median(rel))


def test_int8_integer_dot_products_fit_their_types():
    """|sum of 1024 products of int8 values| <= 1024 * 127^2 = 16 516 096 < 2^24: the int32 accumulator is exact and its
    conversion to fp32 (the screen multiplies it by the row and query scales) is exact too"""
    assert 1024 * 127 * 127 < 2**24 < 2**31


def test_two_term_int8_query_is_second_order():
    rng = np.random.default_rng(5)
    q = rng.standard_normal((64, 768)).astype(np.float32)
    a, sa, b, sb = int8_query(q)
    one = np.linalg.norm(q - sa * a, axis=1) / np.linalg.norm(q, axis=1)
    two = np.linalg.norm(q.astype(np.float64) - (sa.astype(np.float64) * a + sb.astype(np.float64) * b), axis=1) / np.linalg.norm(q, axis=1)
    assert one.max() < 0.02 and two.max() < 2e-4, (one.max(), two.max())


def test_lower_bound_never_exceeds_the_exact_distance():
    """the screen's test as evaluated in fp32 (screen.cuh, consumers): L2 and inner product, both shadows"""
    rng = np.random.default_rng(11)
    d = 768
    v = rng.standard_normal((2000, d)).astype(np.float32)
    q = rng.standard_normal((16, d)).astype(np.float32)
    DOT_SLACK, ACC = np.float32(2e-5), {"bf16": np.float32(2.0**-14), "int8": np.float32(2.0**-17)}
    for kind in ("bf16", "int8"):
        if kind == "bf16":
            vs, qs = bf16(v).astype(np.float64), bf16(q).astype(np.float64)
        else:
            n, s = int8_row(v)
            vs = s.astype(np.float64) * n
            a, sa, b, sb = int8_query(q)
            qs = sa.astype(np.float64) * a + sb.astype(np.float64) * b
        e_v = (np.linalg.norm(v.astype(np.float64) - vs, axis=1) * 1.0002).astype(np.float32)
        e_q = (np.linalg.norm(q.astype(np.float64) - qs, axis=1) * 1.0002).astype(np.float32)
        vn = (v.astype(np.float32) ** 2).sum(1, dtype=np.float32)
        qn = (q.astype(np.float32) ** 2).sum(1, dtype=np.float32)
        nv = (np.sqrt(vn) * np.float32(1.00001)).astype(np.float32)
        nq = (np.sqrt(qn) * np.float32(1.0002)).astype(np.float32)
        dot = (qs @ vs.T).astype(np.float32)
        uc = (2 * e_q + 2 * ACC[kind] * nq).astype(np.float32)
        wc = (2 * (nq + e_q)).astype(np.float32)
        slack = nv[None, :] * uc[:, None] + e_v[None, :] * wc[:, None]
        lb_l2 = (vn[None, :] * (1 - DOT_SLACK) + qn[:, None] * (1 - DOT_SLACK)) - 2 * dot - slack
        lb_ip = -dot - np.float32(0.5) * slack
        exact_l2 = ((q.astype(np.float64)[:, None, :] - v.astype(np.float64)[None, :, :]) ** 2).sum(-1)
        exact_ip = -(q.astype(np.float64) @ v.astype(np.float64).T)
        assert (lb_l2 <= exact_l2).all() and (lb_ip <= exact_ip).all()
        # tight enough to be useful: the L2 bound sits within 1 % (bf16) / 3 % (int8) of the distance on this data
        assert np.median((exact_l2 - lb_l2) / exact_l2) < (0.01 if kind == "bf16" else 0.03)
