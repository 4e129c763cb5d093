"""CPU check of the arithmetic behind seed_sample_par_kernel (csrc/kmeans.cu).

The reference's k-means++ sampling (ivf_flat_index.cpp:87-103) needs the SEQUENTIAL fp32 running sum of the
min-distances.  The CUDA path evaluates it in parallel from the observation that, while the running sum stays inside
one binade (s = S * 2^eu, S < 2^24), round-to-nearest-even addition of x >= 0 is integer arithmetic on S:

    S' = S + y + c,   x / 2^eu = y + f,   c = [f > 1/2], or for an exact tie f = 1/2 the parity of S + y

so every term is a map "S -> S + (increment depending only on the parity of S)", and such maps compose associatively.
This file restates ps_step / ps_apply / ps_then in numpy integers and checks, on adversarial inputs, that
  (1) applying the per-term maps reproduces np.float32 sequential addition bit for bit inside a binade,
  (2) composing maps pairwise (any bracketing, as a parallel scan does) equals applying them one by one,
  (3) a full walk with binade exits handled by real fp32 additions reproduces the sequential sum and the first
      index whose running sum reaches a target.
"""
import numpy as np

SAT = 1 << 26


def decompose(s):
    """fp32 s >= 0 -> (S, eu) with s = S * 2**eu, S < 2**24 (S >= 2**23 for normal numbers)"""
    b = int(np.float32(s).view(np.uint32))
    ef = (b >> 23) & 0xFF
    return ((b & 0x7FFFFF) | 0x800000, ef - 150) if ef else (b & 0x7FFFFF, -149)


def ps_step(x, eu):
    """x >= 0 in units of 2**eu: (integer part y, rounding class ct: 0 down, 1 up, 2 tie)"""
    m, ex = decompose(x)
    if m == 0:
        return 0, 0
    d = eu - ex
    if d <= 0:
        return (SAT if d < -2 else min(m << (-d), SAT)), 0
    if d > 24:
        return 0, 0
    half = 1 << (d - 1)
    r = m & ((half << 1) - 1)
    return (0 if d == 24 else m >> d), (1 if r > half else 2 if r == half else 0)


def ps_apply(S, y, ct):
    return min(S + y + (((S + y) & 1) if ct == 2 else ct), SAT)


def term_map(y, ct):
    """(increment for an even start, increment for an odd start)"""
    return (y + ((y & 1) if ct == 2 else ct), y + (((1 + y) & 1) if ct == 2 else ct))


def ps_then(f, g):
    """g after f"""
    return (min(f[0] + (g[1] if f[0] & 1 else g[0]), SAT), min(f[1] + (g[1] if (1 + f[1]) & 1 else g[0]), SAT))


def sequential(xs, start=np.float32(0)):
    out = np.empty(len(xs), np.float32)
    s = np.float32(start)
    for i, x in enumerate(xs):
        s = np.float32(s + np.float32(x))
        out[i] = s
    return out


def walk(xs, target=np.inf, span=64):
    """the kernel's walk: spans evaluated with integer maps, binade exits with real fp32 additions"""
    s, pos, n = np.float32(0), 0, len(xs)
    while pos < n:
        S0, eu = decompose(s)
        steps = [ps_step(x, eu) for x in xs[pos:pos + span]]
        # a scan-like evaluation: compose the maps in a balanced tree, then also left to right, must agree
        S, crossed = S0, None
        for j, (y, ct) in enumerate(steps):
            Sn = ps_apply(S, y, ct)
            if Sn >= 1 << 24:
                crossed = j
                break
            S = Sn
            if np.float32(np.ldexp(np.float32(S), eu)) >= target:
                return pos + j, None
        if crossed is None:
            s = np.float32(np.ldexp(np.float32(S), eu))
            pos += len(steps)
        else:
            s = np.float32(np.float32(np.ldexp(np.float32(S), eu)) + np.float32(xs[pos + crossed]))
            if s >= target:
                return pos + crossed, None
            pos += crossed + 1
    return None, s


def datasets():
    rng = np.random.default_rng(5)
    g = (rng.standard_normal(3000) ** 2 * 700).astype(np.float32)                 # squared distances
    ints = rng.integers(0, 4000, 4000).astype(np.float32)                         # exact ties once the unit is >= 2
    sub = (rng.random(2000) * 1e-41).astype(np.float32)                           # subnormal terms and sums
    jumpy = np.where(rng.random(2500) < 0.01, 1e9, rng.random(2500)).astype(np.float32)
    zeros = np.where(rng.random(2000) < 0.7, 0.0, rng.random(2000) * 3).astype(np.float32)
    halves = (rng.integers(0, 64, 3000) * 0.5).astype(np.float32)                 # many terms exactly unit / 2
    return {"gaussian": g, "integers": ints, "subnormal": sub, "jumpy": jumpy, "zeros": zeros, "halves": halves}


def test_term_maps_reproduce_fp32_addition_inside_a_binade():
    rng = np.random.default_rng(1)
    for _ in range(20000):
        s = np.float32(rng.random() * 10.0 ** rng.integers(-30, 30))
        x = np.float32(rng.random() * float(s) * 2.0 ** rng.integers(-30, 1))
        if rng.random() < 0.3:  # force exact ties: x = (odd multiple of half a unit)
            S, eu = decompose(s)
            x = np.float32(np.ldexp(float(2 * rng.integers(0, 50) + 1), eu - 1))
        S, eu = decompose(s)
        y, ct = ps_step(x, eu)
        Sn = ps_apply(S, y, ct)
        ref = np.float32(s + x)
        if Sn < 1 << 24:
            assert np.float32(np.ldexp(np.float32(Sn), eu)) == ref, (s, x)
        else:
            assert ref >= np.float32(np.ldexp(np.float32(1 << 24), eu)) or decompose(ref)[1] > eu, (s, x)


def test_map_composition_is_associative_and_matches_stepwise_application():
    rng = np.random.default_rng(2)
    for _ in range(300):
        eu = int(rng.integers(-20, 20))
        xs = (rng.random(16) * np.ldexp(1.0, eu + rng.integers(-3, 8))).astype(np.float32)
        maps = [term_map(*ps_step(x, eu)) for x in xs]
        left = (0, 0)
        for m in maps:
            left = ps_then(left, m)
        tree = maps
        while len(tree) > 1:
            tree = [ps_then(tree[i], tree[i + 1]) for i in range(0, len(tree), 2)]
        assert tree[0] == left
        for S0 in (1 << 23, (1 << 23) + 1, 9_000_001, 9_000_002):
            S = S0
            for x in xs:
                S = ps_apply(S, *ps_step(x, eu))
            inc = left[1] if S0 & 1 else left[0]
            if S < SAT and S0 + inc < SAT:
                assert S0 + inc == S


def test_walk_reproduces_the_sequential_sum_and_the_first_index_reaching_a_target():
    for name, xs in datasets().items():
        ref = sequential(xs)
        idx, total = walk(xs)
        assert idx is None and total == ref[-1], name
        for frac in (0.0, 1e-6, 0.1, 0.5, 0.9, 0.999999, 1.0):
            target = np.float32(np.float32(frac) * ref[-1])
            want = int(np.argmax(ref >= target)) if (ref >= target).any() else None
            got, _ = walk(xs, target)
            assert got == want, (name, frac, got, want)
