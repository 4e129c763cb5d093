"""Parity checks between the CUDA path and the oracle.

The bar (BASELINE.json north_star): neighbour ids bit-exact except on distance
ties, distances within 1e-5 relative for fp32.  The GPU sums the same fp32
terms in a different order than the reference's strict left-to-right loop, so
two candidates whose reference distances differ by less than the tolerance may
legitimately swap ranks (or swap across the k-th boundary); everything else
must match exactly.

For inner product on near-orthogonal data the dot product is ~0 relative to
|q||v|, so "relative" is taken against |q|*max|v| (absolute on that scale);
no reordering of an fp32 sum can do better (hard parts).
"""
import numpy as np

RTOL = 1e-5
FLT_MAX = np.float32(3.4028234663852886e38)
ID_PAD = np.uint64(0xFFFFFFFFFFFFFFFF)


def check_search(D, I, D_ref, I_ref, scale=None, rtol=RTOL):
    """Returns the number of tie-explained id differences; raises on a real mismatch."""
    D = np.asarray(D, np.float32)
    I = np.asarray(I).astype(np.uint64)
    D_ref = np.asarray(D_ref, np.float32)
    I_ref = np.asarray(I_ref, np.uint64)
    assert D.shape == D_ref.shape and I.shape == I_ref.shape, (D.shape, D_ref.shape)
    nq, k = D.shape
    pad_ref = I_ref == ID_PAD
    assert np.array_equal(I == ID_PAD, pad_ref), "padding pattern differs"
    assert np.all(D[pad_ref] == FLT_MAX) and np.all(D_ref[pad_ref] == FLT_MAX)
    if scale is None:
        tol = rtol * np.abs(D_ref.astype(np.float64))
    else:
        tol = rtol * np.broadcast_to(np.asarray(scale, np.float64).reshape(nq, -1), D.shape)
    tol = np.where(pad_ref, 0.0, tol)
    err = np.abs(D.astype(np.float64) - D_ref.astype(np.float64))
    err = np.where(pad_ref, 0.0, err)
    bad = err > tol
    assert not bad.any(), f"distance mismatch: max err {err.max()} at {np.argwhere(bad)[:5]}, D={D[bad][:5]} ref={D_ref[bad][:5]}"
    ties = 0
    for q in np.argwhere((I != I_ref).any(axis=1)).ravel():
        ref_pos = {int(i): j for j, i in enumerate(I_ref[q])}
        for i in np.argwhere(I[q] != I_ref[q]).ravel():
            gid = int(I[q, i])
            t = 2 * tol[q, i]
            if gid in ref_pos:  # same candidate at a different rank: the two ranks must be a tie
                j = ref_pos[gid]
                assert abs(float(D_ref[q, j]) - float(D_ref[q, i])) <= t, \
                    f"q{q} pos{i}: id {gid} ranked {j} by the reference, distances {D_ref[q, j]} vs {D_ref[q, i]}"
            else:  # swapped across the k-th boundary: must tie with the reference's last result
                last = k - 1 - int(pad_ref[q].sum())
                assert abs(float(D[q, i]) - float(D_ref[q, last])) <= 2 * tol[q, last], \
                    f"q{q} pos{i}: id {gid} absent from the reference top-k and not a boundary tie"
            ties += 1
    return ties


def ip_scale(queries, db):
    qn = np.linalg.norm(np.asarray(queries, np.float64), axis=1)
    vn = np.linalg.norm(np.asarray(db, np.float64), axis=1).max()
    return qn * vn


def check_probes(got, ref, queries, centroids, metric, rtol=RTOL):
    """Probe lists vs select_nprobe_lists: equal, or different only where the centroid distances tie within
    the tolerance (rank swaps, or a swap across the nprobe-th boundary).  Returns the number of such ties."""
    got = np.asarray(got).astype(np.int64)
    ref = np.asarray(ref).astype(np.int64)
    assert got.shape == ref.shape
    q = np.asarray(queries, np.float64)
    c = np.asarray(centroids, np.float64)
    ties = 0
    for i in np.argwhere((got != ref).any(axis=1)).ravel():
        if metric == 0:
            d = ((q[i][None, :] - c) ** 2).sum(1)
            tol = 2 * rtol * d
        else:
            d = -(c @ q[i])
            tol = np.full_like(d, 2 * rtol * np.linalg.norm(q[i]) * np.linalg.norm(c, axis=1).max())
        assert len(set(got[i].tolist())) == got.shape[1], "duplicate list id in a probe list"
        for pos in np.argwhere(got[i] != ref[i]).ravel():
            a, b = got[i, pos], ref[i, pos]
            assert abs(d[a] - d[b]) <= max(tol[a], tol[b]), \
                f"query {i} rank {pos}: list {a} (d={d[a]}) vs reference list {b} (d={d[b]})"
            ties += 1
    return ties
