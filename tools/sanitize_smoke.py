"""Smallest end-to-end pass over the hot path for compute-sanitizer (memcheck): train, add, search (synchronous and
pipelined), an odd dimension (masked scan path), exact brute force, kmeans_assign.  Results are checked against the
oracle so that a clean sanitizer run is a run of the real code paths."""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O  # noqa: E402
from parity import check_search  # noqa: E402

pkg = importlib.import_module("cuda-acceleratedvectordatabaseengine_b200")
for dim, nlist in ((64, 16), (50, 8)):
    n, nq, nprobe, k = 4000, 8, 4, 5
    x = O.gaussian(42 + dim, n + nq, dim)
    db, q = x[:n], x[n:]
    ora = O.OracleIndex(dim, nlist)
    ora.train(db[:800])
    ora.add(db)
    Dr, Ir = ora.search(q, nprobe, k)
    ix = pkg.IVFFlatIndex(pkg.Config(dimension=dim, nlist=nlist))
    ix.train(db[:800])
    assert np.array_equal(ix.centroids, ora.centroids)
    ix.add(db)
    D, I = ix.search(q, nprobe, k)
    check_search(D, I, Dr, Ir)
    D2, I2 = np.empty_like(D), np.empty_like(I)
    ts = [ix.search_submit(np.ascontiguousarray(q[lo:lo + 4]), nprobe, k, D2[lo:lo + 4], I2[lo:lo + 4]) for lo in (0, 4)]
    for t in ts:
        ix.search_wait(t)
    assert np.array_equal(D2, D) and np.array_equal(I2, I)
    Db, Ib = pkg.bruteforce_search(db, q, k)
    Df, If = O.flat_search(db, q, k)
    check_search(Db, Ib, Df, If)
    a = pkg.kmeans_assign(db[:500], ora.centroids)
    assert np.array_equal(a, ora.assign(db[:500]))
    ix.close()
print("sanitize smoke ok")
