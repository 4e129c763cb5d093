"""Generates tests/golden/*.npz from the REFERENCE's own CPU path
(oracle/_ref/libvdbref.so = /root/reference/engine/ivf_flat_index.cpp compiled
unmodified, use_gpu=false).  Run in the build container, where /root/reference
exists; the fixtures are committed because the reference cannot travel to the
GPU box.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as O  # noqa: E402

# name: (seed, n, dim, nlist, ntrain, nq, nprobe, k, metric)
CASES = {
    # test/simple_test.cpp:111-138,159-165
    "simple_test": (42, 1000, 64, 16, 100, 5, 4, 5, O.METRIC_L2),
    # test/CMakeLists.txt:56 runs gpu_vs_cpu_test as "10000 100 64 32"; nprobe 8, k 10, seed 12345
    "ctest_gpu_vs_cpu": (12345, 10000, 64, 32, 10000, 100, 8, 10, O.METRIC_L2),
    "small_ip": (12345, 4000, 48, 24, 2000, 32, 6, 10, O.METRIC_IP),
    # BASELINE.json configs[0]: 100K x 128, nlist 128, nprobe 16, k 10, batch 64, train on 10K
    "config1": (12345, 100000, 128, 128, 10000, 64, 16, 10, O.METRIC_L2),
}


def make(name):
    seed, n, dim, nlist, ntrain, nq, nprobe, k, metric = CASES[name]
    x = O.gaussian(seed, n + nq, dim)
    db, q = x[:n], x[n:]
    ref = O.RefIndex(dim, nlist, metric)
    ref.train(db[:ntrain])
    cent = ref.centroids
    ref.add(db)
    D, I = ref.search(q, nprobe, k)
    probes = np.stack([ref.select_nprobe(q[i], nprobe) for i in range(nq)])
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"),
        params=np.array([seed, n, dim, nlist, ntrain, nq, nprobe, k, metric], np.int64),
        centroids=cent, list_sizes=ref.list_sizes(), probes=probes, D=D, I=I,
        # a few input rows so the data generator itself is pinned
        db_head=db[:4].copy(), q_head=q[:2].copy())
    print(name, "ok", D[0, :3], I[0, :3])


if __name__ == "__main__":
    for nm in (sys.argv[1:] or CASES):
        make(nm)
