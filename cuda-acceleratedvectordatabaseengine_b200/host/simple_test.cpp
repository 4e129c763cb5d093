// The reference's smoke test shape (test/simple_test.cpp:99-204) written
// against the C++ mirror: 1000 x 64D, nlist 16, train on the first 100,
// nprobe 4, k 5, mt19937(42) normal(0,1).  Passes iff every id is < n or the
// UINT64_MAX pad, and distances come back sorted and finite.  Prints the first
// query's results so the python test can compare them with the golden fixture.
#include <cmath>
#include <cstdio>
#include <iostream>
#include <random>

#include "ivf_flat_index.h"

int main() {
    using namespace vdb;
    const uint32_t n = 1000, dim = 64, nlist = 16, nq = 5, nprobe = 4, k = 5;
    std::mt19937 gen(42);
    std::normal_distribution<float> dist(0.0f, 1.0f);
    std::vector<float> db((size_t)n * dim), q((size_t)nq * dim);
    for (auto& v : db) v = dist(gen);
    for (auto& v : q) v = dist(gen);
    std::vector<uint64_t> ids(n);
    for (uint32_t i = 0; i < n; ++i) ids[i] = i;
    try {
        TransferManager::Config tc;
        tc.pinned_pool_size = 16 << 20;
        tc.device_pool_size = 64 << 20;
        TransferManager tm(tc);
        IVFFlatIndex::Config cfg{};
        cfg.dimension = dim;
        cfg.nlist = nlist;
        cfg.metric = kernels::Metric::L2;
        IVFFlatIndex index(cfg, &tm);
        index.train(db.data(), 100);
        index.add(db.data(), ids.data(), n);
        IVFFlatIndex::SearchParams sp;
        sp.nprobe = nprobe;
        sp.k = k;
        std::vector<float> D((size_t)nq * k);
        std::vector<uint64_t> I((size_t)nq * k);
        index.search(q.data(), nq, sp, D.data(), I.data());
        bool ok = index.get_total_vectors() == n && index.get_dimension() == dim;
        for (uint32_t i = 0; i < nq * k; ++i) {
            ok = ok && (I[i] < n || I[i] == UINT64_MAX) && std::isfinite(D[i]) && D[i] >= 0.f;
            if (i % k) ok = ok && D[i] >= D[i - 1];
        }
        for (uint32_t i = 0; i < nq * k; ++i) std::printf("R %u %llu %.9g\n", i / k, (unsigned long long)I[i], D[i]);
        bool threw = false;
        try {
            IVFFlatIndex::Config bad{};
            bad.dimension = 0;
            bad.nlist = 4;
            bad.metric = kernels::Metric::L2;
            IVFFlatIndex b(bad, &tm);
        } catch (const std::invalid_argument&) {
            threw = true;
        }
        ok = ok && threw;
        // the kernel seam (engine/kernels.cu launchers): nprobe = nlist makes IVF exact, so it must agree with the
        // brute-force launcher; every row's nearest centroid by the assign launcher must be a valid list
        std::vector<float> Db((size_t)nq * k), De((size_t)nq * k);
        std::vector<uint64_t> Ib((size_t)nq * k), Ie((size_t)nq * k);
        kernels::launch_bruteforce_search<float>(db.data(), q.data(), ids.data(), n, nq, dim, k, Db.data(), Ib.data(),
                                                 kernels::Metric::L2, nullptr);
        sp.nprobe = nlist;
        index.search(q.data(), nq, sp, De.data(), Ie.data());
        for (uint32_t i = 0; i < nq * k; ++i) ok = ok && Ib[i] == Ie[i] && std::fabs(Db[i] - De[i]) <= 1e-5f * De[i];
        std::vector<uint32_t> assign(n);
        std::vector<float> cent((size_t)nlist * dim);
        for (uint32_t c = 0; c < nlist; ++c)
            for (uint32_t d = 0; d < dim; ++d) cent[(size_t)c * dim + d] = db[(size_t)(c * 7) * dim + d];
        kernels::launch_kmeans_assign<float>(db.data(), cent.data(), assign.data(), nullptr, n, nlist, dim, nullptr);
        for (uint32_t c = 0; c < nlist; ++c) ok = ok && assign[c * 7] == c;  // a row that IS a centroid picks it
        std::printf(ok ? "PASSED\n" : "FAILED\n");
        return ok ? 0 : 1;
    } catch (const std::exception& e) {
        std::printf("EXCEPTION %s\n", e.what());
        return 2;
    }
}
