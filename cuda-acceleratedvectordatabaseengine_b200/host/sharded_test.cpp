// One vdb::IVFFlatIndex object over several GPUs, driven through the C++ mirror exactly like the single-GPU one
// (Config::devices is the only difference): BASELINE.json configs[0] -- 100 000 x 128-D, nlist 128, train on the
// first 10 000, nprobe 16, k 10, 64 queries, std::mt19937(12345) + normal(0,1) as test/gpu_vs_cpu_test.cpp:83-94
// generates its data -- i.e. the inputs of the golden fixture tests/golden/config1.npz.  Prints the centroids'
// checksum and every result so that the python test can compare them with the fixture; also saves the index as an
// epoch directory and loads it into a second multi-GPU index (IVFFlatIndex::save / load_from_epoch).
//   usage: sharded_test <dev0,dev1,...> <tmp dir>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <sstream>

#include "ivf_flat_index.h"

struct Epoch {  // what server/query_service.cpp:242-245 hands to load_from_epoch (format/storage.h:175-181)
    std::string id, base_path;
};

int main(int argc, char** argv) {
    using namespace vdb;
    if (argc < 3) {
        std::printf("usage: sharded_test <devices> <tmp dir>\n");
        return 2;
    }
    std::vector<int> devices;
    {
        std::stringstream ss(argv[1]);
        std::string tok;
        while (std::getline(ss, tok, ',')) devices.push_back(std::atoi(tok.c_str()));
    }
    const uint32_t n = 100000, dim = 128, nlist = 128, ntrain = 10000, nq = 64, nprobe = 16, k = 10;
    std::mt19937 gen(12345);
    std::normal_distribution<float> dist(0.0f, 1.0f);
    std::vector<float> x((size_t)(n + nq) * dim);
    for (auto& v : x) v = dist(gen);
    const float* db = x.data();
    const float* q = x.data() + (size_t)n * dim;
    std::vector<uint64_t> ids(n);
    for (uint32_t i = 0; i < n; ++i) ids[i] = i;
    try {
        TransferManager::Config tc;
        tc.pinned_pool_size = 16 << 20;
        tc.device_pool_size = 32 << 20;
        tc.device = devices[0];
        TransferManager tm(tc);
        IVFFlatIndex::Config cfg{};
        cfg.dimension = dim;
        cfg.nlist = nlist;
        cfg.metric = kernels::Metric::L2;
        cfg.devices = devices;
        IVFFlatIndex index(cfg, &tm);
        index.train(db, ntrain);
        index.add(db, ids.data(), n / 2);
        index.add(db + (size_t)(n / 2) * dim, ids.data() + n / 2, n - n / 2);
        IVFFlatIndex::SearchParams sp;
        sp.nprobe = nprobe;
        sp.k = k;
        std::vector<float> D((size_t)nq * k), D2((size_t)nq * k), D3((size_t)nq * k);
        std::vector<uint64_t> I((size_t)nq * k), I2((size_t)nq * k), I3((size_t)nq * k);
        index.search(q, nq, sp, D.data(), I.data());
        bool ok = index.get_total_vectors() == n;
        // the pipelined form: four 16-query batches in flight
        uint64_t t[4];
        for (int b = 0; b < 4; ++b)
            t[b] = index.search_submit(q + (size_t)b * 16 * dim, 16, sp, D2.data() + (size_t)b * 16 * k, I2.data() + (size_t)b * 16 * k);
        for (int b = 3; b >= 0; --b) index.search_wait(t[b]);
        ok = ok && std::memcmp(D.data(), D2.data(), D.size() * 4) == 0 && std::memcmp(I.data(), I2.data(), I.size() * 8) == 0;
        // save -> load_from_epoch into a fresh multi-GPU index
        Epoch ep{"epoch_000001", std::string(argv[2]) + "/epoch_000001"};
        index.save(ep.base_path);
        IVFFlatIndex loaded(cfg, &tm);
        loaded.load_from_epoch(&ep);
        loaded.search(q, nq, sp, D3.data(), I3.data());
        ok = ok && loaded.get_total_vectors() == n && std::memcmp(D.data(), D3.data(), D.size() * 4) == 0 &&
             std::memcmp(I.data(), I3.data(), I.size() * 8) == 0;
        for (uint32_t i = 0; i < nq * k; ++i) std::printf("R %u %llu %.9g\n", i / k, (unsigned long long)I[i], D[i]);
        std::printf("shards %zu gpu_bytes %zu\n", devices.size(), index.get_gpu_memory_usage());
        std::printf(ok ? "PASSED\n" : "FAILED\n");
        return ok ? 0 : 1;
    } catch (const std::exception& e) {
        std::printf("EXCEPTION %s\n", e.what());
        return 2;
    }
}
