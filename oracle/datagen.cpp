// TEST INFRASTRUCTURE ONLY.  Synthetic inputs exactly as the reference's own
// drivers make them: std::mt19937(seed) feeding std::normal_distribution<float>
// (0,1), database first, then queries from the same stream
// (test/gpu_vs_cpu_test.cpp:83-94, test/simple_test.cpp:119-138).  The
// sequence is libstdc++-specific, so it is generated with the same C++
// generators rather than numpy's.
#include <cstdint>
#include <random>
#include <vector>

extern "C" void gen_gaussian(uint32_t seed, float* out, uint64_t n) {
    std::mt19937 gen(seed);
    std::normal_distribution<float> dist(0.0f, 1.0f);
    for (uint64_t i = 0; i < n; ++i) out[i] = dist(gen);
}

// Clustered data (not in the reference's drivers): n_centers Gaussian blobs of
// the given spread.  Exercises small distances relative to the vector norms,
// where a |q|^2+|v|^2-2qv formulation would lose digits.
extern "C" void gen_clustered(uint32_t seed, float* out, uint64_t n, uint32_t dim,
                              uint32_t n_centers, float spread) {
    std::mt19937 gen(seed);
    std::normal_distribution<float> dist(0.0f, 1.0f);
    std::vector<float> centers((size_t)n_centers * dim);
    for (auto& c : centers) c = 4.0f * dist(gen);
    std::uniform_int_distribution<uint32_t> pick(0, n_centers - 1);
    for (uint64_t i = 0; i < n; ++i) {
        const float* c = centers.data() + (size_t)pick(gen) * dim;
        for (uint32_t d = 0; d < dim; ++d) out[i * dim + d] = c[d] + spread * dist(gen);
    }
}
