// Microbenchmark: how fast can one B200 stream HBM -> SM through (a) LDG.128 and
// (b) a ring of 1-D bulk TMA copies (the list-scan kernel's ingest path) with no
// math attached.  Sets the practical ceiling for the scan's roofline.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o stream_bench stream_bench.cu && ./stream_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t ph) {
    asm volatile("{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(smem_u32(b)), "r"(ph) : "memory");
}
__device__ __forceinline__ void tma(void* d, const void* s, uint32_t n, uint64_t* b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(d)), "l"(s), "r"(n), "r"(smem_u32(b)) : "memory");
}

__global__ void ldg_kernel(const float4* __restrict__ x, size_t n4, float* out) {
    float acc = 0.f;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
    for (; i + 3 * st < n4; i += 4 * st) {
        float4 a = __ldcs(x + i), b = __ldcs(x + i + st), c = __ldcs(x + i + 2 * st), d = __ldcs(x + i + 3 * st);
        acc += a.x + b.y + c.z + d.w;
    }
    if (acc == 123.456f) *out = acc;
}

// one producer thread + 8 consumer warps; consumers touch one float4 per lane per row and release
__global__ void __launch_bounds__(288, 1) tma_kernel(const char* __restrict__ x, size_t bytes, uint32_t stage_bytes, uint32_t S,
                                                     uint32_t chunk_bytes, float* out) {
    extern __shared__ __align__(128) uint8_t sm[];
    uint64_t* full = (uint64_t*)(sm + (size_t)S * stage_bytes);
    uint64_t* empty = full + 8;
    if (threadIdx.x == 0) {
        for (uint32_t i = 0; i < S; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const size_t nchunks = bytes / chunk_bytes;
    const uint32_t per = chunk_bytes / stage_bytes;
    uint32_t st = 0, ph = 0;
    if (threadIdx.x == 256) {
        for (size_t c = blockIdx.x; c < nchunks; c += gridDim.x)
            for (uint32_t k = 0; k < per; ++k) {
                mbar_wait(&empty[st], ph ^ 1);
                mbar_expect_tx(&full[st], stage_bytes);
                tma(sm + (size_t)st * stage_bytes, x + c * chunk_bytes + (size_t)k * stage_bytes, stage_bytes, &full[st]);
                if (++st == S) { st = 0; ph ^= 1; }
            }
    } else if (threadIdx.x < 256) {
        float acc = 0.f;
        for (size_t c = blockIdx.x; c < nchunks; c += gridDim.x)
            for (uint32_t k = 0; k < per; ++k) {
                mbar_wait(&full[st], ph);
                const float4* p = (const float4*)(sm + (size_t)st * stage_bytes);
                for (uint32_t i = threadIdx.x; i < stage_bytes / 16; i += 256) { float4 v = p[i]; acc += v.x + v.w; }
                __syncwarp();
                if ((threadIdx.x & 31) == 0) mbar_arrive(&empty[st]);
                if (++st == S) { st = 0; ph ^= 1; }
            }
        if (acc == 123.456f) *out = acc;
    }
}

int main(int argc, char** argv) {
    const size_t bytes = 24ull << 30;
    char* x; float* out;
    cudaMalloc(&x, bytes); cudaMalloc(&out, 4);
    cudaMemset(x, 1, bytes);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms;
    if (argc > 1) {  // "soak": stream for a few seconds so that nvidia-smi can sample clocks and power beside it
        cudaFuncSetAttribute(tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        const uint32_t stage = 49152, S = 4, chunk = 786432;
        const int reps = 800;
        cudaEventRecord(e0);
        for (int r = 0; r < reps; ++r) tma_kernel<<<148, 288, (size_t)S * stage + 256>>>(x, bytes, stage, S, chunk, out);
        cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        printf("soak: TMA ring 48 KB x 4, %d x 24 GiB in %.1f ms: %.0f GB/s\n", reps, ms, (double)reps * (bytes / chunk * chunk) / ms / 1e6);
        return 0;
    }
    for (int blocks : {148 * 4, 148 * 8, 148 * 16}) {
        ldg_kernel<<<blocks, 512>>>((const float4*)x, bytes / 16, out);
        cudaEventRecord(e0);
        for (int r = 0; r < 3; ++r) ldg_kernel<<<blocks, 512>>>((const float4*)x, bytes / 16, out);
        cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        printf("LDG.128 grid=%d x512: %.0f GB/s\n", blocks, 3.0 * bytes / ms / 1e6);
    }
    cudaFuncSetAttribute(tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    struct Cfg { uint32_t stage, S; } cfgs[] = {{49152, 4}, {49152, 3}, {24576, 8}, {24576, 4}, {16384, 12}, {16384, 6}, {32768, 6}, {65536, 3}, {12288, 16}, {8192, 16}};
    for (auto c : cfgs) {
        uint32_t chunk = 786432 / c.stage * c.stage;
        size_t smem = (size_t)c.S * c.stage + 256;
        tma_kernel<<<148, 288, smem>>>(x, bytes, c.stage, c.S, chunk, out);
        cudaEventRecord(e0);
        for (int r = 0; r < 3; ++r) tma_kernel<<<148, 288, smem>>>(x, bytes, c.stage, c.S, chunk, out);
        cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        size_t moved = bytes / chunk * chunk;
        printf("TMA ring stage=%u KB x %u stages (1 CTA/SM): %.0f GB/s  [%s]\n", c.stage / 1024, c.S, 3.0 * moved / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
