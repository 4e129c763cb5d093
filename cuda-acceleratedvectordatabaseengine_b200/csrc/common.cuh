// Shared declarations of the B200 IVF-Flat library (host + device).
#pragma once
#include <cuda_runtime.h>
#include <cfloat>
#include <cstdint>
#include <cstdio>
#include <string>

#include "vdb_b200.h"

namespace vdb {

constexpr uint64_t ID_PAD = 0xFFFFFFFFFFFFFFFFull;  // UINT64_MAX padding, ivf_flat_index.cpp:382
constexpr int NUM_SMS_B200 = 148;

void set_last_error(const std::string& msg);

struct Status {
    int32_t code = VDB_OK;
    bool ok() const { return code == VDB_OK; }
};

#define VDB_CUDA_TRY(expr)                                                                      \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess) {                                                                \
            ::vdb::set_last_error(std::string(#expr) + ": " + cudaGetErrorString(_e) + " at " + \
                                  __FILE__ + ":" + std::to_string(__LINE__));                   \
            return (_e == cudaErrorMemoryAllocation) ? VDB_OUT_OF_MEMORY : VDB_CUDA_ERROR;      \
        }                                                                                       \
    } while (0)

#define VDB_TRY(expr)                    \
    do {                                 \
        int32_t _s = (expr);             \
        if (_s != VDB_OK) return _s;     \
    } while (0)

#define VDB_REQUIRE(cond, msg)                     \
    do {                                           \
        if (!(cond)) {                             \
            ::vdb::set_last_error(msg);            \
            return VDB_INVALID_ARGUMENT;           \
        }                                          \
    } while (0)

// A paged view of row-major fp32 rows grouped into lists.  Inverted lists are
// chains of fixed-size HBM pages ([page_rows][ld] fp32 followed by
// [page_rows] u64 ids); a flat array (centroids, a brute-force database) is
// viewed as one list whose pages are consecutive row blocks with implicit ids.
struct ListTable {
    const uint32_t* rows;      // [nlist] rows in each list
    const uint32_t* page_off;  // [nlist+1] first page of each list in page_vec/page_ids
    const uint64_t* page_vec;  // [npages] device address of the page's row block
    const uint64_t* page_ids;  // [npages] device address of the page's ids, 0 = implicit
    const uint64_t* ids_flat;  // optional ids for a flat view (row number indexes it), may be null
    uint32_t nlist;
    uint32_t page_rows;
    uint32_t ld;  // floats per row, multiple of 4
    // Low-precision shadow of every page for the tensor-core screen of the list scan (0 = none): byte offset, from
    // the page's row block, of the page's rows rounded to bf16 (mirror_kind 1) or quantised to int8 with one scale
    // per row (mirror_kind 2), stored as the shared-memory image of [128 rows][128 bytes] tensor-core operand tiles
    // (128-byte swizzle), row tile major, then K block.  Behind the page's norms sit [page_rows] fp32
    // |v - shadow(v)| and (int8) [page_rows] fp32 row scales.
    uint32_t mirror_off = 0;
    uint32_t mirror_kind = 0;
};

constexpr uint32_t MIRROR_NONE = 0, MIRROR_BF16 = 1, MIRROR_I8 = 2;
constexpr uint32_t MIRROR_TILE_ROWS = 128;   // rows per operand tile of the shadow (UMMA M)
constexpr uint32_t MIRROR_TILE_BYTES = MIRROR_TILE_ROWS * 128;  // a tile row is 128 bytes = one swizzle atom
__host__ __device__ inline uint32_t mirror_elem_bytes(uint32_t kind) { return kind == MIRROR_I8 ? 1u : 2u; }

// byte offset, inside a page's (or a query block's) shadow, of element e of row r (elements of eb = 1 or 2 bytes):
// tile (r / tile_rows, e / (128 / eb)), row r % tile_rows at 128 bytes per row, 16-byte chunks XOR-swizzled by the
// row's low three bits -- exactly what a SWIZZLE_128B tensor map would leave in shared memory, so a plain bulk copy
// of a tile yields a UMMA operand
__host__ __device__ inline uint32_t mirror_elem_off(uint32_t r, uint32_t e, uint32_t ld, uint32_t eb,
                                                    uint32_t tile_rows = MIRROR_TILE_ROWS) {
    const uint32_t b = e * eb;  // byte position inside the row
    const uint32_t rt = r / tile_rows, rr = r % tile_rows, kb = b >> 7, c = (b >> 4) & 7u, w = b & 15u;
    return (rt * ((ld * eb) >> 7) + kb) * (tile_rows * 128u) + rr * 128u + ((c ^ (rr & 7u)) << 4) + w;
}

inline uint32_t next_pow2(uint32_t v) {
    uint32_t p = 1;
    while (p < v) p <<= 1;
    return p;
}

inline uint32_t round_up(uint32_t v, uint32_t m) { return (v + m - 1) / m * m; }

}  // namespace vdb
