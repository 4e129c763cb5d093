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
};

inline uint32_t next_pow2(uint32_t v) {
    uint32_t p = 1;
    while (p < v) p <<= 1;
    return p;
}

inline uint32_t round_up(uint32_t v, uint32_t m) { return (v + m - 1) / m * m; }

}  // namespace vdb
