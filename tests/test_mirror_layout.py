"""The shadow of a page is stored as ready-made tensor-core operand tiles (csrc/common.cuh, mirror_elem_off): a plain
1-D bulk copy of 16 KB must leave in shared memory exactly what a SWIZZLE_128B tensor map would -- rows of 128 bytes,
16-byte chunks XOR-ed with the row's low three bits.  Compiled for the host with nvcc and checked here, on the CPU:
every element of a page gets its own bytes, tiles are contiguous 16 KB blocks in (row tile, K block) order, and inside
a tile the address is the canonical swizzle of the linear one."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = r"""
#include <cstdio>
#include <vector>
#include "common.cuh"
using namespace vdb;
int main() {
    for (uint32_t eb = 1; eb <= 2; ++eb) {            // int8, bf16
        for (uint32_t ld : {128u, 256u, 768u, 1024u}) {
            const uint32_t page_rows = 256, row_bytes = ld * eb, nkb = row_bytes / 128;
            std::vector<int> seen(page_rows * row_bytes, 0);
            for (uint32_t r = 0; r < page_rows; ++r)
                for (uint32_t e = 0; e < ld; ++e) {
                    const uint32_t off = mirror_elem_off(r, e, ld, eb);
                    if (off + eb > seen.size()) { std::printf("out of range\n"); return 1; }
                    for (uint32_t b = 0; b < eb; ++b) if (seen[off + b]++) { std::printf("overlap\n"); return 1; }
                    // tile (row tile, K block), 16 KB each, in that order
                    const uint32_t byte = e * eb, tile = (r / 128) * nkb + byte / 128;
                    if (off / MIRROR_TILE_BYTES != tile) { std::printf("tile order\n"); return 1; }
                    // inside the tile: linear address = row * 128 + byte in row; swizzle = bits [4,7) ^= bits [7,10)
                    const uint32_t lin = (r % 128) * 128 + byte % 128;
                    const uint32_t swz = lin ^ (((lin >> 7) & 7u) << 4);
                    if (off % MIRROR_TILE_BYTES != swz) { std::printf("swizzle\n"); return 1; }
                }
            for (int c : seen) if (c != 1) { std::printf("hole\n"); return 1; }
        }
    }
    // the query image uses the same function with 64-row (bf16) / 128-row (int8) tiles
    if (mirror_elem_off(63, 0, 768, 2, 64) != 63 * 128 + ((0 ^ 7) << 4)) { std::printf("query image\n"); return 1; }
    if (mirror_elem_off(64 + 5, 16, 768, 1, 128) != 69 * 128 + ((1 ^ 5) << 4)) { std::printf("query image i8\n"); return 1; }
    std::printf("ok\n");
    return 0;
}
"""


@pytest.mark.skipif(shutil.which("nvcc") is None, reason="needs nvcc (host-only compile)")
def test_shadow_layout_is_the_128_byte_swizzle_image(tmp_path):
    src = tmp_path / "layout.cu"
    src.write_text(SRC)
    exe = tmp_path / "layout"
    csrc = os.path.join(ROOT, "cuda-acceleratedvectordatabaseengine_b200", "csrc")
    subprocess.run(["nvcc", "-std=c++17", "-O1", "-I", csrc, "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)],
                   check=True, capture_output=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.strip() == "ok", out.stdout + out.stderr
