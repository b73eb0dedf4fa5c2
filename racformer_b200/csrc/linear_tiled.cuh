// The pre-tiled operand format of the tcgen05 Linear kernel (csrc/linear.cu).
//
// A [rows][K] fp32 matrix is stored as its three bf16 pieces, cut into [128 rows][32 k] tiles that are laid out in
// global memory exactly as the kernel wants them in shared memory (K-major, 64-byte rows, 64-byte swizzle: the 16-byte
// chunk c of row r sits at chunk c ^ ((r >> 1) & 3)), ordered [row tile][k block][piece]. One pipeline stage of one
// operand -- three 8 KB piece tiles -- is therefore ONE contiguous 24 KB block that a single cp.async.bulk moves in
// full 128-byte lines (the tensor-map path fetches 64-byte rows and tops out at 7.7 TB/s, profiles/r01c_linear_ncu_summary.json).
// Rows beyond `rows` need not be initialised (they only feed output rows that are never stored); k beyond K must be 0.
#pragma once
#include <stdint.h>

namespace racf {

constexpr int kTileRows = 128;
constexpr int kTileK = 32;
constexpr int kTilePieceElems = kTileRows * kTileK;     // 4096 bf16 = 8 KB
constexpr int kTileStageBytes = 3 * kTilePieceElems * 2;

// element offset (in bf16 units) of (row, k) of piece `piece`; num_kblocks = ceil(K / 32)
__host__ __device__ __forceinline__ long long tiled_offset(long long row, int k, int num_kblocks, int piece) {
    const long long row_tile = row >> 7;
    const int rr = (int)(row & 127), kb = k >> 5, kk = k & 31;
    const int chunk = (kk >> 3) ^ ((rr >> 1) & 3);
    return ((row_tile * num_kblocks + kb) * 3 + piece) * (long long)kTilePieceElems + rr * kTileK + chunk * 8 + (kk & 7);
}

}  // namespace racf
