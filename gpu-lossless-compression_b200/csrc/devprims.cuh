// Device-wide building blocks used by the BWT paths: a segmented LSD radix sort of
// (key, value) pairs and single-pass prefix scans.  Hand-written for sm_100a (devprims.cu);
// they replace the toolkit's cub::DeviceRadixSort / cub::DeviceScan on the hot path.
//
// The reference's counterparts are the four CUB sorts + moderngpu merge per skew level in
// cudpp-inpar/src/cudpp/app/sa_app.cu:125-298 and the Thrust sorts of
// cuda-bzip2-ipdpsw/gpuBWTSort.cu:290-418.
#pragma once
#include "common.cuh"

namespace b200lc {
namespace prims {

// Elements per sort call; tile status words carry a 30-bit running count.
constexpr u64 kSortMaxElems = (1ull << 30) - 1;

// Scratch for sort_pairs over n elements in segments of seg_len (seg_len == n: one segment).
size_t sort_scratch_bytes(u64 n, u64 seg_len);

// Stable LSD radix sort on key bits [begin_bit, end_bit), 8 bits per pass, every segment of
// seg_len consecutive elements sorted on its own (the last segment may be shorter).  The data
// ping-pongs between (keys_a, vals_a) and (keys_b, vals_b), starting in a; *result_in_b says where
// the sorted pairs ended up.  n <= kSortMaxElems.  Asynchronous.
template <typename K>
int sort_pairs(K *keys_a, K *keys_b, u32 *vals_a, u32 *vals_b, u64 n, u64 seg_len, int begin_bit,
               int end_bit, void *scratch, size_t scratch_bytes, cudaStream_t stream, int *result_in_b);

// Single-pass scans (decoupled look-back) over n u32 elements; in == out is allowed.
size_t scan_scratch_bytes(u64 n);
int exclusive_sum_u32(const u32 *in, u32 *out, u64 n, void *scratch, size_t scratch_bytes,
                      cudaStream_t stream);
int inclusive_max_u32(const u32 *in, u32 *out, u64 n, void *scratch, size_t scratch_bytes,
                      cudaStream_t stream);

}  // namespace prims
}  // namespace b200lc
