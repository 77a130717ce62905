// primitives.cuh -- device-wide building blocks written for this library (no CUB/Thrust):
// a stable LSD radix sort of (uint32 key, uint32 value) pairs and an exclusive scan.
#pragma once
#include "common.cuh"

namespace vdet {

// ---- exclusive scan of uint32 ----------------------------------------------------------
// out[i] = sum(in[0..i)), *total (device, optional) = sum of all.  in == out allowed.
// Scratch: scan_scratch_elems(n) uint32.
size_t scan_scratch_elems(int64_t n);
int exclusive_scan_u32(const uint32_t* in, uint32_t* out, int64_t n, uint32_t* total,
                       uint32_t* scratch, cudaStream_t st);

// ---- stable radix sort of pairs ---------------------------------------------------------
// Sorts ascending by the key bits [begin_bit, end_bit) (8 bits per pass); equal keys keep
// their input order.  Ping-pongs between (keys,vals) and (keys_alt,vals_alt); returns 0 if
// the result is in (keys,vals), 1 if it is in the alt buffers, <0 on error.
// Scratch: radix_scratch_elems(n) uint32.
size_t radix_scratch_elems(int64_t n);
int radix_sort_pairs(uint32_t* keys, uint32_t* vals, uint32_t* keys_alt, uint32_t* vals_alt,
                     int64_t n, int begin_bit, int end_bit, uint32_t* scratch, cudaStream_t st);

}  // namespace vdet
