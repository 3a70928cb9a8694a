// hostpack.h — host-side passes over the input columns of a batch (compiled by the host compiler alone, hostpack.cpp):
// the scan of the dataset column for its runs and the packing of Nmeth / coverage for the bus, on a persistent pool of
// worker threads.  No CUDA here.
#pragma once
#include <cstdint>
#include <functional>
#include <vector>

namespace ml {

// run fn(item) for item = 0 .. n_items-1 on up to `threads` threads (the caller is one of them); items are handed out one
// at a time.  The workers live as long as the process; one parallel_for at a time (callers are serialised).
void parallel_for(int n_items, int threads, const std::function<void(int)>& fn);

// rows [lo, hi): out[2 i] = Nmeth[i], out[2 i + 1] = coverage[i] as bytes / 16-bit words (i = row index of the OUTPUT
// layout; the input row is i - shift).  Returns false when a value does not fit (or is negative); the rest of the piece
// is still written.
bool pack_counts_u8(const int32_t* nmeth, const int32_t* cov, int64_t shift, int64_t lo, int64_t hi, uint8_t* out);
bool pack_counts_u16(const int32_t* nmeth, const int32_t* cov, int64_t shift, int64_t lo, int64_t hi, uint16_t* out);

// indices i in [lo, hi), i >= 1, with v[i] != v[i - 1], appended to out in increasing order
void transitions_in(const int32_t* v, int64_t lo, int64_t hi, std::vector<int64_t>& out);

}  // namespace ml
