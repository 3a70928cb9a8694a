// codec.cu — methylation count files parsed on the GPU (what MethyLasso.R:246-262 does
// with data.table::fread).  The text of one file is uploaded as it is; kernels find the line starts (newline count per
// tile, scan, scatter), parse the four selected fields of every line (codec_core.cuh: exact decimal -> double or a
// flag for the host, never an approximation) and mark where the chromosome field changes, so that the host only has
// to look at one name per run.  HBM-bound byte work: the text is read three times, ~40 bytes per line are written.
#include <cub/device/device_select.cuh>
#include <cub/iterator/counting_input_iterator.cuh>

#include <algorithm>
#include <vector>

#include "codec.cuh"
#include "codec_core.cuh"
#include "hd.h"

namespace ml {
namespace {

constexpr int TXT_TILE = 4096;            // bytes per block
constexpr int TXT_PER_THREAD = TXT_TILE / 256;

__global__ void __launch_bounds__(256)
k_txt_count(const char* __restrict__ text, int64_t nbytes, int32_t* __restrict__ tile_count) {
  __shared__ int32_t sm[8];
  const int64_t base = (int64_t)blockIdx.x * TXT_TILE + (int64_t)threadIdx.x * TXT_PER_THREAD;
  int c = 0;
#pragma unroll
  for (int k = 0; k < TXT_PER_THREAD; ++k) c += (base + k < nbytes && text[base + k] == '\n');
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int k = 0; k < 8; ++k) t += sm[k];
    tile_count[blockIdx.x] = t;
  }
}

// exclusive scan of the per-tile newline counts (tables of a few 10^5 entries: one block walks them in chunks)
__global__ void __launch_bounds__(1024)
k_txt_scan_tiles(const int32_t* __restrict__ count, int64_t n, int64_t* __restrict__ base, int64_t* __restrict__ total) {
  __shared__ int64_t warp_sum[32];
  __shared__ int64_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int64_t start = 0; start < n; start += 1024) {
    const int64_t i = start + threadIdx.x;
    const int64_t v = i < n ? count[i] : 0;
    int64_t x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const int64_t y = __shfl_up_sync(0xffffffffu, x, d); if (lane >= d) x += y; }
    if (lane == 31) warp_sum[wid] = x;
    __syncthreads();
    if (wid == 0) {
      int64_t w = warp_sum[lane];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) { const int64_t y = __shfl_up_sync(0xffffffffu, w, d); if (lane >= d) w += y; }
      warp_sum[lane] = w;
    }
    __syncthreads();
    const int64_t before = carry + (wid ? warp_sum[wid - 1] : 0) + (x - v);
    if (i < n) base[i] = before;
    __syncthreads();
    if (threadIdx.x == 1023) carry += warp_sum[31];
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = carry;
}

// line_start[k + 1] = byte after the k-th newline (line_start[0] = 0 is written by the host side)
__global__ void __launch_bounds__(256)
k_txt_starts(const char* __restrict__ text, int64_t nbytes, const int64_t* __restrict__ tile_base,
             int64_t* __restrict__ line_start) {
  __shared__ int32_t sm[8];
  const int64_t base = (int64_t)blockIdx.x * TXT_TILE + (int64_t)threadIdx.x * TXT_PER_THREAD;
  unsigned mask = 0;
#pragma unroll
  for (int k = 0; k < TXT_PER_THREAD; ++k) mask |= (unsigned)(base + k < nbytes && text[base + k] == '\n') << k;
  const int c = __popc(mask);
  int x = c;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, d); if (lane >= d) x += y; }
  if (lane == 31) sm[wid] = x;
  __syncthreads();
  int before = x - c;
  for (int k = 0; k < wid; ++k) before += sm[k];
  int64_t o = tile_base[blockIdx.x] + before + 1;
  while (mask) {
    const int k = __ffs(mask) - 1;
    mask &= mask - 1;
    line_start[o++] = base + k + 1;
  }
}

__global__ void __launch_bounds__(256)
k_txt_parse(const char* __restrict__ text, int64_t nbytes, const int64_t* __restrict__ line_start, int64_t n_lines_total,
            int64_t skip, int col_a, int col_b, char sep, int32_t* __restrict__ pos, double* __restrict__ a,
            double* __restrict__ b, int8_t* __restrict__ status, int32_t* __restrict__ chr_off, int32_t* __restrict__ chr_len) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t i = r + skip;
  if (i >= n_lines_total) return;
  const int64_t s = line_start[i];
  int64_t e = i + 1 < n_lines_total ? line_start[i + 1] - 1 : nbytes;      // without the newline
  if (e > s && e == nbytes && text[e - 1] == '\n') --e;                    // last line terminated by a newline
  if (e < s) e = s;
  int64_t ee = e;
  if (ee > s && text[ee - 1] == '\r') --ee;
  if (ee == s) {                                                           // blank line
    pos[r] = 0; a[r] = 0; b[r] = 0; chr_off[r] = 0; chr_len[r] = 0; status[r] = TXT_EMPTY;
    return;
  }
  const TxtLine L = txt_parse_line(text + s, text + e, col_a, col_b, sep);
  pos[r] = (int32_t)L.pos;
  a[r] = L.a;
  b[r] = L.b;
  chr_off[r] = L.chr_off;
  chr_len[r] = L.chr_len;
  status[r] = (int8_t)L.status;
}

// 1 where the chromosome field of data line r differs from that of line r - 1 (or r == 0)
__global__ void __launch_bounds__(256)
k_txt_chr_heads(const char* __restrict__ text, const int64_t* __restrict__ line_start, int64_t skip, int64_t n,
                const int32_t* __restrict__ chr_off, const int32_t* __restrict__ chr_len, int32_t* __restrict__ flag) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  bool head = r == 0 || chr_len[r] != chr_len[r - 1];
  if (!head) {
    const char* p = text + line_start[r + skip] + chr_off[r];
    const char* q = text + line_start[r - 1 + skip] + chr_off[r - 1];
    for (int k = 0; k < chr_len[r] && !head; ++k) head = p[k] != q[k];
  }
  flag[r] = head ? 1 : 0;
}

}  // namespace

void text_parse(Engine& eng, const char* text, int64_t nbytes, int64_t skip_lines, int col_a, int col_b, char sep,
                TextTable& out) {
  cudaStream_t st = eng.stream;
  out = TextTable();
  if (nbytes <= 0) return;
  if (col_a < 2 || col_b < 2) throw MlError(19, "value columns must come after the chromosome and position columns");
  out.text.alloc(st, (size_t)nbytes);
  ML_CUDA(cudaMemcpyAsync(out.text.p, text, (size_t)nbytes, cudaMemcpyHostToDevice, st));
  const int64_t n_tiles = (nbytes + TXT_TILE - 1) / TXT_TILE;
  DevBuf<int32_t> tile_count(st, (size_t)n_tiles);
  DevBuf<int64_t> tile_base(st, (size_t)n_tiles), total(st, 1);
  ML_LAUNCH(eng, "k_txt_count", k_txt_count<<<(unsigned)n_tiles, 256, 0, st>>>(out.text.p, nbytes, tile_count.p));
  ML_LAUNCH(eng, "k_txt_scan_tiles", k_txt_scan_tiles<<<1, 1024, 0, st>>>(tile_count.p, n_tiles, tile_base.p, total.p));
  int64_t newlines = 0;
  total.download(&newlines, 1);
  stream_sync(st);
  const bool open_end = text[nbytes - 1] != '\n';
  const int64_t n_total = newlines + (open_end ? 1 : 0);           // lines of the file
  out.line_start.alloc(st, (size_t)newlines + 2);
  ML_CUDA(cudaMemsetAsync(out.line_start.p, 0, sizeof(int64_t), st));      // line 0 starts at byte 0
  ML_LAUNCH(eng, "k_txt_starts", k_txt_starts<<<(unsigned)n_tiles, 256, 0, st>>>(out.text.p, nbytes, tile_base.p, out.line_start.p));
  out.skip = std::min<int64_t>(std::max<int64_t>(skip_lines, 0), n_total);
  const int64_t n = n_total - out.skip;
  out.n = n;
  if (n == 0) { stream_sync(st); return; }
  const size_t c = (size_t)n;
  out.pos.alloc(st, c); out.a.alloc(st, c); out.b.alloc(st, c); out.status.alloc(st, c);
  out.chr_off.alloc(st, c); out.chr_len.alloc(st, c);
  ML_LAUNCH(eng, "k_txt_parse", k_txt_parse<<<cdiv(n, 256), 256, 0, st>>>(
      out.text.p, nbytes, out.line_start.p, n_total, out.skip, col_a, col_b, sep, out.pos.p, out.a.p, out.b.p,
      out.status.p, out.chr_off.p, out.chr_len.p));
  DevBuf<int32_t> flag(st, c);
  ML_LAUNCH(eng, "k_txt_chr_heads", k_txt_chr_heads<<<cdiv(n, 256), 256, 0, st>>>(
      out.text.p, out.line_start.p, out.skip, n, out.chr_off.p, out.chr_len.p, flag.p));
  // run_first = the data lines whose flag is set, in order (cub stream compaction: library plumbing)
  out.run_first.alloc(st, c);
  size_t tmp_bytes = 0;
  cub::CountingInputIterator<int64_t> line_index(0);
  if (n >= 0x7fffffff) throw MlError(19, "too many lines");
  ML_CUDA(cub::DeviceSelect::Flagged(nullptr, tmp_bytes, line_index, flag.p, out.run_first.p, total.p, (int)n, st));
  DevBuf<char> tmp(st, tmp_bytes + 16);
  ML_CUDA(cub::DeviceSelect::Flagged(tmp.p, tmp_bytes, line_index, flag.p, out.run_first.p, total.p, (int)n, st));
  eng.launches++;
  int64_t runs = 0;
  total.download(&runs, 1);
  stream_sync(st);
  out.n_runs = runs;
}

}  // namespace ml
