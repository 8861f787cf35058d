// clb_prep.cuh -- device-side row preparation (sm_100a): the reference's input tuple (careless/models/base.py:22-31,
// built by io/formatter.py:382-400, 631-653) -> the sorted / padded SoA rows of the observation kernels.
//
// The host version of the same transformation (plan_rows / fill_rows in clb_api.cu, a single-threaded counting sort) costs
// 0.35 s at 10 M rows and 8.8 s at 200 M; here the O(n) work runs on the GPU, bit-identical to it:
//   k_prep_keys        range checks (first offending row by atomicMin), the 32-bit sort key (refl_id | harmonic_id | image_id)
//                      and, where the row positions depend on run lengths (Laue spots, image tiles), a histogram of the keys;
//   k_rs_hist / k_rs_scatter
//                      STABLE least-significant-digit radix sort of (key, row) pairs, 8 bits per pass, ceil(log2(n_keys) / 8)
//                      passes: a block owns 4 096 consecutive rows, a warp 512 of them; the rank of a row among the rows of equal
//                      digit is (digit total of earlier blocks, from a digit-major exclusive scan) + (earlier warps of the block)
//                      + (earlier 32-row chunks of the warp) + (lower lanes with the same digit, __match_any_sync) -- the order
//                      of equal keys is the input order, exactly what the host's counting sort produces;
//   k_scan_tile / k_scan_add
//                      exclusive prefix sums (block histograms of a pass; key offsets), recursive over 2 048-element tiles;
//   k_prep_pos         padded position of every sorted row from its key's start position (computed on the host from the
//                      n_keys run lengths: the padding rules of Laue spots / image tiles are a sequential recurrence over KEYS,
//                      not rows) -- only for the orders that pad;
//   k_fill_defaults / k_fill_gather
//                      the SoA rows: padding values, then one gather per row through the sort permutation.
// All of it is HBM-bound integer / byte work: coalesced streaming reads and writes except the gather through the permutation
// (one 32-byte sector per array and row) and the radix scatter (256 write streams per block).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace clb {
namespace prep {

constexpr int kRsThreads = 256, kRsItems = 16, kRsTile = kRsThreads * kRsItems;     // rows per block of a radix pass
constexpr int kRsWarps = kRsThreads / 32, kRsWarpItems = kRsTile / kRsWarps;
constexpr int kScanThreads = 256, kScanItems = 8, kScanTile = kScanThreads * kScanItems;
constexpr unsigned long long kNoBadRow = ~0ull;

// order: CLB_ORDER_REFL 1, CLB_ORDER_SPOT 2, CLB_ORDER_IMAGE 3 (include/careless_b200.h)
__global__ void __launch_bounds__(256) k_prep_keys(int64_t n, const int64_t* refl_id, const int64_t* image_id, const int64_t* harmonic_id,
                                                   const int64_t* obs_index, int64_t R, int64_t n_images, int64_t n_total, int laue, int order,
                                                   uint32_t* key, unsigned long long* first_bad, uint32_t* count) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = refl_id[i];
    bool bad = r < 0 || r >= R;
    int64_t img = 0, hm = 0;
    if (image_id != nullptr) { img = image_id[i]; if (n_images > 0 && (img < 0 || img >= n_images)) bad = true; }
    if (obs_index != nullptr) { const int64_t o = obs_index[i]; if (o < 0 || o >= n_total) bad = true; }
    if (laue) { hm = harmonic_id[i]; if (hm < 0 || hm >= n) bad = true; }
    if (bad) { atomicMin(first_bad, (unsigned long long)i); key[i] = 0u; continue; }
    const uint32_t k = (uint32_t)(order == 1 ? r : order == 2 ? hm : img);
    key[i] = k;
    if (count != nullptr) atomicAdd(&count[k], 1u);
  }
}

// Digit histogram of one block's 4 096 rows -> G[digit][block] (digit-major: one exclusive scan over the whole table gives every
// (digit, block) its first output position).
__global__ void __launch_bounds__(kRsThreads) k_rs_hist(const uint32_t* key, int64_t n, int shift, uint32_t* G, int nblocks) {
  __shared__ uint32_t h[256];
  h[threadIdx.x] = 0u;
  __syncthreads();
  const int64_t base = (int64_t)blockIdx.x * kRsTile;
#pragma unroll
  for (int c = 0; c < kRsItems; ++c) {
    const int64_t i = base + c * kRsThreads + threadIdx.x;
    if (i < n) atomicAdd(&h[(key[i] >> shift) & 255u], 1u);
  }
  __syncthreads();
  G[(size_t)threadIdx.x * nblocks + blockIdx.x] = h[threadIdx.x];
}

// Stable scatter of one pass.  idx_in == nullptr: the first pass, the value of a row is its own index.
__global__ void __launch_bounds__(kRsThreads) k_rs_scatter(const uint32_t* key_in, const uint32_t* idx_in, int64_t n, int shift,
                                                           const uint32_t* G, int nblocks, uint32_t* key_out, uint32_t* idx_out) {
  __shared__ uint32_t wcnt[kRsWarps][257];          // [warp][digit]: running count, then first output position; bin 256 = rows past the end
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < kRsWarps * 257; i += kRsThreads) (&wcnt[0][0])[i] = 0u;
  __syncthreads();
  const int64_t wbase = (int64_t)blockIdx.x * kRsTile + (int64_t)warp * kRsWarpItems;
  const unsigned lt = (1u << lane) - 1u;
  uint32_t k[kRsItems]; uint32_t lr[kRsItems]; uint32_t v[kRsItems];
#pragma unroll
  for (int c = 0; c < kRsItems; ++c) {               // all of the warp's loads are in flight before the (serial) ranking starts
    const int64_t i = wbase + c * 32 + lane;
    k[c] = i < n ? __ldcs(&key_in[i]) : 0u;
    v[c] = (i < n && idx_in != nullptr) ? __ldcs(&idx_in[i]) : (uint32_t)i;
  }
#pragma unroll
  for (int c = 0; c < kRsItems; ++c) {
    const int64_t i = wbase + c * 32 + lane;
    const bool valid = i < n;
    const uint32_t kk = k[c];
    const uint32_t d = valid ? ((kk >> shift) & 255u) : 256u;
    const unsigned same = __match_any_sync(0xffffffffu, d);
    const uint32_t r = __popc(same & lt);
    const uint32_t b = wcnt[warp][d];
    __syncwarp();
    if (r == 0u) wcnt[warp][d] = b + __popc(same);
    __syncwarp();
    lr[c] = b + r;
  }
  __syncthreads();
  {                                                  // digit tid: first output position of every warp's rows of that digit
    uint32_t run = G[(size_t)tid * nblocks + blockIdx.x];
#pragma unroll
    for (int w = 0; w < kRsWarps; ++w) { const uint32_t t = wcnt[w][tid]; wcnt[w][tid] = run; run += t; }
  }
  __syncthreads();
#pragma unroll
  for (int c = 0; c < kRsItems; ++c) {
    const int64_t i = wbase + c * 32 + lane;
    if (i < n) {
      const uint32_t p = wcnt[warp][(k[c] >> shift) & 255u] + lr[c];
      key_out[p] = k[c];
      idx_out[p] = v[c];
    }
  }
}

// Exclusive prefix sum of one 2 048-element tile in place; the tile total goes to tile_sums[block].
__global__ void __launch_bounds__(kScanThreads) k_scan_tile(uint32_t* data, int64_t m, uint32_t* tile_sums) {
  __shared__ uint32_t wsum[kScanThreads / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)tid * kScanItems;
  uint32_t v[kScanItems], s = 0u;
#pragma unroll
  for (int j = 0; j < kScanItems; ++j) { v[j] = (base + j < m) ? data[base + j] : 0u; s += v[j]; }
  uint32_t inc = s;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
  if (lane == 31) wsum[warp] = inc;
  __syncthreads();
  uint32_t woff = 0u, total = 0u;
#pragma unroll
  for (int w = 0; w < kScanThreads / 32; ++w) { if (w < warp) woff += wsum[w]; total += wsum[w]; }
  uint32_t run = woff + inc - s;
#pragma unroll
  for (int j = 0; j < kScanItems; ++j) { if (base + j < m) data[base + j] = run; run += v[j]; }
  if (tid == 0) tile_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(256) k_scan_add(uint32_t* data, int64_t m, const uint32_t* tile_offsets) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < m) data[i] += tile_offsets[i / kScanTile];
}

// Image of the first row of every non-empty Laue spot (image layers: a new image starts a new tile); -1 for empty spots.
__global__ void __launch_bounds__(256) k_prep_first_image(int64_t n_keys, const uint32_t* count, const uint32_t* off, const uint32_t* perm,
                                                          const int64_t* image_id, int32_t* kimg) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n_keys) kimg[k] = count[k] > 0u ? (int32_t)image_id[perm[off[k]]] : -1;
}

// Padded position of every sorted row: its key's start position + its rank inside the key's run.
__global__ void __launch_bounds__(256) k_prep_pos(int64_t n, const uint32_t* skey, const uint32_t* off, const uint32_t* kpos, uint32_t* pos) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s < n) { const uint32_t k = skey[s]; pos[s] = kpos[k] + ((uint32_t)s - off[k]); }
}

__global__ void __launch_bounds__(256) k_fill_defaults(int64_t npad, int d, int32_t* refl, int32_t* img, int32_t* spot, uint32_t* oidx,
                                                       float* meta, float* iobs, float* sig) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npad) return;
  refl[p] = -1; oidx[p] = 0u; iobs[p] = 0.f; sig[p] = 1.f;
  if (img != nullptr) img[p] = 0;
  if (spot != nullptr) spot[p] = -1;
  for (int j = 0; j < d; ++j) meta[(size_t)j * npad + p] = 0.f;
}

// One thread per sorted row: gather the reference tuple's row perm[s] into padded position pos[s] (pos == nullptr: s itself).
__global__ void __launch_bounds__(256) k_fill_gather(int64_t n, int64_t npad, int d, const uint32_t* perm, const uint32_t* pos,
                                                     const int64_t* refl_id, const int64_t* image_id, const int64_t* harmonic_id,
                                                     const int64_t* obs_index, const float* metadata, const float* iobs_in, const float* sig_in,
                                                     int32_t* refl, int32_t* img, int32_t* spot, uint32_t* oidx, float* meta, float* iobs, float* sig) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  const int64_t i = perm[s];
  const int64_t p = pos != nullptr ? (int64_t)pos[s] : s;
  refl[p] = (int32_t)refl_id[i];
  if (img != nullptr) img[p] = (int32_t)image_id[i];
  oidx[p] = (uint32_t)(obs_index != nullptr ? obs_index[i] : i);
  for (int j = 0; j < d; ++j) meta[(size_t)j * npad + p] = metadata[(size_t)i * d + j];
  if (spot != nullptr) {
    const int64_t k = harmonic_id[i];
    spot[p] = (int32_t)k; iobs[p] = iobs_in[k]; sig[p] = sig_in[k];       // formatter.py:637-640: spot k's value sits at index k
  } else { iobs[p] = iobs_in[i]; sig[p] = sig_in[i]; }
}

}  // namespace prep
}  // namespace clb
