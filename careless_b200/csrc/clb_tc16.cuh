// clb_tc16.cuh -- tensor-core observation kernel for NARROW scale MLPs (padded width 16: the careless CLI default
// --mlp-width 10 lives here).  Same algorithm, same global layouts and the same epilogue as k_obs_tc2 (clb_kernels.cuh);
// what changes is the tiling: one thread per observation row carries all 16 features, a CTA is 128 threads / 128 rows,
// FOUR CTAs share an SM (128 tensor-memory columns, ~37 KB of shared memory, <= 128 registers each).
//   chain pass   : D[128 x 16] += A[128 x 16] (tensor memory, hi | lo) x B[16 x 16] (shared, K-major no-swizzle images by
//                  TMA); kind::tf32, K = 8 per instruction -> 2 k-steps x 3 products = 6 tcgen05.mma per pass;
//   dW product   : one 128-byte row per observation holds [x_hi (16) | x_lo (16)], so the MN-major SWIZZLE_128B_BASE32B
//                  image of `a` IS [a_hi | a_lo] (one 32-element MN group) and likewise for delta-p:
//                  D_dw[64 x 32] = [a_hi | a_lo | (junk group)]^T [dp_hi | dp_lo], 16 tcgen05.mma (M = 64, N = 32, K = 8);
//                  rows 0..15 (a_hi) sit in lanes 0..15 of warp 0, rows 16..31 (a_lo) in lanes 0..15 of warp 1; each folds
//                  its two 16-column blocks and REDs them into the CTA's FP32 partial.
// Included by clb_kernels.cuh (needs ObsArgs, obs_epilogue, bias_red16).
#pragma once

namespace clb {
namespace tc16 {

using namespace tc;

constexpr int kRows = 128;                         // rows (= threads) per CTA tile
constexpr uint32_t kLBO16 = 272;                   // bytes between K-adjacent core matrices of the 16 x 16 weight image
constexpr uint32_t kImg16 = 4 * kLBO16;            // one 16 x 16 tf32 operand image (1088 B)
constexpr uint32_t kCols16 = 128;                  // tensor-memory columns per CTA (four CTAs per SM)
constexpr uint32_t kA_hi = 0, kA_mid = 16, kA_lo = 32, kD = 48, kDdw = 64;   // forward: three-way split of X (hi, mid, lo)
constexpr uint32_t kDwImg16 = kRows * 128;         // one MN-major operand image: 128 rows x 128 B
constexpr uint32_t kIdesc16 = (1u << 4) | (2u << 7) | (2u << 10) | ((16u >> 3) << 17) | ((128u >> 4) << 24);
constexpr uint32_t kIdescDw16 = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((32u >> 3) << 17) | ((64u >> 4) << 24);
constexpr int PSLOT16 = 16 * 16 + 16;              // one layer of the CTA's FP32 partial
constexpr int PSLOT16_DET = 2 * 256 + 4 * 16;      // deterministic mode: [kernel from a_hi rows | from a_lo rows | bias per warp [4][16]]

__device__ __forceinline__ uint64_t make_desc16(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((kLBO16 >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((kSBO >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
// MN-major 128B/32B-base swizzle; LBO (distance to the second MN group, whose rows of D are never read) = one image
__device__ __forceinline__ uint64_t make_desc_mn16(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((kDwImg16 >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((kDwSBO >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46) | ((uint64_t)1 << 61);
}
__device__ __forceinline__ void mma16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n"
               :: "r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(kIdesc16), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_alloc16(uint32_t slot_smem) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(slot_smem), "r"(kCols16) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc16(uint32_t tbase) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tbase), "r"(kCols16) : "memory");
}
// images of one (layer, direction) in global memory: [hi, mid, lo][kImg16]; a forward pass fetches all three, a backward
// pass the first two... the backward images are stored as [hi, lo, unused]
__device__ __forceinline__ void tma_fetch16(uint32_t dst_smem, const float* src, uint32_t n_images, uint32_t bar) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(n_images * kImg16) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(dst_smem), "l"(src), "r"(n_images * kImg16), "r"(bar) : "memory");
}

// Three-way TF32 split for the FORWARD chain: hi = tf32(x), mid = tf32(x - hi), lo = the exact rest.  With the six
// products hi.lo, lo.hi, mid.mid, hi.mid, mid.hi, hi.hi the layer output is exact to ~2^-24 like an FP32 FMA chain:
// the surrogate gradients depend on the forward pass only (through the predicted intensity), and 20 narrow layers are
// conditioning-limited in FP32, so this is where 3xTF32 (2^-21) is not enough.
__device__ __forceinline__ void split16_3(const float (&x)[16], uint32_t (&hi)[16], uint32_t (&mid)[16], uint32_t (&lo)[16]) {
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const float h = tf32_rna(x[k]);
    const float r = x[k] - h;
    const float m = tf32_rna(r);
    hi[k] = __float_as_uint(h); mid[k] = __float_as_uint(m); lo[k] = __float_as_uint(r - m);
  }
}

// Rounded-hi split (3 instructions per value, 7e-7 rms per product against 1.4e-6 for the truncating split16): the chain
// passes of the narrow kernel use it -- 20 layers of width 10 are conditioning-limited in FP32 and the parity criterion
// (3 x the FP32 noise floor of the oracle) leaves no room for the cheaper split there; the dW operands keep split16.
__device__ __forceinline__ void split16_rna(const float (&x)[16], uint32_t (&hi)[16], uint32_t (&lo)[16]) {
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const float h = tf32_rna(x[k]);
    hi[k] = __float_as_uint(h);
    lo[k] = __float_as_uint(x[k] - h);
  }
}

struct Ctx16 {
  uint32_t row_addr, base;            // tensor memory: this thread's lane, the CTA's base
  uint32_t mbar, parity;              // chain barrier
  uint32_t mbar_dw, parity_dw;        // dW barrier
  uint32_t wimg0, wimg1, wbar0, wbar1;   // weight image buffers [hi | lo] and their TMA barriers
  uint32_t pass, wphase;
  char* dw_a; char* dw_b;
  uint32_t dwa_s, dwb_s;              // shared addresses of the operand images
  int tid;
};

// One row of an MN-major operand image: chunks 0, 1 = hi[0..7], hi[8..15], chunks 2, 3 = lo[0..7], lo[8..15]; with sw = 1
// the caller passes the two 4-value blocks of every chunk swapped and they land in the swapped halves (conflict-free).
__device__ __forceinline__ void dw_store_row16(char* img, int k, const uint32_t (&hi)[16], const uint32_t (&lo)[16], int sw) {
  const int r = k & 3;
  char* row = img + (size_t)(k >> 2) * kDwSBO + r * 128 + 16 * sw;
  const int odd = 16 - 32 * sw;
#pragma unroll
  for (int cc = 0; cc < 2; ++cc) {
    char* p = row + ((cc ^ r) * 32);
    char* q = row + (((2 + cc) ^ r) * 32);
    *reinterpret_cast<uint4*>(p) = make_uint4(hi[8 * cc], hi[8 * cc + 1], hi[8 * cc + 2], hi[8 * cc + 3]);
    *reinterpret_cast<uint4*>(p + odd) = make_uint4(hi[8 * cc + 4], hi[8 * cc + 5], hi[8 * cc + 6], hi[8 * cc + 7]);
    *reinterpret_cast<uint4*>(q) = make_uint4(lo[8 * cc], lo[8 * cc + 1], lo[8 * cc + 2], lo[8 * cc + 3]);
    *reinterpret_cast<uint4*>(q + odd) = make_uint4(lo[8 * cc + 4], lo[8 * cc + 5], lo[8 * cc + 6], lo[8 * cc + 7]);
  }
}

// All threads, after the pass's __syncthreads(): warp 0 waits for the layer's images and issues the MMAs: six products of
// the three-way split in a forward pass (12 MMAs), three products of the two-way split in a dX pass (6 MMAs).
// `next` / `next_n`: global images of the next pass and how many of them to prefetch.
__device__ __forceinline__ void issue_chain16(Ctx16& c, bool fwd, const float* next, uint32_t next_n, bool tma = true) {
  const uint32_t b = c.pass & 1u;
  const uint32_t warp = uniform32((uint32_t)c.tid >> 5);
  if (warp == 0u) {
    fence_after();
    const uint32_t base = uniform32(c.base);
    const uint32_t img = uniform32(b ? c.wimg1 : c.wimg0);
    const uint64_t w0 = make_desc16(img), w1 = make_desc16(img + kImg16), w2 = make_desc16(img + 2 * kImg16);
    const uint32_t d = base + kD;
    const uint32_t bar = uniform32(c.mbar);
    const uint32_t wb = uniform32(b ? c.wbar1 : c.wbar0), wph = uniform32((c.wphase >> b) & 1u);
    const uint32_t nb = uniform32(b ? c.wbar0 : c.wbar1), ndst = uniform32(b ? c.wimg0 : c.wimg1);
    if (elect_one()) {
      if (tma) mbar_wait(wb, wph);      // image layers: the threads built this pass's images themselves
#define CLB_MMA16(acol, wdesc, first) \
      mma16_ts(d, base + (acol), (wdesc), (first) ? 0u : 1u); \
      mma16_ts(d, base + (acol) + 8u, (wdesc) + (uint64_t)((2u * kLBO16) >> 4), 1u)
      if (fwd) {          // images [hi, mid, lo]
        CLB_MMA16(kA_hi, w2, true);  CLB_MMA16(kA_lo, w0, false); CLB_MMA16(kA_mid, w1, false);
        CLB_MMA16(kA_hi, w1, false); CLB_MMA16(kA_mid, w0, false); CLB_MMA16(kA_hi, w0, false);
      } else {            // images [hi, lo]; delta-p as (hi at kA_hi, lo at kA_lo -- CLB_ST32: at kA_mid, next to hi)
        CLB_MMA16(kA_hi, w1, true);  CLB_MMA16(CLB_ST32 ? kA_mid : kA_lo, w0, false); CLB_MMA16(kA_hi, w0, false);
      }
#undef CLB_MMA16
      commit(bar);
      if (next != nullptr) tma_fetch16(ndst, next, next_n, nb);
    }
    __syncwarp();
  }
  if (tma) c.wphase ^= (1u << b);
  c.pass += 1u;
}

// Image layers (scaling/image.py:66-125): the tile's own 16 x 16 kernel (FP32 [in][out] in shared memory) is turned into
// the operand images of this pass -- forward [hi, mid, lo] of B[n][k] = W[k][n], backward [hi, lo] of B[n][k] = W[n][k] --
// in buffer (pass & 1) by threads 0..63 (one 16-byte group of four k each).  The caller's fences + __syncthreads() publish them.
template <bool BWD>
__device__ __forceinline__ void build_images16(Ctx16& c, const float* Wsm, char* w_img) {
  if (c.tid < 64) {
    const int n = c.tid & 15, kq = c.tid >> 4;
    float w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) w[i] = BWD ? Wsm[n * 16 + 4 * kq + i] : Wsm[(4 * kq + i) * 16 + n];
    float4 hi, mid, lo;
    float* ph = &hi.x; float* pm = &mid.x; float* pl = &lo.x;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float h = tf32_rna(w[i]), r = w[i] - h, m = tf32_rna(r);
      ph[i] = h;
      if (BWD) { pm[i] = r; pl[i] = 0.f; } else { pm[i] = m; pl[i] = r - m; }
    }
    char* dst = w_img + (size_t)(c.pass & 1u) * 3 * kImg16 + kq * kLBO16 + (n >> 3) * kSBO + (n & 7) * 16;
    *reinterpret_cast<float4*>(dst) = hi;
    *reinterpret_cast<float4*>(dst + kImg16) = mid;
    if (!BWD) *reinterpret_cast<float4*>(dst + 2 * kImg16) = lo;
  }
}

__device__ __forceinline__ void issue_fwd16(Ctx16& c, const float (&x)[16], const float* next, uint32_t next_n,
                                            const float* build_from = nullptr, char* w_img = nullptr) {
  {
    uint32_t hi[16], mid[16], lo[16];
    split16_3(x, hi, mid, lo);
#if CLB_ST32
    {                                               // hi | mid sit in adjacent columns: one tcgen05.st.x32 + one .x16 instead of three .x16
      uint32_t v[32];
#pragma unroll
      for (int k = 0; k < 16; ++k) { v[k] = hi[k]; v[16 + k] = mid[k]; }
      CLB_TMEM_ST32(c.row_addr + kA_hi, v);
    }
#else
    CLB_TMEM_ST16(c.row_addr + kA_hi, hi);
    CLB_TMEM_ST16(c.row_addr + kA_mid, mid);
#endif
    CLB_TMEM_ST16(c.row_addr + kA_lo, lo);
  }
  if (build_from != nullptr) { build_images16<false>(c, build_from, w_img); fence_async_smem(); }
  wait_st();
  fence_before();
  __syncthreads();
  issue_chain16(c, true, next, next_n, build_from == nullptr);
}

__device__ __forceinline__ void collect16(Ctx16& c, float (&y)[16]) {
  mbar_wait(c.mbar, c.parity);
  c.parity ^= 1u;
  fence_after();
  uint32_t v[16];
  CLB_TMEM_LD16(c.row_addr + kD, v);
  wait_ld();
#pragma unroll
  for (int k = 0; k < 16; ++k) y[k] = __uint_as_float(v[k]);
}

__device__ __forceinline__ void issue_bwd16(Ctx16& c, const float (&dp)[16], const float (&ain)[16], bool need_dx, const float* next, uint32_t next_n,
                                            const float* build_from = nullptr, char* w_img = nullptr, bool ones15 = false) {
  {
    uint32_t hi[16], lo[16];
    split16_rna(dp, hi, lo);
    if (need_dx) {
#if CLB_ST32
      uint32_t v[32];                               // backward: the lo part goes to the (unused) mid columns next to hi: one .x32 store
#pragma unroll
      for (int k = 0; k < 16; ++k) { v[k] = hi[k]; v[16 + k] = lo[k]; }
      CLB_TMEM_ST32(c.row_addr + kA_hi, v);
#else
      CLB_TMEM_ST16(c.row_addr + kA_hi, hi);
      CLB_TMEM_ST16(c.row_addr + kA_lo, lo);
#endif
    }
    const int sw = (c.tid >> 2) & 1;
    swap_blocks(hi, sw); swap_blocks(lo, sw);
    dw_store_row16(c.dw_b, c.tid, hi, lo, sw);
    uint32_t a2[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) a2[k] = __float_as_uint(ain[k]);
    if (ones15) a2[15] = 0x3f800000u;      // the padding feature carries 1.0: row 15 of dW = sum over the tile's rows of delta-p = the bias gradient
    swap_blocks(a2, sw);
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      hi[k] = a2[k];
      lo[k] = __float_as_uint(__uint_as_float(a2[k]) - __uint_as_float(a2[k] & 0xFFFFE000u));
    }
    dw_store_row16(c.dw_a, c.tid, hi, lo, sw);
  }
  if (need_dx && build_from != nullptr) build_images16<true>(c, build_from, w_img);
  wait_st();
  fence_async_smem();
  fence_before();
  __syncthreads();
  if (need_dx) issue_chain16(c, false, next, next_n, build_from == nullptr);
  const uint32_t warp = uniform32((uint32_t)c.tid >> 5);
  if (warp == 3u) {
    fence_after();
    const uint32_t d = uniform32(c.base) + kDdw;
    const uint64_t a0 = make_desc_mn16(uniform32(c.dwa_s)), b0 = make_desc_mn16(uniform32(c.dwb_s));
    const uint32_t bar = uniform32(c.mbar_dw);
    if (elect_one()) {
#pragma unroll
      for (int ks = 0; ks < kRows / 8; ++ks)
        mma_tf32_ss(d, a0 + (uint64_t)((2u * kDwSBO * (uint32_t)ks) >> 4), b0 + (uint64_t)((2u * kDwSBO * (uint32_t)ks) >> 4), kIdescDw16, ks > 0 ? 1u : 0u);
      commit(bar);
    }
    __syncwarp();
  }
}

// D_dw rows 0..15 (a_hi) = lanes 0..15 of warp 0, rows 16..31 (a_lo) = lanes 0..15 of warp 1; 32 columns [.dp_hi | .dp_lo].
// il_w > 0: an image layer's kernel gradient, stored (out, in) with width il_w in that image's slot (scalar REDs).
// ones15: row 15 (lane 15 of warp 0) is the bias gradient (see issue_bwd16) and goes to bk instead of the kernel's slot.
__device__ __forceinline__ void collect_dw16(Ctx16& c, float* wk, int il_w = 0, int lo_off = 0, float* bk = nullptr, bool ones15 = false) {
  mbar_wait(c.mbar_dw, c.parity_dw);
  c.parity_dw ^= 1u;
  fence_after();
  const int warp = c.tid >> 5, lane = c.tid & 31;
  if (warp < 2) {
    uint32_t v[32];
    CLB_TMEM_LD32(c.row_addr + kDdw, v);
    wait_ld();
    if (ones15 && warp == 0 && lane == 15) {
      if (bk != nullptr) {
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (il_w == 0 || j < il_w) atomicAdd(&bk[j], __uint_as_float(v[j]) + __uint_as_float(v[16 + j]));
      }
    } else if (lane < 16 && wk != nullptr) {
      if (il_w == 0) {          // slot layout [j / 4][i][j % 4]: each RED instruction covers 256 contiguous bytes
        float4* dst = reinterpret_cast<float4*>(wk + (warp == 1 ? lo_off : 0)) + lane;      // warp 1 holds the a_lo rows
#pragma unroll
        for (int q = 0; q < 4; ++q)
          atomicAdd(dst + q * 16, make_float4(__uint_as_float(v[4 * q]) + __uint_as_float(v[16 + 4 * q]),
                                         __uint_as_float(v[4 * q + 1]) + __uint_as_float(v[16 + 4 * q + 1]),
                                         __uint_as_float(v[4 * q + 2]) + __uint_as_float(v[16 + 4 * q + 2]),
                                         __uint_as_float(v[4 * q + 3]) + __uint_as_float(v[16 + 4 * q + 3])));
      } else if (lane < il_w) {          // row i = lane (a_hi rows in warp 0, a_lo rows in warp 1) of dK[i = in][j = out]
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (j < il_w) atomicAdd(&wk[j * il_w + lane], __uint_as_float(v[j]) + __uint_as_float(v[16 + j]));
      }
    }
  }
}

}  // namespace tc16

struct ObsSmem16 {
  static size_t bytes(int n_layers, int n_img_layers = 0) {
    return 2 * (size_t)tc16::kDwImg16 + 6 * (size_t)tc16::kImg16 + 64 + sizeof(float) * (32 + (size_t)n_layers * 16 + (size_t)n_img_layers * (256 + 16))
           + 64 * sizeof(double) + 128;
  }
};

// IL = the model has image layers (their code is compiled out otherwise): rows are image-major, no image straddles a tile.
template <int LIK, bool IL>
__global__ void __launch_bounds__(tc16::kRows, 4) k_obs_tc16(ObsArgs a) {
  using namespace tc16;
  constexpr int WP = 16, TR = kRows, T = kRows, NC = 4;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const int NL = a.lay.n_layers, L = NL - 1, K = IL ? a.n_img_layers : 0, LT = L + K;
  char* dw_a = reinterpret_cast<char*>(smem_raw);
  char* dw_b = dw_a + kDwImg16;
  char* w_img = dw_b + kDwImg16;                                     // [2 buffers][hi, mid, lo][kImg16]
  uint64_t* bars = reinterpret_cast<uint64_t*>(w_img + 6 * kImg16);  // [0] chain, [1] dW, [2..3] image buffers
  uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 4);
  float* Whead = reinterpret_cast<float*>(slot + 8);                 // [16][2]
  float* bsm = Whead + 32;                                           // [NL][16]
  float* Wimg = bsm + (size_t)NL * WP;                               // [K][16][16] this tile's image-layer kernels as [in][out]
  float* bimg = Wimg + (size_t)K * WP * WP;                          // [K][16]
  double* red = reinterpret_cast<double*>(bimg + (size_t)K * WP + (((NL + K) * WP) & 1));

  const int tid = threadIdx.x, lane = tid & 31;
  if (tid == 0) { for (int i = 0; i < 4; ++i) tc::mbar_init(tc::smem_u32(bars + i), 1); }
  if (tid < 32) tmem_alloc16(tc::smem_u32(slot));
  tc::fence_before();
  for (int idx = tid; idx < WP * 2; idx += T) {
    const int i = idx / 2, j = idx % 2;
    Whead[idx] = (i < a.lay.in_dim[L] && j < a.lay.out_dim[L]) ? a.theta_mlp[a.lay.koff[L] + i * a.lay.out_dim[L] + j] : 0.f;
  }
  for (int idx = tid; idx < NL * WP; idx += T) {
    const int k = idx / WP, j = idx % WP;
    bsm[idx] = (j < a.lay.out_dim[k]) ? a.theta_mlp[a.lay.boff[k] + j] : 0.f;
  }
  __syncthreads();
  Ctx16 c{};
  {
    tc::fence_after();
    c.base = *slot;
    c.row_addr = c.base + ((uint32_t)(32 * (tid >> 5)) << 16);
    c.mbar = tc::smem_u32(bars); c.mbar_dw = tc::smem_u32(bars + 1);
    c.wbar0 = tc::smem_u32(bars + 2); c.wbar1 = tc::smem_u32(bars + 3);
    c.wimg0 = tc::smem_u32(w_img); c.wimg1 = tc::smem_u32(w_img + 3 * kImg16);
    c.dw_a = dw_a; c.dw_b = dw_b; c.dwa_s = tc::smem_u32(dw_a); c.dwb_s = tc::smem_u32(dw_b);
    c.tid = tid;
  }
  constexpr size_t IMGF = kImg16 / 4;
  // [hi, mid, lo] / [hi, lo, -] of hidden layer k; null for image layers, whose per-tile kernels the threads turn into images
  auto gimg = [&](int k, int dir) -> const float* { return (!IL || k < L) ? a.wimg + ((size_t)(k * 2 + dir) * 3) * IMGF : nullptr; };
  auto wsrc = [&](int k) -> const float* { return (IL && k >= L) ? Wimg + (size_t)(k - L) * WP * WP : nullptr; };
  const int PSLOT = a.det ? PSLOT16_DET : PSLOT16;
  const int BOFF = a.det ? 2 * WP * WP + 16 * (tid >> 5) : WP * WP;       // this warp's bias slot inside a layer's slot
  const int lo_off = a.det ? WP * WP : 0;
  const bool ones15 = a.bias_feat15 != 0;
  float* part32 = a.partials32 + (size_t)(blockIdx.x % a.n_partials) * NL * PSLOT;
  float4* scr = a.scratch + (size_t)blockIdx.x * LT * NC * TR;
  double ll_sum = 0.0;
  float ev_f = 1.f, ev_a = 0.f, ev_b = 0.f;
  if (a.theta_lik != nullptr) { ev_f = softplusf(a.theta_lik[0]); ev_a = softplusf(a.theta_lik[1]); ev_b = softplusf(a.theta_lik[2]); }
  const int64_t n_tiles = (a.n_rows + TR - 1) / TR;
  if (tid == 0 && LT > 0 && blockIdx.x < n_tiles) tma_fetch16(c.wimg0, gimg(0, 0), 3u, c.wbar0);

  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const bool more_tiles = tile + gridDim.x < n_tiles;
    const int64_t row = tile * TR + tid;
    const bool inb = row < a.n_rows;
    const int refl = inb ? __ldcs(&a.refl[row]) : -1;
    const bool active = refl >= 0;
    const int timg = (IL && K > 0) ? a.image[tile * TR] : 0;
    if (IL && K > 0) {
      __syncthreads();                       // the previous tile is done with Wimg
      const int w = a.il_width;
      const size_t lstride = (size_t)a.il_n_images * w * (w + 1);
      for (int idx = tid; idx < K * WP * WP; idx += T) {
        const int l = idx / (WP * WP), i = (idx / WP) % WP, j = idx % WP;
        Wimg[idx] = (i < w && j < w) ? a.theta_il[l * lstride + ((size_t)timg * w + j) * w + i] : 0.f;   // stored (out, in)
      }
      for (int idx = tid; idx < K * WP; idx += T) {
        const int l = idx / WP, j = idx % WP;
        bimg[idx] = (j < w) ? a.theta_il[l * lstride + (size_t)a.il_n_images * w * w + (size_t)timg * w + j] : 0.f;
      }
      __syncthreads();
    }
    // ---------------- forward ----------------
    float h[WP];
#pragma unroll
    for (int i = 0; i < WP; ++i) h[i] = (inb && i < a.d) ? __ldcs(&a.meta[(size_t)i * a.n_rows + row]) : 0.f;
    for (int k = 0; k < LT; ++k) {
      const float* bk = (IL && k >= L) ? bimg + (size_t)(k - L) * WP : bsm + (size_t)k * WP;
      const bool next_bwd = !(k + 1 < LT) && a.train_mlp && LT > 1;
      const float* next = (k + 1 < LT) ? gimg(k + 1, 0) : next_bwd ? gimg(LT - 1, 1) : (more_tiles ? gimg(0, 0) : nullptr);
      if (ones15) h[15] = 1.f;               // the padding feature: multiplies the bias row of a hidden layer's forward image
      issue_fwd16(c, h, next, next_bwd ? 2u : 3u, wsrc(k), w_img);
      float o[WP];
      collect16(c, o);
      const bool bias_in_product = ones15 && !(IL && k >= L);      // (image layers: per-tile images without a bias row)
#pragma unroll
      for (int j = 0; j < WP; ++j) { const float v = bias_in_product ? o[j] : o[j] + bk[j]; h[j] = fmaxf(v, kLeak * v); }
      if (a.train_mlp && k + 1 < LT) {          // the last layer's output stays in registers for the head
#pragma unroll
        for (int q = 0; q < 4; ++q) scr[((size_t)k * NC + q) * TR + tid] = make_float4(h[4 * q], h[4 * q + 1], h[4 * q + 2], h[4 * q + 3]);
      }
    }
    float out0 = bsm[L * WP], out1 = bsm[L * WP + 1];
#pragma unroll
    for (int i = 0; i < WP; ++i) {
      const float2 w = *reinterpret_cast<const float2*>(&Whead[i * 2]);
      out0 = fmaf(h[i], w.x, out0); out1 = fmaf(h[i], w.y, out1);
    }
    float dmu, drho;
    obs_epilogue<LIK>(a, row, inb, active, refl, lane, out0, out1, ev_f, ev_a, ev_b, ll_sum, dmu, drho);
    if (!a.train_mlp) continue;
    // ---------------- backward ----------------
    float dp[WP], nxt[WP];
    auto load_act = [&](float (&dst)[WP], int k) {
      if (k > 0) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 v = __ldcg(&scr[((size_t)(k - 1) * NC + q) * TR + tid]);
          dst[4 * q] = v.x; dst[4 * q + 1] = v.y; dst[4 * q + 2] = v.z; dst[4 * q + 3] = v.w;
        }
      } else {
#pragma unroll
        for (int i = 0; i < WP; ++i) dst[i] = (inb && i < a.d) ? __ldcs(&a.meta[(size_t)i * a.n_rows + row]) : 0.f;
      }
    };
    auto layer_backward = [&](const float (&ain)[WP], bool need_dx, const float* next, uint32_t next_n, float* wk, float* bk2, int il_w,
                              const float* build_from, const float4* dead) {
      const bool o15 = ones15 && il_w == 0;      // (hidden layers and the head; image layers keep the shuffle reduction: measured faster)
      issue_bwd16(c, dp, ain, need_dx, next, next_n, build_from, w_img, o15);
      // every warp has consumed `ain` (issue_bwd16 ends after a __syncthreads()): warp 2 drops the layer's 8 KB scratch slot from the L2
      if (dead != nullptr && (tid >> 5) == 2) { discard_line(reinterpret_cast<const char*>(dead) + (size_t)lane * 128);
                                                discard_line(reinterpret_cast<const char*>(dead) + (size_t)(32 + lane) * 128); }
      if (!o15) bias_red16(dp, bk2, lane, il_w > 0 ? il_w : 16);
      if (need_dx) {
        collect16(c, dp);                  // delta a_k
        // delta p_{k-1} = delta a_k * leaky'(pre-activation of layer k-1); sign(a_k) == sign(pre-activation), a_k is still in registers
#pragma unroll
        for (int j = 0; j < WP; ++j) dp[j] = ain[j] > 0.f ? dp[j] : kLeak * dp[j];
      }
      collect_dw16(c, wk, il_w, lo_off, bk2, o15);
    };
    if (LT > 0) load_act(nxt, LT - 1);
#pragma unroll
    for (int j = 0; j < WP; ++j) dp[j] = 0.f;
    dp[0] = dmu; dp[1] = drho;
    layer_backward(h, false, nullptr, 0u, part32 + (size_t)L * PSLOT, part32 + (size_t)L * PSLOT + BOFF, 0, nullptr, nullptr);     // head: dW_out = a_L^T [dmu, drho]
#pragma unroll
    for (int i = 0; i < WP; ++i) {         // delta a_LT from the head, times leaky' of the last hidden layer (sign of its output h)
      const float2 w = *reinterpret_cast<const float2*>(&Whead[i * 2]);
      const float da = w.x * dmu + w.y * drho;
      dp[i] = h[i] > 0.f ? da : kLeak * da;
    }
    for (int k = LT - 1; k >= 0; --k) {
      float ain[WP];
#pragma unroll
      for (int i = 0; i < WP; ++i) ain[i] = nxt[i];
      if (k > 0) load_act(nxt, k - 1);
      const float* next = (k > 1) ? gimg(k - 1, 1) : (more_tiles ? gimg(0, 0) : nullptr);
      const bool is_il = IL && k >= L;
      float* wk = part32 + (size_t)k * PSLOT; float* bk2 = wk + BOFF; int il_w = 0;
      if (is_il) {            // image layers send their gradient to the tile's image slot
        const int w = a.il_width;
        const size_t lstride = (size_t)a.il_n_images * w * (w + 1);
        wk = a.g_il != nullptr ? a.g_il + (size_t)(k - L) * lstride + (size_t)timg * w * w : nullptr;
        bk2 = a.g_il != nullptr ? a.g_il + (size_t)(k - L) * lstride + (size_t)a.il_n_images * w * w + (size_t)timg * w : nullptr;
        il_w = w;
      }
      const float4* dead = (a.discard_scratch && k > 0) ? scr + (size_t)(k - 1) * NC * TR : nullptr;
      layer_backward(ain, k > 0, next, (k > 1) ? 2u : 3u, wk, bk2, il_w, wsrc(k), dead);
    }
  }
  __syncthreads();
  ll_sum = warp_sum(ll_sum);
  if (lane == 0) red[tid >> 5] = ll_sum;
  tc::fence_before();
  __syncthreads();
  if (tid == 0) {
    double t = 0.0;
    for (int i = 0; i < T / 32; ++i) t += red[i];
    flush_ll(a.ll_part, a.acc, t);
  }
  if (tid < 32) tmem_dealloc16(*slot);
}

// Ready-made 16 x 16 B operand images of every hidden layer for k_obs_tc16: per layer [fwd, bwd][hi, lo][kImg16].
// bias15: the forward image's input row 15 (a padding feature that the kernel feeds with 1.0) holds the layer's BIAS, so the chain
// product already includes it (three-way split like the weights: exact to 2^-24).
__global__ void __launch_bounds__(256) k_pack_images16(const float* theta_mlp, MlpLayout lay, float* wimg, int bias15) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int L = lay.n_layers - 1;
  if (idx >= L * 2 * 256) return;
  const int layer = idx >> 9, dir = (idx >> 8) & 1, n = (idx >> 4) & 15, k = idx & 15;
  const int i = dir ? n : k, j = dir ? k : n;          // W[i = in][j = out]
  float w = (i < lay.in_dim[layer] && j < lay.out_dim[layer]) ? theta_mlp[lay.koff[layer] + i * lay.out_dim[layer] + j] : 0.f;
  if (bias15 && dir == 0 && i == 15 && j < lay.out_dim[layer]) w = theta_mlp[lay.boff[layer] + j];
  const float hi = tc::tf32_rna(w);
  const float r = w - hi;
  const float mid = tc::tf32_rna(r);
  const size_t IMGF = tc16::kImg16 / 4;
  float* base = wimg + ((size_t)(layer * 2 + dir) * 3) * IMGF;
  const uint32_t off = ((k >> 2) * tc16::kLBO16 + (n >> 3) * tc::kSBO + (n & 7) * 16 + (k & 3) * 4) / 4;
  base[off] = hi;
  if (dir == 0) { base[IMGF + off] = mid; base[2 * IMGF + off] = r - mid; }     // forward: [hi, mid, lo]
  else base[IMGF + off] = r;                                                     // backward: [hi, lo]
}

__global__ void __launch_bounds__(256) k_reduce_partials16(const float* partials, int rows, MlpLayout lay, float* grad, int det) {
  // 32 parameters per block, the rows split into 8 contiguous chunks (one per warp) that are summed in parallel and combined in a
  // fixed order: the same deterministic result whatever the timing, 8x shorter dependent-load chains (the one-thread-per-parameter
  // version took 52 us for 74 rows -- a sixth of a small problem's step)
  __shared__ double part[8][32];
  const int pl = threadIdx.x & 31, g = threadIdx.x >> 5;
  const int p = blockIdx.x * 32 + pl;
  const bool live = p < lay.n_params;
  const int PSLOT = det ? tc16::PSLOT16_DET : tc16::PSLOT16, PP = lay.n_layers * PSLOT;
  const int BIAS = det ? 512 : 256;
  int extra = 0, n_extra = 0;
  int src = -1;
  for (int k = 0; live && k < lay.n_layers; ++k) {
    const int nk = lay.in_dim[k] * lay.out_dim[k];
    if (p >= lay.koff[k] && p < lay.koff[k] + nk) {
      const int i = (p - lay.koff[k]) / lay.out_dim[k], j = (p - lay.koff[k]) % lay.out_dim[k];
      src = k * PSLOT + (((j >> 2) * 16 + i) << 2) + (j & 3);
      if (det) { extra = 256; n_extra = 1; }
      break;
    }
    if (p >= lay.boff[k] && p < lay.boff[k] + lay.out_dim[k]) { src = k * PSLOT + BIAS + (p - lay.boff[k]); if (det) { extra = 16; n_extra = 3; } break; }
  }
  double acc = 0.0;
  const int per = (rows + 7) / 8, r0 = g * per, r1 = min(rows, r0 + per);
  if (live && src >= 0) for (int r = r0; r < r1; ++r)
    for (int e = 0; e <= n_extra; ++e) acc += (double)partials[(size_t)r * PP + src + e * extra];
  part[g][pl] = acc;
  __syncthreads();
  if (g == 0 && live) {
    double t = 0.0;
#pragma unroll
    for (int q = 0; q < 8; ++q) t += part[q][pl];
    grad[p] = (float)t;
  }
}

}  // namespace clb
