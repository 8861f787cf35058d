// clb_tc.cuh -- tcgen05 / TMEM building blocks for the scale-MLP products of the observation kernels (sm_100a only).
//
// One "pass" multiplies the CTA's 128 x 32 activation (or delta) tile by one 32 x 32 weight matrix on the
// 5th-generation tensor cores, error-compensated 3xTF32 so that the result keeps FP32-level accuracy:
//     X W  ~=  X_hi W_lo + X_lo W_hi + X_hi W_hi          (hi = the upper 19 bits the MMA reads, lo = the exact rest)
// * A operand (X): written to TENSOR MEMORY with tcgen05.st (lane = row) -- no layout puzzle;
// * B operand (W): shared-memory images in the canonical K-major / no-swizzle UMMA layout (8 x 16-byte core matrices,
//   LBO = 528 B between K-adjacent core matrices, SBO = 128 B between 8-row groups); k_obs_tc2 receives them ready-made
//   by TMA (cp.async.bulk + mbarrier), the older k_obs<32, LIK, true> builds them per pass;
// * D (FP32 accumulators): tensor memory, one accumulator region for all three products (12 MMAs, K = 8 each, issued
//   back to back by ONE elected lane with warp-uniform operands), read back with tcgen05.ld;
// * dW = A^T dP: both operands from shared memory in the MN-major SWIZZLE_128B_BASE32B layout, M = N = 64.
// First half of the file: helpers shared by all kernels and the one-thread-per-row generation (k_obs<32, LIK, true>,
// 128-thread CTAs); second half: two threads per row (k_obs_tc2, 256-thread CTAs).  clb_tc16.cuh: narrow models.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace clb {
namespace tc {

constexpr uint32_t kLBO = 528;           // bytes between K-adjacent core matrices (528 % 128 == 16: conflict-free image builds)
constexpr uint32_t kSBO = 128;           // bytes between 8-row groups
constexpr uint32_t kImgBytes = 8 * kLBO; // one 32x32 tf32 operand image
constexpr uint32_t kTmemCols = 256;      // per CTA (two CTAs per SM)
constexpr int kThreads = 128;            // rows per CTA tile
// tensor-memory column map of the CTA
constexpr uint32_t kColAhi = 0, kColAlo = 32, kColD = 64;   // D0, D1, D2 at 64, 96, 128; dW accumulators at 160
// dW = A^T dP (contraction over the 128 observations of the tile): both operands come from shared memory in the
// MN-major layout, which for 32-bit types exists only as SWIZZLE_128B_BASE32B (layout type 1): atoms of 4 K-rows
// x 128 B (32 MN elements), 32-byte chunk index XOR (row % 4); SBO = 512 B between 4-row K groups, LBO = 16 KB
// between 32-element MN groups (group 0 = hi parts, group 1 = lo parts).  M = N = 64, K = 8 per instruction:
// D = [A_hi; A_lo]^T [dP_hi | dP_lo]  ->  dW = D[0:32,0:32] + D[0:32,32:64] + D[32:64,0:32] (+ lo*lo).
// M = 64 accumulators occupy lanes (r % 16) + 32 (r / 16) (verified with tools/tc_probe2.cu).
constexpr uint32_t kDwSBO = 512, kDwLBO = (kThreads / 4) * 512;
constexpr uint32_t kDwImgBytes = 2 * kDwLBO;                 // 32 KB per operand
constexpr uint32_t kColDw = 160;                              // 64 accumulator columns of the dW product
constexpr uint32_t kIdescDw = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((64u >> 3) << 17) | ((64u >> 4) << 24);
constexpr int kStageStride = 36;                             // floats per row of the dW reduction stage (padded)
// instruction descriptor: D=F32 (1<<4), A=B=TF32 (2<<7, 2<<10), both K-major, N=32 (4<<17), M=128 (8<<24)
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((32u >> 3) << 17) | ((128u >> 4) << 24);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Round to the nearest TF32 (10-bit mantissa), ties away from zero: integer add + mask (2 instructions;
// cvt.rna.tf32.f32 expands to ~5).  Overflow into the exponent is the correct rounding; inf/nan stay inf/nan-like.
__device__ __forceinline__ float tf32_rna(float x) {
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
}

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((kLBO >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((kSBO >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);   // version 1, no swizzle
}

__device__ __forceinline__ uint64_t make_desc_mn(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((kDwLBO >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((kDwSBO >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46) | ((uint64_t)1 << 61);   // SWIZZLE_128B_BASE32B
}

__device__ __forceinline__ void mma_tf32_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
               :: "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}

__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n"
               :: "r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(kIdesc), "r"(accumulate) : "memory");
}

// Warp-uniform helpers: the compiler keeps values produced by a constant-lane shuffle in uniform registers,
// so the tcgen05.mma operands need no per-instruction R2UR "waterfall" loop (which costs ~100 cycles per MMA).
__device__ __forceinline__ uint32_t uniform32(uint32_t x) { return __shfl_sync(0xffffffffu, x, 0); }
__device__ __forceinline__ uint64_t uniform64(uint64_t x) {
  return ((uint64_t)__shfl_sync(0xffffffffu, (uint32_t)(x >> 32), 0) << 32) | (uint64_t)__shfl_sync(0xffffffffu, (uint32_t)x, 0);
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ void commit(uint32_t mbar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(mbar) : "memory");
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_init(uint32_t mbar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(mbar), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
  asm volatile("{\n\t.reg .pred P1;\n\tCLB_WAIT:\n\t"
               "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
               "@P1 bra CLB_DONE;\n\tbra CLB_WAIT;\n\tCLB_DONE:\n\t}\n" :: "r"(mbar), "r"(parity) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t slot_smem) {   // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(slot_smem), "r"(kTmemCols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t tbase) {     // the same warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tbase), "r"(kTmemCols) : "memory");
}

#define CLB_TMEM_ST32(taddr, v) asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" \
  :: "r"(taddr), "r"(v[0]),"r"(v[1]),"r"(v[2]),"r"(v[3]),"r"(v[4]),"r"(v[5]),"r"(v[6]),"r"(v[7]),"r"(v[8]),"r"(v[9]),"r"(v[10]),"r"(v[11]),"r"(v[12]),"r"(v[13]),"r"(v[14]),"r"(v[15]), \
     "r"(v[16]),"r"(v[17]),"r"(v[18]),"r"(v[19]),"r"(v[20]),"r"(v[21]),"r"(v[22]),"r"(v[23]),"r"(v[24]),"r"(v[25]),"r"(v[26]),"r"(v[27]),"r"(v[28]),"r"(v[29]),"r"(v[30]),"r"(v[31]) : "memory")

#define CLB_TMEM_LD32(taddr, v) asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];" \
  : "=r"(v[0]),"=r"(v[1]),"=r"(v[2]),"=r"(v[3]),"=r"(v[4]),"=r"(v[5]),"=r"(v[6]),"=r"(v[7]),"=r"(v[8]),"=r"(v[9]),"=r"(v[10]),"=r"(v[11]),"=r"(v[12]),"=r"(v[13]),"=r"(v[14]),"=r"(v[15]), \
    "=r"(v[16]),"=r"(v[17]),"=r"(v[18]),"=r"(v[19]),"=r"(v[20]),"=r"(v[21]),"=r"(v[22]),"=r"(v[23]),"=r"(v[24]),"=r"(v[25]),"=r"(v[26]),"=r"(v[27]),"=r"(v[28]),"=r"(v[29]),"=r"(v[30]),"=r"(v[31]) \
  : "r"(taddr) : "memory")

// Per-thread view of the CTA's tensor-core state.
struct Ctx {
  uint32_t row_addr;     // tensor-memory address of this thread's row: base + (32 warp) << 16
  uint32_t mbar;         // shared-memory address of the pass barrier (3 arrivals per pass)
  uint32_t parity;
  char* img_hi; char* img_lo;          // B operand images in shared memory
  uint64_t desc_hi, desc_lo;
  int tid;
  // dW
  uint32_t mbar_dw, parity_dw;         // barrier of the dW product (1 arrival)
  char* dw_a; char* dw_b;              // MN-major operand images (kDwImgBytes each); dw_a doubles as the reduction stage
  uint64_t desc_dwa, desc_dwb;
  uint32_t base;                       // tensor-memory base address
  // two-threads-per-row kernels (k_obs_tc2): this thread's tile row, feature half and its column offset (16 hf)
  int row, hf; uint32_t col;
  // weight images delivered by TMA (k_obs_tc2): two shared-memory buffers [hi | lo], their descriptors and barriers
  uint64_t wdesc_hi[2], wdesc_lo[2];
  uint32_t wimg[2], wbar[2];
  uint32_t pass, wphase;               // chain passes started so far (buffer = pass & 1); phase bit of each buffer's barrier
  // barrier-free hand-over (k_obs_tc2): arrival counters in shared memory; the deferred dW collection of the previous layer
  uint32_t cnt_chain, cnt_dw, n_warps_m1;
  float* pend_wk; float* pend_bk; int pend_ilw; bool dw_pending;
  int lo_off;                          // deterministic mode: float offset of the slot that takes the a_lo rows of a kernel gradient (else 0)
};

__device__ __forceinline__ void split32(const float (&x)[32], uint32_t (&hi)[32], uint32_t (&lo)[32]) {
#pragma unroll
  for (int k = 0; k < 32; ++k) {
    const float h = tf32_rna(x[k]);
    hi[k] = __float_as_uint(h);
    lo[k] = __float_as_uint(x[k] - h);      // exact; the MMA ignores the low 13 bits (measured: tools/tc_probe mode 5)
  }
}

// One row (observation k) of an MN-major operand image: 32 hi values in MN group 0, 32 lo values in group 1.
__device__ __forceinline__ void dw_store_row(char* img, int k, const uint32_t (&hi)[32], const uint32_t (&lo)[32]) {
  const int r = k & 3;
  char* row = img + (size_t)(k >> 2) * kDwSBO + r * 128;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    char* p = row + ((c ^ r) * 32);
    *reinterpret_cast<uint4*>(p) = make_uint4(hi[8 * c], hi[8 * c + 1], hi[8 * c + 2], hi[8 * c + 3]);
    *reinterpret_cast<uint4*>(p + 16) = make_uint4(hi[8 * c + 4], hi[8 * c + 5], hi[8 * c + 6], hi[8 * c + 7]);
    *reinterpret_cast<uint4*>(p + kDwLBO) = make_uint4(lo[8 * c], lo[8 * c + 1], lo[8 * c + 2], lo[8 * c + 3]);
    *reinterpret_cast<uint4*>(p + kDwLBO + 16) = make_uint4(lo[8 * c + 4], lo[8 * c + 5], lo[8 * c + 6], lo[8 * c + 7]);
  }
}

// The 8 weights this thread contributes to the B operand image of one layer, fetched from the padded FP32
// copy in global memory (L2/L1 resident, 4 KB per layer; image layers: the tile's copy in shared memory, hence
// generic loads) one pass ahead of their use.
//   forward  (B[n][k] = W[k][n]): thread (n = tid%32, kq = tid/32 and kq+4) gathers 4 consecutive k each;
//   backward (B[n][k] = W[n][k]): thread (kq = tid%8, n = tid/8 and n+16) copies 4 consecutive k each.
template <bool BWD>
__device__ __forceinline__ void load_w(const float* Wg, int tid, float (&w)[8]) {
  if (!BWD) {
    const int n = tid & 31, kq = tid >> 5;
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int r = 0; r < 4; ++r) w[4 * h + r] = Wg[(4 * (kq + 4 * h) + r) * 32 + n];
  } else {
    const int kq = tid & 7, n = tid >> 3;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const float4 v = *reinterpret_cast<const float4*>(&Wg[(n + 16 * h) * 32 + 4 * kq]);
      w[4 * h] = v.x; w[4 * h + 1] = v.y; w[4 * h + 2] = v.z; w[4 * h + 3] = v.w;
    }
  }
}

// Build the chain's B operand image (hi and lo) from the prefetched weights.
// element (n, k) of the [N][K] operand at (k/4) LBO + (n/8) SBO + (n%8) 16 + (k%4) 4 bytes
template <bool BWD>
__device__ __forceinline__ void build_weight_image(Ctx& c, const float (&w)[8]) {
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    int n, kq;
    if (!BWD) { n = c.tid & 31; kq = (c.tid >> 5) + 4 * h; }
    else { kq = c.tid & 7; n = (c.tid >> 3) + 16 * h; }
    const uint32_t off = kq * kLBO + (n >> 3) * kSBO + (n & 7) * 16;
    float4 hi, lo;
    hi.x = tf32_rna(w[4 * h]); hi.y = tf32_rna(w[4 * h + 1]); hi.z = tf32_rna(w[4 * h + 2]); hi.w = tf32_rna(w[4 * h + 3]);
    lo.x = w[4 * h] - hi.x; lo.y = w[4 * h + 1] - hi.y;
    lo.z = w[4 * h + 2] - hi.z; lo.w = w[4 * h + 3] - hi.w;
    *reinterpret_cast<float4*>(c.img_hi + off) = hi;
    *reinterpret_cast<float4*>(c.img_lo + off) = lo;
  }
}

// The chain issuer: warp 0 issues the 12 tcgen05.mma of the three products back to back into ONE accumulator
// (X_hi W_lo, X_lo W_hi first, then X_hi W_hi), so that collect() needs a single tensor-memory load.
__device__ __forceinline__ void issue_chain_mmas(Ctx& c) {
  const uint32_t warp = uniform32((uint32_t)c.tid >> 5);
  if (warp == 0u) {
    fence_after();
    const uint32_t base = uniform32(c.base);
    const uint64_t bhi = uniform64(c.desc_hi), blo = uniform64(c.desc_lo);
    const uint32_t d = base + kColD;
    const uint32_t bar = uniform32(c.mbar);
    if (elect_one()) {
#ifndef CLB_ABL_CHAIN
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
        mma_tf32_ts(d, base + kColAhi + 8u * (uint32_t)ks, blo + (uint64_t)((2u * kLBO * (uint32_t)ks) >> 4), ks > 0 ? 1u : 0u);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
        mma_tf32_ts(d, base + kColAlo + 8u * (uint32_t)ks, bhi + (uint64_t)((2u * kLBO * (uint32_t)ks) >> 4), 1u);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
        mma_tf32_ts(d, base + kColAhi + 8u * (uint32_t)ks, bhi + (uint64_t)((2u * kLBO * (uint32_t)ks) >> 4), 1u);
#endif
      commit(bar);
    }
    __syncwarp();
  }
}

// Backward of one layer: start delta_a = dp W^T (if need_dx; w = prefetched weights) and dW = ain^T dp on the
// tensor cores.  Contains one __syncthreads().  collect() returns delta_a; collect_dw() the weight gradient.
__device__ __forceinline__ void issue_backward(Ctx& c, const float (&dp)[32], const float (&ain)[32], const float (&w)[8], bool need_dx) {
  {
    uint32_t hi[32], lo[32];
    split32(dp, hi, lo);
    if (need_dx) {
      CLB_TMEM_ST32(c.row_addr + kColAhi, hi);
      CLB_TMEM_ST32(c.row_addr + kColAlo, lo);
    }
    dw_store_row(c.dw_b, c.tid, hi, lo);
    split32(ain, hi, lo);
    dw_store_row(c.dw_a, c.tid, hi, lo);
  }
  if (need_dx) build_weight_image<true>(c, w);
  wait_st();
  fence_async_smem();
  fence_before();
  __syncthreads();
  if (need_dx) issue_chain_mmas(c);
  const uint32_t warp = uniform32((uint32_t)c.tid >> 5);
  if (warp == 3u) {                                 // the dW issuer: 16 k-steps over the tile's 128 observations
    fence_after();
    const uint32_t d = uniform32(c.base) + kColDw;
    const uint64_t a0 = uniform64(c.desc_dwa), b0 = uniform64(c.desc_dwb);
    const uint32_t bar = uniform32(c.mbar_dw);
    if (elect_one()) {
#pragma unroll
      for (int ks = 0; ks < kThreads / 8; ++ks)
        mma_tf32_ss(d, a0 + (uint64_t)((2u * kDwSBO * (uint32_t)ks) >> 4), b0 + (uint64_t)((2u * kDwSBO * (uint32_t)ks) >> 4), kIdescDw, ks > 0 ? 1u : 0u);
      commit(bar);
    }
    __syncwarp();
  }
}

// Wait for dW and combine: D rows live at lanes (r%16) + 32 (r/16); every thread with lane < 16 folds its row's
// two column halves and parks it in the stage [r >= 32][i = r % 32][kStageStride] (aliases dw_a).
// After the caller's __syncthreads() the stage holds the 2 partial copies of the 32x32 gradient.
__device__ __forceinline__ void collect_dw(Ctx& c) {
  mbar_wait(c.mbar_dw, c.parity_dw);
  c.parity_dw ^= 1u;
  fence_after();
  const int warp = c.tid >> 5, lane = c.tid & 31;
  const uint32_t addr = c.row_addr + kColDw;
  uint32_t v0[32], v1[32];
  CLB_TMEM_LD32(addr, v0);
  CLB_TMEM_LD32(addr + 32, v1);
  wait_ld();
  if (lane < 16) {
    const int r = 16 * warp + lane;
    float* dst = reinterpret_cast<float*>(c.dw_a) + ((size_t)(r >> 5) * 32 + (r & 31)) * kStageStride;
#pragma unroll
    for (int q = 0; q < 8; ++q)
      *reinterpret_cast<float4*>(dst + 4 * q) = make_float4(__uint_as_float(v0[4 * q]) + __uint_as_float(v1[4 * q]),
                                                            __uint_as_float(v0[4 * q + 1]) + __uint_as_float(v1[4 * q + 1]),
                                                            __uint_as_float(v0[4 * q + 2]) + __uint_as_float(v1[4 * q + 2]),
                                                            __uint_as_float(v0[4 * q + 3]) + __uint_as_float(v1[4 * q + 3]));
  }
}

// Start one pass:  Y = X W_k (BWD = false)  or  Y = X W_k^T (BWD = true) for the CTA's 128 rows.
// x: this thread's row; w: the thread's share of the layer's weights (load_w<BWD>).
// Ends with the tcgen05.mma's in flight; tc::collect() waits for them.  Contains one __syncthreads().
template <bool BWD>
__device__ __forceinline__ void issue(Ctx& c, const float (&x)[32], const float (&w)[8]) {
  {
    uint32_t hi[32], lo[32];
    split32(x, hi, lo);
    CLB_TMEM_ST32(c.row_addr + kColAhi, hi);
    CLB_TMEM_ST32(c.row_addr + kColAlo, lo);
  }
  build_weight_image<BWD>(c, w);
  wait_st();
  fence_async_smem();           // generic-proxy smem writes -> visible to the tensor core's async proxy
  fence_before();
  __syncthreads();
  issue_chain_mmas(c);
}

// Wait for the pass and read this thread's row of the result.
__device__ __forceinline__ void collect(Ctx& c, float (&y)[32]) {
  mbar_wait(c.mbar, c.parity);
  c.parity ^= 1u;
  fence_after();
  uint32_t v[32];
  CLB_TMEM_LD32(c.row_addr + kColD, v);
  wait_ld();
#pragma unroll
  for (int k = 0; k < 32; ++k) y[k] = __uint_as_float(v[k]);
}


// ---------------------------------------------------------------------------------------
// Two threads per observation row (k_obs_tc2): 256-thread CTAs, thread (row = tid % 128, hf = tid / 128) owns
// features [16 hf, 16 hf + 16) of its row.  Warps w and w + 4 share the tensor-memory lanes 32 (w % 4) .. + 31 and use
// different columns.  Twice the warps per SM for the same shared / tensor memory, half the serial work per thread.
// ---------------------------------------------------------------------------------------
constexpr int kThreads2 = 256;
#define CLB_TMEM_ST16(taddr, v) asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" \
  :: "r"(taddr), "r"(v[0]),"r"(v[1]),"r"(v[2]),"r"(v[3]),"r"(v[4]),"r"(v[5]),"r"(v[6]),"r"(v[7]),"r"(v[8]),"r"(v[9]),"r"(v[10]),"r"(v[11]),"r"(v[12]),"r"(v[13]),"r"(v[14]),"r"(v[15]) : "memory")
#define CLB_TMEM_LD16(taddr, v) asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];" \
  : "=r"(v[0]),"=r"(v[1]),"=r"(v[2]),"=r"(v[3]),"=r"(v[4]),"=r"(v[5]),"=r"(v[6]),"=r"(v[7]),"=r"(v[8]),"=r"(v[9]),"=r"(v[10]),"=r"(v[11]),"=r"(v[12]),"=r"(v[13]),"=r"(v[14]),"=r"(v[15]) \
  : "r"(taddr) : "memory")

// TF32 split with two instructions per value: the MMA reads only the upper 19 bits of an operand (tools/tc_probe: raw
// FP32 bits and explicitly truncated bits give bit-identical products), so the hi part is x itself and
// lo = x - trunc(x) (exact).  3xTF32 with truncated parts is 1.4e-6 rms relative per product against 7e-7 with
// rounded hi parts (tc_probe modes 4 / 5).
__device__ __forceinline__ void split16(const float (&x)[16], uint32_t (&hi)[16], uint32_t (&lo)[16]) {
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    hi[k] = __float_as_uint(x[k]);
    lo[k] = __float_as_uint(x[k] - __uint_as_float(hi[k] & 0xFFFFE000u));
  }
}

// Swap the two 4-value blocks of each 8-value chunk when sw != 0 (two selects per pair of values).
__device__ __forceinline__ void swap_blocks(uint32_t (&v)[16], int sw) {
#pragma unroll
  for (int cc = 0; cc < 2; ++cc) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint32_t x = v[8 * cc + i], y = v[8 * cc + 4 + i];
      v[8 * cc + i] = sw ? y : x;
      v[8 * cc + 4 + i] = sw ? x : y;
    }
  }
}

// This thread's 16 features (32-byte chunks 2 hf and 2 hf + 1) of row k of an MN-major operand image.
// PACKED (the delta-p image of k_obs_tc2, CLB_DWB_PACK): MN group hf holds [hi of features 16 hf .. | lo of the same features], so that
// a thread's 32 accumulator columns of the dW product are adjacent (one tcgen05.ld.x32 instead of two .x16).
template <bool PACKED = false>
__device__ __forceinline__ void dw_store_half(char* img, int k, int hf, const uint32_t (&hi)[16], const uint32_t (&lo)[16], int sw) {
#ifdef CLB_ABL_STS
  if (hi[0] == 0x7fc01234u && lo[3] == 0x7fc04321u) *reinterpret_cast<uint32_t*>(img) = hi[1] ^ lo[2] ^ hi[15] ^ lo[15] ^ hi[8] ^ lo[8];
  return;
#endif
  const int r = k & 3;
  // sw = 1 (lanes with (lane >> 2) odd): the caller passes the two 16-byte blocks of every chunk SWAPPED in the register
  // arrays and they are stored to the swapped halves, so that one instruction's 32 lanes cover both halves of the
  // chunks: 4 wavefronts per STS.128 instead of 8.
  char* row = img + (size_t)(k >> 2) * kDwSBO + r * 128 + 16 * sw;
  const int odd = 16 - 32 * sw;
  if (PACKED) {
    row += (size_t)hf * kDwLBO;
#pragma unroll
    for (int cc = 0; cc < 2; ++cc) {
      char* p = row + ((cc ^ r) * 32);
      char* q = row + (((2 + cc) ^ r) * 32);
      *reinterpret_cast<uint4*>(p) = make_uint4(hi[8 * cc], hi[8 * cc + 1], hi[8 * cc + 2], hi[8 * cc + 3]);
      *reinterpret_cast<uint4*>(p + odd) = make_uint4(hi[8 * cc + 4], hi[8 * cc + 5], hi[8 * cc + 6], hi[8 * cc + 7]);
      *reinterpret_cast<uint4*>(q) = make_uint4(lo[8 * cc], lo[8 * cc + 1], lo[8 * cc + 2], lo[8 * cc + 3]);
      *reinterpret_cast<uint4*>(q + odd) = make_uint4(lo[8 * cc + 4], lo[8 * cc + 5], lo[8 * cc + 6], lo[8 * cc + 7]);
    }
    return;
  }
#pragma unroll
  for (int cc = 0; cc < 2; ++cc) {
    char* p = row + (((2 * hf + cc) ^ r) * 32);
    *reinterpret_cast<uint4*>(p) = make_uint4(hi[8 * cc], hi[8 * cc + 1], hi[8 * cc + 2], hi[8 * cc + 3]);
    *reinterpret_cast<uint4*>(p + odd) = make_uint4(hi[8 * cc + 4], hi[8 * cc + 5], hi[8 * cc + 6], hi[8 * cc + 7]);
    *reinterpret_cast<uint4*>(p + kDwLBO) = make_uint4(lo[8 * cc], lo[8 * cc + 1], lo[8 * cc + 2], lo[8 * cc + 3]);
    *reinterpret_cast<uint4*>(p + kDwLBO + odd) = make_uint4(lo[8 * cc + 4], lo[8 * cc + 5], lo[8 * cc + 6], lo[8 * cc + 7]);
  }
}
#ifndef CLB_DWB_PACK
#define CLB_DWB_PACK 0
#endif
#ifndef CLB_STS_DIVERGE
#define CLB_STS_DIVERGE 0   // 1: the conflict-free store order of the dW operand images (odd lane groups store the second 16-byte block of a chunk
                            // first) by a two-way predicated store sequence instead of swapping the register blocks with selects (-68 SEL, +32
                            // half-populated STS.128 per thread and layer).  Parity-green; measured MUCH slower on B200: 19.77 vs 16.23 ms --
                            // a shared-memory store instruction costs its full wavefronts whatever the number of active lanes
#endif
// Same bytes and the same conflict-free instruction pairing as swap_blocks + dw_store_half, without the 48 selects per thread and layer:
// lanes with sw = 0 store (block 0, block 1), lanes with sw = 1 store (block 1, block 0), each under its own predicate.
__device__ __forceinline__ void dw_store_half_div(char* img, int k, int hf, const uint32_t (&hi)[16], const uint32_t (&lo)[16]) {
  const int r = k & 3;
  char* row = img + (size_t)(k >> 2) * kDwSBO + r * 128;
  char* p0 = row + (((2 * hf) ^ r) * 32);
  char* p1 = row + (((2 * hf + 1) ^ r) * 32);
  if (((k >> 2) & 1) == 0) {
    *reinterpret_cast<uint4*>(p0) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(p0 + 16) = make_uint4(hi[4], hi[5], hi[6], hi[7]);
    *reinterpret_cast<uint4*>(p0 + kDwLBO) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    *reinterpret_cast<uint4*>(p0 + kDwLBO + 16) = make_uint4(lo[4], lo[5], lo[6], lo[7]);
    *reinterpret_cast<uint4*>(p1) = make_uint4(hi[8], hi[9], hi[10], hi[11]);
    *reinterpret_cast<uint4*>(p1 + 16) = make_uint4(hi[12], hi[13], hi[14], hi[15]);
    *reinterpret_cast<uint4*>(p1 + kDwLBO) = make_uint4(lo[8], lo[9], lo[10], lo[11]);
    *reinterpret_cast<uint4*>(p1 + kDwLBO + 16) = make_uint4(lo[12], lo[13], lo[14], lo[15]);
  } else {
    *reinterpret_cast<uint4*>(p0 + 16) = make_uint4(hi[4], hi[5], hi[6], hi[7]);
    *reinterpret_cast<uint4*>(p0) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(p0 + kDwLBO + 16) = make_uint4(lo[4], lo[5], lo[6], lo[7]);
    *reinterpret_cast<uint4*>(p0 + kDwLBO) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    *reinterpret_cast<uint4*>(p1 + 16) = make_uint4(hi[12], hi[13], hi[14], hi[15]);
    *reinterpret_cast<uint4*>(p1) = make_uint4(hi[8], hi[9], hi[10], hi[11]);
    *reinterpret_cast<uint4*>(p1 + kDwLBO + 16) = make_uint4(lo[12], lo[13], lo[14], lo[15]);
    *reinterpret_cast<uint4*>(p1 + kDwLBO) = make_uint4(lo[8], lo[9], lo[10], lo[11]);
  }
}

// The 4 weights this thread contributes to the B operand image of one layer (256 threads x 4 = 32 x 32):
//   forward  (B[n][k] = W[k][n]): thread (n = tid % 32, kq = tid / 32) gathers k = 4 kq .. 4 kq + 3;
//   backward (B[n][k] = W[n][k]): thread (kq = tid % 8, n = tid / 8) copies k = 4 kq .. 4 kq + 3.
template <bool BWD>
__device__ __forceinline__ void load_w2(const float* Wg, int tid, float (&w)[4]) {
  if (!BWD) {
    const int n = tid & 31, kq = tid >> 5;
#pragma unroll
    for (int i = 0; i < 4; ++i) w[i] = Wg[(4 * kq + i) * 32 + n];
  } else {
    const int kq = tid & 7, n = tid >> 3;
    const float4 v = *reinterpret_cast<const float4*>(Wg + n * 32 + 4 * kq);
    w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
  }
}

__device__ __forceinline__ void collect2(Ctx& c, float (&y)[16]) {
  mbar_wait(c.mbar, c.parity);
  c.parity ^= 1u;
  fence_after();
  uint32_t v[16];
  CLB_TMEM_LD16(c.row_addr + kColD + c.col, v);
  wait_ld();
#pragma unroll
  for (int k = 0; k < 16; ++k) y[k] = __uint_as_float(v[k]);
}

// Offset (in floats) of element (i = in, j = out) of a 32 x 32 kernel gradient inside one layer's slot of the FP32 partial.
// The slot is laid out so that every warp-level vector RED of collect_dw_red covers 256 contiguous bytes (16 lanes x 16 B):
// [j / 4][i][j % 4]: the 32 lanes of a warp (rows i = 0..31) write 512 contiguous bytes per RED.128.  (A row-major slot made
// each RED instruction touch 16 different 128-byte lines: 13 L1 requests per instruction, a quarter of all LSU wavefronts.)
__host__ __device__ inline int dw_slot32(int i, int j) {
  return ((((j >> 2) << 5) + i) << 2) + (j & 3);
}
// k_obs_tc2's dW product has M = 128 rows: [a_hi (32) | a_lo (32) | ONES (32) | unused (32)] -- an M = 64 instruction
// costs the tensor pipe exactly as much, and the row of ones turns the bias gradient (column sums of delta-p) into one more
// row of the same product instead of ~55 shuffle / select / add instructions per thread and layer.
constexpr int kPslotDet = 2 * 1024 + 4 * 32;   // deterministic layout of one layer's slot: [kernel from a_hi rows | from a_lo rows | bias [4 quarters][32]]
constexpr uint32_t kIdescDw128 = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((64u >> 3) << 17) | ((128u >> 4) << 24);

// dW rows live at lanes (r % 16) + 32 (r / 16): in every warp the lanes < 16 hold row r = 16 (warp % 4) + lane, i.e.
// feature i = r % 32 (rows 0..31 from a_hi, 32..63 from a_lo).  The thread folds the delta-hi and delta-lo column blocks of
// ITS 16 columns and adds them to wk[i][16 hf ..] with four 16-byte REDs (il_w > 0: an image layer's kernel, stored
// (out, in) with width il_w, scalar REDs).  When this returns the dW MMAs are complete: the operand images and the
// accumulator may be reused after the next __syncthreads().
#ifndef CLB_BIAS_COL
#define CLB_BIAS_COL 0      // 1: the dW product is taken TRANSPOSED, D = [dp_hi; dp_lo]^T [a_hi | a_lo | ONES (8 columns)] (M = 64, N = 72): column 64
                            // of the accumulator is the bias gradient (column sums of delta-p), so the 16 shuffles + ~45 ALU instructions per
                            // thread and layer of bias_red16 go away for 12.5 % more dW tensor time.  The ones sit on the N side (one extra
                            // 32-byte chunk per K row), not on the M side like CLB_BIAS_ONES (which doubled the A operand reads).
                            // Parity-green (all of tests/test_gpu_parity.py); measured SLOWER on B200: 17.28 vs 16.23 ms per 10 M observations.
#endif
#define CLB_TMEM_LD1(taddr, v) asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(taddr) : "memory")
constexpr uint32_t kIdescDwT72 = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((72u >> 3) << 17) | ((64u >> 4) << 24);
#ifndef CLB_BIAS_ONES
#define CLB_BIAS_ONES 0     // 1: bias gradient as a ones row of an M = 128 dW product (parity-green; measured SLOWER on B200, 17.2 vs
                            // 16.8 ms: the M = 128 product reads twice the A operand from shared memory and the dW MMAs get longer)
#endif
#if CLB_BIAS_ONES
constexpr uint32_t kIdescDwTc2 = kIdescDw128;
__device__ __forceinline__ void collect_dw_red(Ctx& c, float* wk, int il_w, float* bk = nullptr) {
  mbar_wait(c.mbar_dw, c.parity_dw);
  c.parity_dw ^= 1u;
  fence_after();
  const int q = (c.tid >> 5) & 3, lane = c.tid & 31;
  if (q == 3) return;                      // rows 96..127 of the product are not used
  const uint32_t addr = c.row_addr + kColDw + c.col;
  uint32_t v0[16], v1[16];
  CLB_TMEM_LD16(addr, v0);
  CLB_TMEM_LD16(addr + 32, v1);
  wait_ld();
  float f[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) f[k] = __uint_as_float(v0[k]) + __uint_as_float(v1[k]);     // . delta-p_hi  +  . delta-p_lo
  if (q < 2) {                             // lane = row i of a_hi^T dp (q = 0) / a_lo^T dp (q = 1): both add into dW[i][16 hf ..]
    if (wk == nullptr) return;
    if (il_w == 0) {
      float4* dst = reinterpret_cast<float4*>(wk) + (4 * c.hf) * 32 + lane;           // dw_slot32(lane, 16 hf + 4 qq) / 4
#pragma unroll
      for (int qq = 0; qq < 4; ++qq) atomicAdd(dst + qq * 32, make_float4(f[4 * qq], f[4 * qq + 1], f[4 * qq + 2], f[4 * qq + 3]));
    } else if (lane < il_w) {              // an image layer's kernel, stored (out, in) with width il_w
#pragma unroll
      for (int k = 0; k < 16; ++k) { const int j = 16 * c.hf + k; if (j < il_w) atomicAdd(&wk[j * il_w + lane], f[k]); }
    }
  } else if (lane == 0 && bk != nullptr) { // row 64 = ones^T dp: the bias gradient of my 16 columns
    if (il_w == 0) {
      float4* dst = reinterpret_cast<float4*>(bk + 16 * c.hf);
#pragma unroll
      for (int qq = 0; qq < 4; ++qq) atomicAdd(dst + qq, make_float4(f[4 * qq], f[4 * qq + 1], f[4 * qq + 2], f[4 * qq + 3]));
    } else {
#pragma unroll
      for (int k = 0; k < 16; ++k) { const int j = 16 * c.hf + k; if (j < il_w) atomicAdd(&bk[j], f[k]); }
    }
  }
}
#elif CLB_BIAS_COL
constexpr uint32_t kIdescDwTc2 = kIdescDwT72;
// Transposed product: D rows (lanes (r % 16) + 32 (r / 16)) are delta-p features j = r % 32 (rows 0..31 from dp_hi, 32..63 from dp_lo), columns
// 0..31 / 32..63 the a_hi / a_lo features i, column 64 the sum over the tile's observations.  Lane j of a warp's lower half adds its 16 values
// i = 16 hf .. to dW[i][j]: the partial's slot is stored TRANSPOSED ([i / 4][j][i % 4], dw_slot32(j, i)) so that these are again four 16-byte
// REDs per thread covering 256 contiguous bytes per warp.  bk = bias slot (or an image layer's bias gradient).
__device__ __forceinline__ void collect_dw_red(Ctx& c, float* wk, int il_w, float* bk = nullptr) {
  mbar_wait(c.mbar_dw, c.parity_dw);
  c.parity_dw ^= 1u;
  fence_after();
  const int q = (c.tid >> 5) & 3, lane = c.tid & 31;
  const uint32_t addr = c.row_addr + kColDw + c.col;
  uint32_t v0[16], v1[16], vb = 0u;
  CLB_TMEM_LD16(addr, v0);
  CLB_TMEM_LD16(addr + 32, v1);
  if (c.hf == 0) CLB_TMEM_LD1(c.row_addr + kColDw + 64u, vb);
  wait_ld();
  if (lane < 16) {
    const int j = (16 * q + lane) & 31;
    if (wk != nullptr) {
      float f[16];
#pragma unroll
      for (int k = 0; k < 16; ++k) f[k] = __uint_as_float(v0[k]) + __uint_as_float(v1[k]);
      if (il_w == 0) {
        float4* dst = reinterpret_cast<float4*>(wk + (q >= 2 ? c.lo_off : 0)) + (4 * c.hf) * 32 + j;      // dw_slot32(j, 16 hf + 4 qq) / 4
#pragma unroll
        for (int qq = 0; qq < 4; ++qq) atomicAdd(dst + qq * 32, make_float4(f[4 * qq], f[4 * qq + 1], f[4 * qq + 2], f[4 * qq + 3]));
      } else if (j < il_w) {               // an image layer's kernel, stored (out, in) with width il_w
#pragma unroll
        for (int k = 0; k < 16; ++k) { const int i = 16 * c.hf + k; if (i < il_w) atomicAdd(&wk[j * il_w + i], f[k]); }
      }
    }
    if (c.hf == 0 && bk != nullptr && (il_w == 0 || j < il_w)) atomicAdd(&bk[j], __uint_as_float(vb));
  }
}
#else
constexpr uint32_t kIdescDwTc2 = kIdescDw;
#ifndef CLB_DW_FOLD
#define CLB_DW_FOLD 0       // 1: the delta-p_hi and delta-p_lo column blocks of the dW product accumulate into the SAME 32 tensor-memory
                            // columns (two N = 32 instruction groups instead of one N = 64 group), so the collection reads 16 columns
                            // per thread and adds nothing (-17 instructions per thread and layer).  Parity-green; measured SLOWER on B200
                            // (16.94 vs 16.41 ms per 10 M observations): 32 MMAs to issue instead of 16 and no tensor time saved
#endif
constexpr uint32_t kIdescDwN32 = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((32u >> 3) << 17) | ((64u >> 4) << 24);
// M = 64 product: D rows live at lanes (r % 16) + 32 (r / 16): in every warp the lanes < 16 hold row r = 16 (warp % 4) + lane,
// i.e. feature i = r % 32 (rows 0..31 from a_hi, 32..63 from a_lo, both added into dW[i][16 hf ..]).
__device__ __forceinline__ void collect_dw_red(Ctx& c, float* wk, int il_w, float* bk = nullptr) {
  mbar_wait(c.mbar_dw, c.parity_dw);
  c.parity_dw ^= 1u;
  fence_after();
  const int q = (c.tid >> 5) & 3, lane = c.tid & 31;
  const uint32_t addr = c.row_addr + kColDw + c.col;
#if CLB_DW_FOLD
  uint32_t v0[16];
  CLB_TMEM_LD16(addr, v0);
  wait_ld();
#elif CLB_DWB_PACK
  uint32_t v0[16], v1[16];
  {
    uint32_t v[32];
    CLB_TMEM_LD32(c.row_addr + kColDw + 32u * (uint32_t)c.hf, v);
    wait_ld();
#pragma unroll
    for (int k = 0; k < 16; ++k) { v0[k] = v[k]; v1[k] = v[16 + k]; }
  }
#else
  uint32_t v0[16], v1[16];
  CLB_TMEM_LD16(addr, v0);
  CLB_TMEM_LD16(addr + 32, v1);
  wait_ld();
#endif
  if (lane < 16 && wk != nullptr) {
    const int i = (16 * q + lane) & 31;
    float f[16];
#pragma unroll
#if CLB_DW_FOLD
    for (int k = 0; k < 16; ++k) f[k] = __uint_as_float(v0[k]);
#else
    for (int k = 0; k < 16; ++k) f[k] = __uint_as_float(v0[k]) + __uint_as_float(v1[k]);
#endif
    if (il_w == 0) {
      float4* dst = reinterpret_cast<float4*>(wk + (q >= 2 ? c.lo_off : 0)) + (4 * c.hf) * 32 + i;      // dw_slot32(i, 16 hf + 4 qq) / 4
#pragma unroll
      for (int qq = 0; qq < 4; ++qq) atomicAdd(dst + qq * 32, make_float4(f[4 * qq], f[4 * qq + 1], f[4 * qq + 2], f[4 * qq + 3]));
    } else if (i < il_w) {
#pragma unroll
      for (int k = 0; k < 16; ++k) { const int j = 16 * c.hf + k; if (j < il_w) atomicAdd(&wk[j * il_w + i], f[k]); }
    }
  }
}
#endif

// The tcgen05.mma's of one tile's dW = [a_hi; a_lo]^T [delta-p_hi | delta-p_lo] (one elected lane; operands warp-uniform).
__device__ __forceinline__ void issue_dw_mmas(uint32_t d, uint64_t a0, uint64_t b0) {
#if CLB_BIAS_COL && !CLB_BIAS_ONES
  // transposed: A operand = the delta-p images, B operand = [a_hi | a_lo | ones] (the ones block follows the a images)
#pragma unroll
  for (int ks = 0; ks < kThreads / 8; ++ks)
    mma_tf32_ss(d, b0 + (uint64_t)((2u * kDwSBO * (uint32_t)ks) >> 4), a0 + (uint64_t)((2u * kDwSBO * (uint32_t)ks) >> 4), kIdescDwT72, ks > 0 ? 1u : 0u);
#elif CLB_DW_FOLD && !CLB_BIAS_ONES
  // x delta-p_hi, then x delta-p_lo (MN group 1: + kDwLBO bytes) into the same 32 accumulator columns
#pragma unroll
  for (int ks = 0; ks < kThreads / 8; ++ks)
    mma_tf32_ss(d, a0 + (uint64_t)((2u * kDwSBO * (uint32_t)ks) >> 4), b0 + (uint64_t)((2u * kDwSBO * (uint32_t)ks) >> 4), kIdescDwN32, ks > 0 ? 1u : 0u);
  const uint64_t b1 = b0 + (uint64_t)(kDwLBO >> 4);
#pragma unroll
  for (int ks = 0; ks < kThreads / 8; ++ks)
    mma_tf32_ss(d, a0 + (uint64_t)((2u * kDwSBO * (uint32_t)ks) >> 4), b1 + (uint64_t)((2u * kDwSBO * (uint32_t)ks) >> 4), kIdescDwN32, 1u);
#else
#pragma unroll
  for (int ks = 0; ks < kThreads / 8; ++ks)
    mma_tf32_ss(d, a0 + (uint64_t)((2u * kDwSBO * (uint32_t)ks) >> 4), b0 + (uint64_t)((2u * kDwSBO * (uint32_t)ks) >> 4), kIdescDwTc2, ks > 0 ? 1u : 0u);
#endif
}

// ---- weight images by TMA --------------------------------------------------------------------------------------
// The hi / lo B-operand images of every hidden layer (both orientations) are prepared once per step in global memory
// in their final shared-memory byte layout (k_pack_images); a pass copies its two images (4224 B each) with
// cp.async.bulk into the buffer (pass & 1) while the previous pass is still computing.  No thread builds images, no
// register staging, no generic-proxy stores: the MMA issuer alone waits for the copy.
__device__ __forceinline__ void tma_fetch_image(uint32_t dst_smem, const float* src_hi_lo, uint32_t bar) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(2u * kImgBytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(dst_smem), "l"(src_hi_lo), "r"(2u * kImgBytes), "r"(bar) : "memory");
}

// Issue one chain pass from buffer b; `tma` = the image came by TMA (wait for it), `next` = global image of the next
// pass to prefetch into the other buffer (or null).  Called by all threads after the pass's __syncthreads().
#ifndef CLB_ST32
#define CLB_ST32 1          // 1: the chain's A operand (hi and lo parts of a thread's 16 features) is written with ONE tcgen05.st.x32 into 32
                            // adjacent columns [hi 16 | lo 16] per feature half instead of two .x16 stores into separate hi / lo regions
#endif
// first tensor-memory column (relative to the CTA base) of K-slice ks (features 8 ks .. 8 ks + 7) of the hi / lo operand
__device__ __forceinline__ constexpr uint32_t acol_hi(int ks) { return CLB_ST32 ? (uint32_t)(32 * (ks >> 1) + 8 * (ks & 1)) : kColAhi + 8u * (uint32_t)ks; }
__device__ __forceinline__ constexpr uint32_t acol_lo(int ks) { return CLB_ST32 ? (uint32_t)(32 * (ks >> 1) + 16 + 8 * (ks & 1)) : kColAlo + 8u * (uint32_t)ks; }
// store this thread's hi / lo operand halves (two threads per row)
__device__ __forceinline__ void store_operand16(uint32_t row_addr, int hf, const uint32_t (&hi)[16], const uint32_t (&lo)[16]) {
#if CLB_ST32
  uint32_t v[32];
#pragma unroll
  for (int k = 0; k < 16; ++k) { v[k] = hi[k]; v[16 + k] = lo[k]; }
  CLB_TMEM_ST32(row_addr + 32u * (uint32_t)hf, v);
#else
  CLB_TMEM_ST16(row_addr + kColAhi + 16u * (uint32_t)hf, hi);
  CLB_TMEM_ST16(row_addr + kColAlo + 16u * (uint32_t)hf, lo);
#endif
}
#ifndef CLB_BIAS_IN_MMA
#define CLB_BIAS_IN_MMA 0   // 1: a forward pass starts from an accumulator PRELOADED with the layer's bias (one tcgen05.st per thread
                            // and every MMA accumulating) instead of 16 FADDs per thread after the collection.  Parity-green; measured
                            // SLOWER on B200 (16.89 vs 16.41 ms): the third tcgen05.st per pass costs more than the 16 FADDs it saves
#endif
__device__ __forceinline__ void issue_chain_mmas2(Ctx& c, bool tma, const float* next, bool preloaded = false) {
  const uint32_t b = c.pass & 1u;
  const uint32_t warp = uniform32((uint32_t)c.tid >> 5);
  if (warp == 0u) {
    fence_after();
    const uint32_t base = uniform32(c.base);
    const uint64_t bhi = uniform64(b ? c.wdesc_hi[1] : c.wdesc_hi[0]), blo = uniform64(b ? c.wdesc_lo[1] : c.wdesc_lo[0]);
    const uint32_t d = base + kColD;
    const uint32_t bar = uniform32(c.mbar);
    const uint32_t wb = uniform32(b ? c.wbar[1] : c.wbar[0]), wph = uniform32((c.wphase >> b) & 1u);
    const uint32_t nb = uniform32(b ? c.wbar[0] : c.wbar[1]), ndst = uniform32(b ? c.wimg[0] : c.wimg[1]);
    if (elect_one()) {
      if (tma) mbar_wait(wb, wph);
#ifndef CLB_ABL_CHAIN
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
        mma_tf32_ts(d, base + acol_hi(ks), blo + (uint64_t)((2u * kLBO * (uint32_t)ks) >> 4), (ks > 0 || preloaded) ? 1u : 0u);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
        mma_tf32_ts(d, base + acol_lo(ks), bhi + (uint64_t)((2u * kLBO * (uint32_t)ks) >> 4), 1u);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
        mma_tf32_ts(d, base + acol_hi(ks), bhi + (uint64_t)((2u * kLBO * (uint32_t)ks) >> 4), 1u);
#endif
      commit(bar);
      if (next != nullptr) tma_fetch_image(ndst, next, nb);
    }
    __syncwarp();
  }
  if (tma) c.wphase ^= (1u << b);
  c.pass += 1u;
}

// Build the images of this pass in buffer (pass & 1) from FP32 weights `Wg` (image layers: the tile's kernels in
// shared memory).  All threads; the caller's fence + __syncthreads() publish them.
template <bool BWD>
__device__ __forceinline__ void build_weight_image3(Ctx& c, const float* Wg, char* img_base) {
  float w[4];
  load_w2<BWD>(Wg, c.tid, w);
  int n, kq;
  if (!BWD) { n = c.tid & 31; kq = c.tid >> 5; }
  else { kq = c.tid & 7; n = c.tid >> 3; }
  const uint32_t off = kq * kLBO + (n >> 3) * kSBO + (n & 7) * 16;
  float4 hi, lo;
  hi.x = tf32_rna(w[0]); hi.y = tf32_rna(w[1]); hi.z = tf32_rna(w[2]); hi.w = tf32_rna(w[3]);
  lo.x = w[0] - hi.x; lo.y = w[1] - hi.y; lo.z = w[2] - hi.z; lo.w = w[3] - hi.w;
  char* dst = img_base + (size_t)(c.pass & 1u) * 2 * kImgBytes;
  *reinterpret_cast<float4*>(dst + off) = hi;
  *reinterpret_cast<float4*>(dst + kImgBytes + off) = lo;
}

// Forward pass of one layer.  build_from == nullptr: the layer's images were prefetched by TMA.
// bias16 (CLB_BIAS_IN_MMA): this thread's 16 biases of the layer (shared memory, 16-byte aligned); they are stored into the thread's
// accumulator columns and every MMA of the pass accumulates on top.
__device__ __forceinline__ void issue3(Ctx& c, const float (&x)[16], const float* build_from, char* img_base, const float* next,
                                       const float* bias16 = nullptr) {
  {
    uint32_t hi[16], lo[16];
    split16(x, hi, lo);
    store_operand16(c.row_addr, c.hf, hi, lo);
  }
  if (bias16 != nullptr) {
    uint32_t b[16];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const uint4 v = *reinterpret_cast<const uint4*>(bias16 + 4 * q);
      b[4 * q] = v.x; b[4 * q + 1] = v.y; b[4 * q + 2] = v.z; b[4 * q + 3] = v.w;
    }
    CLB_TMEM_ST16(c.row_addr + kColD + c.col, b);
  }
  if (build_from != nullptr) { build_weight_image3<false>(c, build_from, img_base); fence_async_smem(); }
  wait_st();
  fence_before();
  __syncthreads();
  issue_chain_mmas2(c, build_from == nullptr, next, bias16 != nullptr);
}

// Backward of one layer: dX chain (if need_dx) and dW.
__device__ __forceinline__ void issue_backward3(Ctx& c, const float (&dp)[16], const float (&ain)[16], bool need_dx,
                                                const float* build_from, char* img_base, const float* next) {
  {
    uint32_t hi[16], lo[16];
    split16(dp, hi, lo);
    if (need_dx) store_operand16(c.row_addr, c.hf, hi, lo);
#if CLB_STS_DIVERGE
    dw_store_half_div(c.dw_b, c.row, c.hf, hi, lo);
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      hi[k] = __float_as_uint(ain[k]);
      lo[k] = __float_as_uint(ain[k] - __uint_as_float(hi[k] & 0xFFFFE000u));
    }
    dw_store_half_div(c.dw_a, c.row, c.hf, hi, lo);
#else
    const int sw = (c.row >> 2) & 1;       // conflict-free image stores: see dw_store_half
    swap_blocks(hi, sw); swap_blocks(lo, sw);
    dw_store_half<CLB_DWB_PACK != 0>(c.dw_b, c.row, c.hf, hi, lo, sw);
    {
      uint32_t a2[16];
#pragma unroll
      for (int k = 0; k < 16; ++k) a2[k] = __float_as_uint(ain[k]);
      swap_blocks(a2, sw);
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        hi[k] = a2[k];
        lo[k] = __float_as_uint(__uint_as_float(a2[k]) - __uint_as_float(a2[k] & 0xFFFFE000u));
      }
    }
    dw_store_half(c.dw_a, c.row, c.hf, hi, lo, sw);
#endif
  }
  if (need_dx && build_from != nullptr) build_weight_image3<true>(c, build_from, img_base);
  wait_st();
  fence_async_smem();
  fence_before();
  __syncthreads();
  if (need_dx) issue_chain_mmas2(c, build_from == nullptr, next);
  const uint32_t warp = uniform32((uint32_t)c.tid >> 5);
#ifndef CLB_DW_ISSUER
#define CLB_DW_ISSUER 7     // which warp issues the dW product: 7 = concurrently with warp 0's chain (default), 0 = warp 0, after its chain
#endif
  if (warp == (uint32_t)CLB_DW_ISSUER) {
    fence_after();
    const uint32_t d = uniform32(c.base) + kColDw;
    const uint64_t a0 = uniform64(c.desc_dwa), b0 = uniform64(c.desc_dwb);
    const uint32_t bar = uniform32(c.mbar_dw);
    if (elect_one()) {
#ifndef CLB_ABL_DW
      issue_dw_mmas(d, a0, b0);
#endif
      commit(bar);
    }
    __syncwarp();
  }
}


// ---- barrier-free hand-over (k_obs_tc2, round 2) -----------------------------------------------------------------
// A pass used to be: all warps store operands -> __syncthreads() -> warp 0 issues -> everybody waits.  ncu (round 1): 13 % of
// the stall samples sat at that barrier, and every warp paid for the slowest one twice (at the barrier and at the mbarrier).
// Now each warp ARRIVES on a shared-memory counter (one acq_rel atomic per warp) and goes on; the warp that arrives LAST
// finds all operands in place and issues the tcgen05.mma's itself.  No warp ever waits for another warp -- only for the
// tensor pipe.  The backward step of a layer hands over the critical chain (delta-a = delta-p W^T) first, builds the dW operand
// images while it runs, and collects the dW accumulator one layer later (after the next chain has been handed over).
__device__ __forceinline__ bool arrive_last(uint32_t cnt_smem, int lane, uint32_t last) {
  __syncwarp();
  uint32_t old = 0;
  if (lane == 0) asm volatile("atom.acq_rel.cta.shared.inc.u32 %0, [%1], %2;" : "=r"(old) : "r"(cnt_smem), "r"(last) : "memory");   // wraps to 0
  old = __shfl_sync(0xffffffffu, old, 0);
  return old == last;
}

// The 12 tcgen05.mma of one chain pass from weight buffer (pass & 1); all 32 lanes of the issuing warp call it.
__device__ __forceinline__ void chain_issue(Ctx& c, bool tma, const float* next) {
  const uint32_t b = c.pass & 1u;
  fence_after();
  const uint32_t base = uniform32(c.base);
  const uint64_t bhi = uniform64(b ? c.wdesc_hi[1] : c.wdesc_hi[0]), blo = uniform64(b ? c.wdesc_lo[1] : c.wdesc_lo[0]);
  const uint32_t d = base + kColD;
  const uint32_t bar = uniform32(c.mbar);
  const uint32_t wb = uniform32(b ? c.wbar[1] : c.wbar[0]), wph = uniform32((c.wphase >> b) & 1u);
  const uint32_t nb = uniform32(b ? c.wbar[0] : c.wbar[1]), ndst = uniform32(b ? c.wimg[0] : c.wimg[1]);
  if (elect_one()) {
    if (tma) mbar_wait(wb, wph);
#pragma unroll
    for (int ks = 0; ks < 4; ++ks)
      mma_tf32_ts(d, base + kColAhi + 8u * (uint32_t)ks, blo + (uint64_t)((2u * kLBO * (uint32_t)ks) >> 4), ks > 0 ? 1u : 0u);
#pragma unroll
    for (int ks = 0; ks < 4; ++ks)
      mma_tf32_ts(d, base + kColAlo + 8u * (uint32_t)ks, bhi + (uint64_t)((2u * kLBO * (uint32_t)ks) >> 4), 1u);
#pragma unroll
    for (int ks = 0; ks < 4; ++ks)
      mma_tf32_ts(d, base + kColAhi + 8u * (uint32_t)ks, bhi + (uint64_t)((2u * kLBO * (uint32_t)ks) >> 4), 1u);
    commit(bar);
    if (next != nullptr) tma_fetch_image(ndst, next, nb);
  }
  __syncwarp();
}
__device__ __forceinline__ void pass_advance(Ctx& c, bool tma) {
  if (tma) c.wphase ^= (1u << (c.pass & 1u));
  c.pass += 1u;
}

// Hand one chain pass over: operands hi / lo of my half row go to tensor memory, (image layers: my piece of the weight image
// to shared memory), then arrive; the last warp issues.
template <bool BWD>
__device__ __forceinline__ void chain_handover(Ctx& c, const uint32_t (&hi)[16], const uint32_t (&lo)[16], const float* build_from,
                                               char* img_base, const float* next, int lane) {
  CLB_TMEM_ST16(c.row_addr + kColAhi + c.col, hi);
  CLB_TMEM_ST16(c.row_addr + kColAlo + c.col, lo);
  if (build_from != nullptr) { build_weight_image3<BWD>(c, build_from, img_base); fence_async_smem(); }
  wait_st();
  fence_before();
  if (arrive_last(c.cnt_chain, lane, c.n_warps_m1)) chain_issue(c, build_from == nullptr, next);
  pass_advance(c, build_from == nullptr);
}

__device__ __forceinline__ void issue4(Ctx& c, const float (&x)[16], const float* build_from, char* img_base, const float* next, int lane) {
  uint32_t hi[16], lo[16];
  split16(x, hi, lo);
  chain_handover<false>(c, hi, lo, build_from, img_base, next, lane);
}

__device__ __forceinline__ void dw_issue(Ctx& c, const float4* dead, int lane) {     // all 32 lanes of the issuing warp
  fence_after();
  const uint32_t d = uniform32(c.base) + kColDw;
  const uint64_t a0 = uniform64(c.desc_dwa), b0 = uniform64(c.desc_dwb);
  const uint32_t bar = uniform32(c.mbar_dw);
  if (elect_one()) {
    issue_dw_mmas(d, a0, b0);
    commit(bar);
  }
  __syncwarp();
  if (dead != nullptr) {      // every warp has consumed the layer's input activations: drop their 128 scratch lines from the L2
#pragma unroll
    for (int i = 0; i < 4; ++i) asm volatile("discard.global.L2 [%0], 128;" :: "l"(reinterpret_cast<const char*>(dead) + (size_t)(32 * i + lane) * 128) : "memory");
  }
}

// One hand-over per backward layer: chain operands (if need_dx) AND the dW operand images, one arrival; the last warp issues
// the chain first, then dW.
__device__ __forceinline__ void bwd_handover(Ctx& c, uint32_t (&hi)[16], uint32_t (&lo)[16], const float (&ain)[16], bool need_dx,
                                             const float* build_from, char* img_base, const float* next, const float4* dead, int lane) {
  if (need_dx) {
    CLB_TMEM_ST16(c.row_addr + kColAhi + c.col, hi);
    CLB_TMEM_ST16(c.row_addr + kColAlo + c.col, lo);
    if (build_from != nullptr) build_weight_image3<true>(c, build_from, img_base);
  }
  const int sw = (c.row >> 2) & 1;       // conflict-free image stores: see dw_store_half
  swap_blocks(hi, sw); swap_blocks(lo, sw);
  dw_store_half<CLB_DWB_PACK != 0>(c.dw_b, c.row, c.hf, hi, lo, sw);
  {
    uint32_t a2[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) a2[k] = __float_as_uint(ain[k]);
    swap_blocks(a2, sw);
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      hi[k] = a2[k];
      lo[k] = __float_as_uint(__uint_as_float(a2[k]) - __uint_as_float(a2[k] & 0xFFFFE000u));
    }
  }
  dw_store_half(c.dw_a, c.row, c.hf, hi, lo, sw);
  wait_st();
  fence_async_smem();
  fence_before();
  if (arrive_last(c.cnt_dw, lane, c.n_warps_m1)) {
    if (need_dx) chain_issue(c, build_from == nullptr, next);
    dw_issue(c, dead, lane);
  }
  if (need_dx) pass_advance(c, build_from == nullptr);
}

// dW operand images of this layer (delta-p split in hi / lo, the layer input `ain`), then arrive; the last warp issues the 16
// tcgen05.mma of dW = a^T delta-p and, the layer's input activations now being consumed by every warp, drops their 128
// scratch lines (`dead`: the 16 KB slot, or null) from the L2.
__device__ __forceinline__ void dw_handover(Ctx& c, uint32_t (&hi)[16], uint32_t (&lo)[16], const float (&ain)[16], const float4* dead, int lane) {
  const int sw = (c.row >> 2) & 1;       // conflict-free image stores: see dw_store_half
  swap_blocks(hi, sw); swap_blocks(lo, sw);
  dw_store_half<CLB_DWB_PACK != 0>(c.dw_b, c.row, c.hf, hi, lo, sw);
  {
    uint32_t a2[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) a2[k] = __float_as_uint(ain[k]);
    swap_blocks(a2, sw);
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      hi[k] = a2[k];
      lo[k] = __float_as_uint(__uint_as_float(a2[k]) - __uint_as_float(a2[k] & 0xFFFFE000u));
    }
  }
  dw_store_half(c.dw_a, c.row, c.hf, hi, lo, sw);
  fence_async_smem();
  fence_before();
  if (arrive_last(c.cnt_dw, lane, c.n_warps_m1)) {
    fence_after();
    const uint32_t d = uniform32(c.base) + kColDw;
    const uint64_t a0 = uniform64(c.desc_dwa), b0 = uniform64(c.desc_dwb);
    const uint32_t bar = uniform32(c.mbar_dw);
    if (elect_one()) {
      issue_dw_mmas(d, a0, b0);
      commit(bar);
    }
    __syncwarp();
    if (dead != nullptr) {
#pragma unroll
      for (int i = 0; i < 4; ++i) asm volatile("discard.global.L2 [%0], 128;" :: "l"(reinterpret_cast<const char*>(dead) + (size_t)(32 * i + lane) * 128) : "memory");
    }
  }
}

}  // namespace tc
}  // namespace clb
