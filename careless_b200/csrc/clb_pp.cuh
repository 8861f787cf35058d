// clb_pp.cuh -- k_obs_pp: the width-32 observation kernel as a warp-specialised, two-tile ping-pong (sm_100a).
//
// Same arithmetic, operand images, tensor-memory products and global layouts as k_obs_tc2 (clb_kernels.cuh / clb_tc.cuh) --
// what changes is WHO waits for WHAT.  k_obs_tc2 runs one 128-row tile per CTA: all 8 warps prepare operands, meet at a
// __syncthreads(), one of them issues the tcgen05.mma's and then everybody sits at an mbarrier until the tensor pipe
// answers (ncu, round 1: 39 % of the stall samples are that wait, 7 % the barrier; issue slots 40 % busy, tensor pipe 28 %).
// Here one CTA per SM owns TWO tiles (A, B) and 9 warps:
//   * warps 0..7 ("workers", two threads per row as before) alternate between the tiles at pass granularity: they hand
//     tile A's operands to the tensor pipe and, instead of waiting, do tile B's share of the same layer; by the time they
//     come back to A its products have long finished.  Hand-over is an mbarrier arrive (one per warp), never a CTA barrier;
//   * warp 8 ("issuer") does nothing but wait for operands and issue: chain(A), dW(A), chain(B), dW(B) per layer, in the
//     order the workers produce them, so the in-order tensor pipe never holds a critical product behind a late one; it also
//     streams the layer's weight images in by TMA, one fetch serving both tiles (they move through the layers in lockstep).
//   * the backward step of a layer starts with the critical chain (delta-a = delta-p W^T) and only then builds the dW operand
//     images; the dW accumulator of layer k is collected one step later, after layer k-1's chain has been handed over.
// Tensor memory: 512 columns (one CTA per SM), 256 per tile: A_hi 0, A_lo 32, D 64, dW accumulator 128 (64 columns).
// Shared memory: per tile the two MN-major dW operand images (64 KB), one double-buffered weight image pair for both tiles.
// Included by clb_kernels.cuh (needs ObsArgs, obs_epilogue, bias_red16, discard_line).
#pragma once

namespace clb {
namespace pp {

using namespace tc;

constexpr int kWorkers = 256;                 // worker threads (8 warps)
constexpr int kThreadsPP = 288;               // + the issuer warp (ptxas budgets 168 registers per thread: 9 warps round up to 12)
constexpr int TR = 128;                       // rows per tile
constexpr uint32_t kTileCols = 256;           // tensor-memory columns per tile
constexpr uint32_t kTmemColsPP = 512;
constexpr uint32_t cAhi = 0, cAlo = 32, cD = 64, cDw = 128;
constexpr int PSLOT = 32 * 32 + 32;
// mbarriers: [x] = tile
enum { B_OPND_CHAIN = 0, B_RES_CHAIN = 2, B_OPND_DW = 4, B_DW_DONE = 6, B_WIMG = 8, N_BARS = 10 };

struct SmemPP {
  static size_t bytes(int n_layers) {
    return 2 * 2 * (size_t)kDwImgBytes + 4 * (size_t)kImgBytes + sizeof(float) * (64 + (size_t)n_layers * 32)
           + 4 * TR * sizeof(float2) + 64 * sizeof(double) + N_BARS * sizeof(uint64_t) + 64 + 1024;
  }
};

__device__ __forceinline__ void tmem_alloc512(uint32_t slot_smem) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(slot_smem), "r"(kTmemColsPP) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc512(uint32_t tbase) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tbase), "r"(kTmemColsPP) : "memory");
}
// one arrival per warp: every lane has fenced its own writes, __syncwarp orders them before lane 0's (releasing) arrive
__device__ __forceinline__ void warp_arrive(uint32_t bar, int lane) {
  __syncwarp();
  if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void workers_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// per-thread state of one of the two tiles
// (everything that is a fixed offset from a per-CTA base -- tensor-memory columns, barriers, operand images, scratch -- is
// derived from the compile-time tile index at the point of use instead of living in registers)
struct Tile {
  int refl; bool inb, active;
  float x[16];                 // forward: my half of the activation vector; afterwards the layer input a_LT for the head
  float y[16];                 // backward: prefetched input activations of the next layer down
  float dmu, drho;
  unsigned mask;
};
template <int X> struct TileIdx { static constexpr int value = X; };

}  // namespace pp

template <int LIK>
__global__ void __launch_bounds__(pp::kThreadsPP, 1) k_obs_pp(ObsArgs a) {
  using namespace pp;
  constexpr int WP = 32, NC = 8, HW = 16;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const int NL = a.lay.n_layers, L = NL - 1, LT = L;           // no image layers in this kernel
  unsigned char* sp = smem_raw;
  char* img_base = reinterpret_cast<char*>(sp);                // [tile][a_hi a_lo | dp_hi dp_lo]
  sp += 4 * (size_t)kDwImgBytes;
  char* w_img = reinterpret_cast<char*>(sp);                   // [2 buffers][hi, lo][kImgBytes]
  sp += 4 * (size_t)kImgBytes;
  float* Whead = reinterpret_cast<float*>(sp);                 // [32][2]
  float* bsm = Whead + 64;                                     // [NL][32]
  float2* xch = reinterpret_cast<float2*>(bsm + (size_t)NL * WP);   // [0..1][TR]: head partial sums for tile A / B; [2..3][TR]: (dmu, drho)
  double* red = reinterpret_cast<double*>(xch + 4 * TR);
  uint64_t* bars = reinterpret_cast<uint64_t*>(red + 64);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bars + N_BARS);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool is_issuer = warp == 8;
  if (tid == 0) {
    for (int x = 0; x < 2; ++x) {
      mbar_init(smem_u32(bars + B_OPND_CHAIN + x), 8); mbar_init(smem_u32(bars + B_OPND_DW + x), 8);
      mbar_init(smem_u32(bars + B_RES_CHAIN + x), 1); mbar_init(smem_u32(bars + B_DW_DONE + x), 1);
      mbar_init(smem_u32(bars + B_WIMG + x), 1);
    }
  }
  if (is_issuer) tmem_alloc512(smem_u32(slot));
  fence_before();
  for (int idx = tid; idx < WP * 2; idx += kThreadsPP) {
    const int i = idx / 2, j = idx % 2;
    Whead[idx] = (i < a.lay.in_dim[L] && j < a.lay.out_dim[L]) ? a.theta_mlp[a.lay.koff[L] + i * a.lay.out_dim[L] + j] : 0.f;
  }
  for (int idx = tid; idx < NL * WP; idx += kThreadsPP) {
    const int k = idx / WP, j = idx % WP;
    bsm[idx] = (j < a.lay.out_dim[k]) ? a.theta_mlp[a.lay.boff[k] + j] : 0.f;
  }
  __syncthreads();
  fence_after();
  const uint32_t tbase = *slot;
  constexpr size_t IMGF = kImgBytes / 4;
  auto gimg = [&](int k, int dir) -> const float* { return a.wimg + ((size_t)(k * 2 + dir) * 2) * IMGF; };
  const int64_t n_tiles = (a.n_rows + TR - 1) / TR;
  const int64_t n_pairs = (n_tiles + 1) / 2;
  const bool train = a.train_mlp != 0;

  if (is_issuer) {
    // =========================================== issuer warp ===========================================
    const uint32_t wbuf0 = smem_u32(w_img), wbuf1 = smem_u32(w_img + 2 * kImgBytes);
    const uint32_t wbar0 = smem_u32(bars + B_WIMG), wbar1 = smem_u32(bars + B_WIMG + 1);
    uint32_t pass = 0;                                         // weight-image passes so far: buffer = pass & 1, phase = (pass >> 1) & 1
    uint32_t par_oc = 0, par_od = 0;                           // bit x: phase of tile x's operand barriers
    const uint32_t tb0 = uniform32(tbase);                     // warp-uniform operands: no R2UR waterfall per MMA (see clb_tc.cuh)
    if (blockIdx.x < n_pairs && elect_one()) tma_fetch_image(wbuf0, gimg(0, 0), wbar0);
    __syncwarp();
    // one chain pass of tile x from weight buffer b: 12 MMAs (X_hi W_lo + X_lo W_hi + X_hi W_hi) into the tile's D
    auto chain = [&](int x, uint32_t b) {
      mbar_wait(smem_u32(bars + B_OPND_CHAIN + x), (par_oc >> x) & 1u); par_oc ^= (1u << x);
      fence_after();
      const uint32_t tb = tb0 + (uint32_t)x * kTileCols;
      const uint32_t wb = b ? wbuf1 : wbuf0;
      const uint64_t bhi = uniform64(make_desc(wb)), blo = uniform64(make_desc(wb + kImgBytes));
      const uint32_t d = tb + cD;
      const uint32_t rbar = uniform32(smem_u32(bars + B_RES_CHAIN + x));
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) mma_tf32_ts(d, tb + cAhi + 8u * (uint32_t)ks, blo + (uint64_t)((2u * kLBO * (uint32_t)ks) >> 4), ks > 0 ? 1u : 0u);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) mma_tf32_ts(d, tb + cAlo + 8u * (uint32_t)ks, bhi + (uint64_t)((2u * kLBO * (uint32_t)ks) >> 4), 1u);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) mma_tf32_ts(d, tb + cAhi + 8u * (uint32_t)ks, bhi + (uint64_t)((2u * kLBO * (uint32_t)ks) >> 4), 1u);
        commit(rbar);
      }
      __syncwarp();
    };
    auto dw = [&](int x) {
      mbar_wait(smem_u32(bars + B_OPND_DW + x), (par_od >> x) & 1u); par_od ^= (1u << x);
      fence_after();
      const uint32_t d = tb0 + (uint32_t)x * kTileCols + cDw;
      const uint64_t a0 = uniform64(make_desc_mn(smem_u32(img_base + (size_t)x * 2 * kDwImgBytes)));
      const uint64_t b0 = uniform64(make_desc_mn(smem_u32(img_base + (size_t)x * 2 * kDwImgBytes + kDwImgBytes)));
      const uint32_t dbar = uniform32(smem_u32(bars + B_DW_DONE + x));
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < TR / 8; ++ks)
          mma_tf32_ss(d, a0 + (uint64_t)((2u * kDwSBO * (uint32_t)ks) >> 4), b0 + (uint64_t)((2u * kDwSBO * (uint32_t)ks) >> 4), kIdescDw, ks > 0 ? 1u : 0u);
        commit(dbar);
      }
      __syncwarp();
    };
    // wait for this pass's weight images; returns the buffer
    auto wait_w = [&]() -> uint32_t {
      const uint32_t b = pass & 1u;
      mbar_wait(b ? wbar1 : wbar0, (pass >> 1) & 1u);
      return b;
    };
    // after both tiles' chains of this pass have been issued their predecessors have been consumed: the other buffer is free
    auto prefetch = [&](const float* next) {
      pass += 1u;
      const uint32_t dst = uniform32((pass & 1u) ? wbuf1 : wbuf0), nb = uniform32((pass & 1u) ? wbar1 : wbar0);
      if (next != nullptr && elect_one()) tma_fetch_image(dst, next, nb);
      __syncwarp();
    };
    for (int64_t pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
      const bool more = pair + gridDim.x < n_pairs;
      for (int k = 0; k < LT; ++k) {
        const uint32_t b = wait_w();
        chain(0, b); chain(1, b);
        prefetch((k + 1 < LT) ? gimg(k + 1, 0) : (train && LT > 1) ? gimg(LT - 1, 1) : (more ? gimg(0, 0) : nullptr));
      }
      if (!train) continue;
      dw(0); dw(1);                                            // head: dW_out = a_L^T [dmu, drho]
      for (int k = LT - 1; k >= 0; --k) {
        if (k > 0) {
          const uint32_t b = wait_w();
          chain(0, b); dw(0); chain(1, b); dw(1);
          prefetch((k > 1) ? gimg(k - 1, 1) : (more ? gimg(0, 0) : nullptr));
        } else {
          dw(0); dw(1);
        }
      }
    }
    fence_before();
  } else {
    // =========================================== worker warps ===========================================
    const int rrow = tid & (TR - 1), hf = tid >> 7;
    const uint32_t col = (uint32_t)(HW * hf);
    Tile T[2];
    const uint32_t tm0 = tbase + ((uint32_t)(32 * (warp & 3)) << 16);      // my row in tile 0's column block (+ kTileCols for tile 1)
    const uint32_t bar0 = smem_u32(bars);
    float4* const scr0 = a.scratch + (size_t)blockIdx.x * 2 * (size_t)LT * NC * TR;     // [tile][LT][8][TR]
    const size_t scr_tile = (size_t)LT * NC * TR;
    uint32_t par = 0;                                          // phase bits: bit x = res_chain of tile x, bit 2 + x = dw_done of tile x
    float* part32 = a.partials32 + (size_t)(blockIdx.x % a.n_partials) * NL * PSLOT;
    double ll_sum = 0.0;
    float ev_f = 1.f, ev_a = 0.f, ev_b = 0.f;
    if (a.theta_lik != nullptr) { ev_f = softplusf(a.theta_lik[0]); ev_a = softplusf(a.theta_lik[1]); ev_b = softplusf(a.theta_lik[2]); }
    const int sw = (rrow >> 2) & 1;                            // conflict-free image stores: see tc::dw_store_half
    int64_t pair = blockIdx.x;
    auto tile_row = [&](int x) -> int64_t { return (2 * pair + x) * TR + rrow; };
    auto bar = [&](int which, int x) -> uint32_t { return bar0 + 8u * (uint32_t)(which + x); };

    // ---- forward: consume the products of pass k-1 (k > 0), hand the operands of pass k (k < LT) to the issuer ----
    auto fwd_step = [&](auto X, int k) {
      constexpr int x = decltype(X)::value;
      Tile& t = T[x];
      const uint32_t tm = tm0 + (uint32_t)x * kTileCols;
      if (k > 0) {
        mbar_wait(bar(B_RES_CHAIN, x), (par >> x) & 1u); par ^= (1u << x);
        fence_after();
        uint32_t v[16];
        CLB_TMEM_LD16(tm + cD + col, v);
        wait_ld();
        const float* bk = bsm + (size_t)(k - 1) * WP + HW * hf;
#pragma unroll
        for (int j = 0; j < HW; ++j) { const float o = __uint_as_float(v[j]) + bk[j]; t.x[j] = fmaxf(o, kLeak * o); }
        if (train && k < LT) {                                 // a_k for the backward pass; a_LT stays in registers for the head
          float4* dst = scr0 + x * scr_tile + ((size_t)(k - 1) * NC + 4 * hf) * TR + rrow;
#pragma unroll
          for (int c = 0; c < 4; ++c) dst[(size_t)c * TR] = make_float4(t.x[4 * c], t.x[4 * c + 1], t.x[4 * c + 2], t.x[4 * c + 3]);
        }
      }
      if (k < LT) {
        uint32_t hi[16], lo[16];
        split16(t.x, hi, lo);
        CLB_TMEM_ST16(tm + cAhi + col, hi);
        CLB_TMEM_ST16(tm + cAlo + col, lo);
        wait_st();
        fence_before();
        warp_arrive(bar(B_OPND_CHAIN, x), lane);
      }
    };

    // collect tile x's dW product of layer `layer`: D rows at lanes (r % 16) + 32 (r / 16), see tc::collect_dw_red
    auto collect_dw = [&](auto X, int layer) {
      constexpr int x = decltype(X)::value;
      const uint32_t tm = tm0 + (uint32_t)x * kTileCols;
      mbar_wait(bar(B_DW_DONE, x), (par >> (2 + x)) & 1u); par ^= (4u << x);
      fence_after();
      uint32_t v0[16], v1[16];
      CLB_TMEM_LD16(tm + cDw + col, v0);
      CLB_TMEM_LD16(tm + cDw + 32 + col, v1);
      wait_ld();
      if (lane < 16) {
        // M = 64 product: lanes < 16 of quarter q hold rows 16 q + lane, i.e. dW row i = (16 q + lane) % 32; tc::dw_slot32 layout
        float4* dst = reinterpret_cast<float4*>(part32 + (size_t)layer * PSLOT) + (4 * hf) * 32 + ((16 * (warp & 3) + lane) & 31);
#pragma unroll
        for (int qq = 0; qq < 4; ++qq)
          atomicAdd(dst + qq * 32, make_float4(__uint_as_float(v0[4 * qq]) + __uint_as_float(v1[4 * qq]),
                                               __uint_as_float(v0[4 * qq + 1]) + __uint_as_float(v1[4 * qq + 1]),
                                               __uint_as_float(v0[4 * qq + 2]) + __uint_as_float(v1[4 * qq + 2]),
                                               __uint_as_float(v0[4 * qq + 3]) + __uint_as_float(v1[4 * qq + 3])));
      }
    };

    // my half of a_k, the input of layer k (k > 0: scratch slot k-1; k == 0: the metadata columns)
    auto load_act = [&](auto X, int k) {
      constexpr int x = decltype(X)::value;
      Tile& t = T[x];
      if (k > 0) {
        const float4* src = scr0 + x * scr_tile + ((size_t)(k - 1) * NC + 4 * hf) * TR + rrow;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float4 v = __ldcg(src + (size_t)c * TR);
          t.y[4 * c] = v.x; t.y[4 * c + 1] = v.y; t.y[4 * c + 2] = v.z; t.y[4 * c + 3] = v.w;
        }
      } else {
        const int64_t row = tile_row(x);
#pragma unroll
        for (int i = 0; i < HW; ++i) { const int f = HW * hf + i; t.y[i] = (t.inb && f < a.d) ? a.meta[(size_t)f * a.n_rows + row] : 0.f; }
      }
    };

    // ---- backward of layer k (k == LT: the Dense(2) head, dW only) ----
    auto bwd_step = [&](auto X, int k) {
      constexpr int x = decltype(X)::value;
      Tile& t = T[x];
      const uint32_t tm = tm0 + (uint32_t)x * kTileCols;
      char* const dw_a = img_base + (size_t)x * 2 * kDwImgBytes;
      char* const dw_b = dw_a + kDwImgBytes;
      float dp[HW], ain[HW];
      if (k == LT) {
#pragma unroll
        for (int j = 0; j < HW; ++j) { dp[j] = 0.f; ain[j] = t.x[j]; }
        if (hf == 0) { dp[0] = t.dmu; dp[1] = t.drho; }
      } else {
        if (k == LT - 1) {                                     // delta a_LT comes from the head, not from a tensor product
#pragma unroll
          for (int i = 0; i < HW; ++i) {
            const float2 w = *reinterpret_cast<const float2*>(&Whead[(HW * hf + i) * 2]);
            dp[i] = w.x * t.dmu + w.y * t.drho;
          }
        } else {
          mbar_wait(bar(B_RES_CHAIN, x), (par >> x) & 1u); par ^= (1u << x);
          fence_after();
          uint32_t v[16];
          CLB_TMEM_LD16(tm + cD + col, v);
          wait_ld();
#pragma unroll
          for (int j = 0; j < HW; ++j) dp[j] = __uint_as_float(v[j]);
        }
#pragma unroll
        for (int j = 0; j < HW; ++j) { dp[j] = ((t.mask >> j) & 1u) ? dp[j] : kLeak * dp[j]; ain[j] = t.y[j]; }
      }
      const bool need_dx = k > 0 && k < LT;
      uint32_t hi[16], lo[16];
      split16(dp, hi, lo);
      if (need_dx) {                                           // the critical product first: delta a_k = delta p_k W_k^T
        CLB_TMEM_ST16(tm + cAhi + col, hi);
        CLB_TMEM_ST16(tm + cAlo + col, lo);
        wait_st();
        fence_before();
        warp_arrive(bar(B_OPND_CHAIN, x), lane);
      }
      if (k < LT) collect_dw(X, k + 1);                        // the layer above: its operand images and accumulator become free
      swap_blocks(hi, sw); swap_blocks(lo, sw);
      dw_store_half(dw_b, rrow, hf, hi, lo, sw);
      {
        uint32_t a2[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) a2[j] = __float_as_uint(ain[j]);
        swap_blocks(a2, sw);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          hi[j] = a2[j];
          lo[j] = __float_as_uint(__uint_as_float(a2[j]) - __uint_as_float(a2[j] & 0xFFFFE000u));
        }
      }
      dw_store_half(dw_a, rrow, hf, hi, lo, sw);
      fence_async_smem();
      warp_arrive(bar(B_OPND_DW, x), lane);
      // in the shadow of the tensor pipe: sign mask of a_k, bias gradient, the next layer's activations, dead scratch lines
      unsigned m = 0u;
#pragma unroll
      for (int i = 0; i < HW; ++i) m |= (ain[i] > 0.f ? 1u : 0u) << i;
      t.mask = m;
      bias_red16(dp, part32 + (size_t)k * PSLOT + WP * WP + HW * hf, lane, 16);
      if (k > 0 && k < LT && a.discard_scratch && (rrow & 1) == 0)
        discard_line(scr0 + x * scr_tile + ((size_t)(k - 1) * NC + 4 * hf + ((rrow & 7) >> 1)) * TR + (rrow - (rrow & 7)));
      if (k > 0) load_act(X, k - 1);
    };

    const TileIdx<0> A; const TileIdx<1> B;
    for (; pair < n_pairs; pair += gridDim.x) {
#pragma unroll
      for (int x = 0; x < 2; ++x) {
        Tile& t = T[x];
        const int64_t row = tile_row(x);
        t.inb = row < a.n_rows;
        t.refl = t.inb ? a.refl[row] : -1;
        t.active = t.refl >= 0;
#pragma unroll
        for (int i = 0; i < HW; ++i) { const int f = HW * hf + i; t.x[i] = (t.inb && f < a.d) ? a.meta[(size_t)f * a.n_rows + row] : 0.f; }
      }
      for (int k = 0; k <= LT; ++k) { fwd_step(A, k); fwd_step(B, k); }
      // ---- head: partial dot products of both halves; the hf = 0 warps finish tile A, the hf = 1 warps tile B ----
      {
        float p0[2], p1[2];
#pragma unroll
        for (int x = 0; x < 2; ++x) {
          p0[x] = 0.f; p1[x] = 0.f;
#pragma unroll
          for (int i = 0; i < HW; ++i) {
            const float2 w = *reinterpret_cast<const float2*>(&Whead[(HW * hf + i) * 2]);
            p0[x] = fmaf(T[x].x[i], w.x, p0[x]); p1[x] = fmaf(T[x].x[i], w.y, p1[x]);
          }
        }
        const int other = hf ^ 1;                              // I finish tile `hf`; the other half's partial of tile `other` goes to its owner
        xch[other * TR + rrow] = hf ? make_float2(p0[0], p1[0]) : make_float2(p0[1], p1[1]);
        workers_sync();
        const float2 o = xch[hf * TR + rrow];
        const float out0 = (hf ? p0[1] : p0[0]) + o.x + bsm[L * WP], out1 = (hf ? p1[1] : p1[0]) + o.y + bsm[L * WP + 1];
        const int64_t erow = tile_row(hf);                     // (no dynamic indexing of T: it must stay in registers)
        const bool einb = hf ? T[1].inb : T[0].inb, eact = hf ? T[1].active : T[0].active;
        const int erefl = hf ? T[1].refl : T[0].refl;
        float dmu, drho;
        obs_epilogue<LIK>(a, erow, einb, eact, erefl, lane, out0, out1, ev_f, ev_a, ev_b, ll_sum, dmu, drho);
        xch[(2 + hf) * TR + rrow] = make_float2(dmu, drho);
        workers_sync();
#pragma unroll
        for (int x = 0; x < 2; ++x) { const float2 g = xch[(2 + x) * TR + rrow]; T[x].dmu = g.x; T[x].drho = g.y; }
      }
      if (!train) continue;
      // ---- backward ----
      // every step leaves behind the sign mask of its own input a_k (leaky' of the layer below) and the prefetched a_{k-1}
      for (int k = LT; k >= 0; --k) { bwd_step(A, k); bwd_step(B, k); }
      collect_dw(A, 0); collect_dw(B, 0);
    }
    // ---- flush: the log-likelihood sum ----
    ll_sum = warp_sum(ll_sum);
    if (lane == 0) red[warp] = ll_sum;
    fence_before();
  }
  __syncthreads();
  if (tid == 0) {
    double t = 0.0;
    for (int i = 0; i < kWorkers / 32; ++i) t += red[i];
    flush_ll(a.ll_part, a.acc, t);
  }
  if (is_issuer) { fence_after(); tmem_dealloc512(tbase); }
}

}  // namespace clb
