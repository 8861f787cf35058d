// clb_pp.cuh -- k_obs_pp: the width-32 observation kernel as a two-tile ping-pong, two CTAs per SM (sm_100a).
//
// Same arithmetic, operand images, tensor-memory products and global layouts as k_obs_tc2 (clb_kernels.cuh / clb_tc.cuh) --
// what changes is WHO waits for WHAT.  k_obs_tc2 runs one 128-row tile per CTA: all 8 warps prepare operands, meet at a
// __syncthreads(), one of them issues the tcgen05.mma's and then everybody sits at an mbarrier until the tensor pipe
// answers (ncu: 35 % of the stall samples are that wait plus the barrier; issue slots 49 % busy, tensor pipe 31 %): with one
// tile per CTA and two CTAs per SM only TWO strictly serial layer chains are in flight per SM.
// Here every CTA owns TWO tiles (A, B) and two CTAs share an SM: FOUR chains in flight per SM.
//   * the 8 warps (two threads per row as before) alternate between the tiles at pass granularity: they hand tile A's
//     operands to the tensor pipe and, instead of waiting, do tile B's share of the same layer; by the time they come back to
//     A its products have long finished.  Hand-over is an arrival counter in shared memory (one acq_rel atomic per warp), never
//     a CTA barrier: the warp that arrives LAST finds all operands in place and issues the tcgen05.mma's itself (and starts
//     the TMA fetch of the next layer's weight images, one fetch serving both tiles -- they move through the layers in lockstep);
//   * the backward step of a layer starts with the critical chain (delta-a = delta-p W^T) and only then builds the dW operand
//     images.  To fit two CTAs per SM the two tiles SHARE one pair of dW operand images (64 KB) and one dW accumulator: a
//     step first collects the dW product of the PREVIOUS step (the other tile's) -- which frees both -- and then stores its
//     own images; the chain hand-over and the bias-gradient shuffles sit between the previous step's dW hand-over and its
//     collection, in the shadow of the 16 dW MMAs;
//   * registers (128 per thread): between two backward steps a tile's only live state is a 16-bit sign mask (leaky' of the
//     layer below) -- delta-p comes back from tensor memory, and the layer input is prefetched from the scratch for the NEXT
//     step only (one 16-register buffer serves both tiles).
// (Round 2's first ping-pong had one CTA per SM, a dedicated issuer warp and per-tile dW buffers: two tiles in flight per SM like
// k_obs_tc2, two warps per scheduler, 18.7 ms against 16.8 ms -- see DESIGN.md 4.2.)
// Tensor memory: 256 columns per CTA: per tile A_hi 0, A_lo 32, D 64 (tile B: + 96); shared dW accumulator at 192 (64 columns).
// Shared memory: one pair of MN-major dW operand images (64 KB), one double-buffered weight image pair for both tiles.
// Included by clb_kernels.cuh (needs ObsArgs, obs_epilogue, bias_red16, discard_line).
#pragma once

namespace clb {
namespace pp {

using namespace tc;

constexpr int kThreadsPP = 256;               // 8 warps, two threads per row
#ifndef CLB_PP_CTAS
#define CLB_PP_CTAS 2
#endif
constexpr int kCtasPerSM = CLB_PP_CTAS;
constexpr int TR = 128;                       // rows per tile
constexpr uint32_t kTileCols = 96;            // tensor-memory columns per tile (A_hi, A_lo, D)
constexpr uint32_t cAhi = 0, cAlo = 32, cD = 64, cDw = 192;     // cDw: the dW accumulator shared by both tiles (absolute column)
constexpr int PSLOT = 32 * 32 + 32;
// mbarriers: [x] = tile; the dW accumulator is shared by the tiles (one barrier, the steps alternate)
enum { B_RES_CHAIN = 0, B_DW_DONE = 2, B_WIMG = 3, N_BARS = 5 };
// arrival counters (uint32, wrap to 0 at the last arrival): chain operands of tile 0 / 1, dW operands
enum { C_CHAIN = 0, C_DW = 2, N_CNT = 3 };

struct SmemPP {
  static size_t bytes(int n_layers) {
    return 2 * (size_t)kDwImgBytes + 4 * (size_t)kImgBytes + sizeof(float) * (64 + (size_t)n_layers * 32)
           + 4 * TR * sizeof(float2) + 64 * sizeof(double) + N_BARS * sizeof(uint64_t) + 64 + 1024;
  }
};

// per-thread state of one of the two tiles
// (everything that is a fixed offset from a per-CTA base -- tensor-memory columns, barriers, operand images, scratch -- is
// derived from the compile-time tile index at the point of use instead of living in registers)
struct Tile {
  int refl; bool inb, active;
  float x[16];                 // forward: my half of the activation vector; afterwards the layer input a_LT for the head
  float dmu, drho;
  unsigned mask;
};
template <int X> struct TileIdx { static constexpr int value = X; };

}  // namespace pp

template <int LIK>
__global__ void __launch_bounds__(pp::kThreadsPP, pp::kCtasPerSM) k_obs_pp(ObsArgs a) {
  using namespace pp;
  constexpr int WP = 32, NC = 8, HW = 16;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const int NL = a.lay.n_layers, L = NL - 1, LT = L;           // no image layers in this kernel
  unsigned char* sp = smem_raw;
  char* img_base = reinterpret_cast<char*>(sp);                // [a_hi a_lo | dp_hi dp_lo], shared by the two tiles
  sp += 2 * (size_t)kDwImgBytes;
  char* w_img = reinterpret_cast<char*>(sp);                   // [2 buffers][hi, lo][kImgBytes]
  sp += 4 * (size_t)kImgBytes;
  float* Whead = reinterpret_cast<float*>(sp);                 // [32][2]
  float* bsm = Whead + 64;                                     // [NL][32]
  float2* xch = reinterpret_cast<float2*>(bsm + (size_t)NL * WP);   // [0..1][TR]: head partial sums for tile A / B; [2..3][TR]: (dmu, drho)
  double* red = reinterpret_cast<double*>(xch + 4 * TR);
  uint64_t* bars = reinterpret_cast<uint64_t*>(red + 64);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bars + N_BARS); // [0] tensor-memory base, [1 + c] arrival counters

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    for (int b = 0; b < N_BARS; ++b) mbar_init(smem_u32(bars + b), 1);
    for (int c = 0; c < N_CNT; ++c) slot[1 + c] = 0u;
  }
  if (warp == 0) tmem_alloc(smem_u32(slot));
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

  const int rrow = tid & (TR - 1), hf = tid >> 7;
  const uint32_t col = (uint32_t)(HW * hf);
  Tile T[2];
  const uint32_t tm0 = tbase + ((uint32_t)(32 * (warp & 3)) << 16);      // my row in tile 0's column block (+ kTileCols for tile 1)
  const uint32_t bar0 = smem_u32(bars), cnt0 = smem_u32(slot + 1);
  const uint32_t wbuf0 = smem_u32(w_img);
  float4* const scr0 = a.scratch + (size_t)blockIdx.x * 2 * (size_t)LT * NC * TR;     // [tile][LT][8][TR]
  const size_t scr_tile = (size_t)LT * NC * TR;
  uint32_t par = 0;                                            // phase bits: bit x = res_chain of tile x, bit 2 = dw_done
  uint32_t pass = 0;                                           // weight-image passes so far: buffer = pass & 1, phase = (pass >> 1) & 1
  float* part32 = a.partials32 + (size_t)(blockIdx.x % a.n_partials) * NL * PSLOT;
  double ll_sum = 0.0;
  float ev_f = 1.f, ev_a = 0.f, ev_b = 0.f;
  if (a.theta_lik != nullptr) { ev_f = softplusf(a.theta_lik[0]); ev_a = softplusf(a.theta_lik[1]); ev_b = softplusf(a.theta_lik[2]); }
  const int sw = (rrow >> 2) & 1;                              // conflict-free image stores: see tc::dw_store_half
  int64_t pair = blockIdx.x;
  auto tile_row = [&](int x) -> int64_t { return (2 * pair + x) * TR + rrow; };
  auto bar = [&](int which) -> uint32_t { return bar0 + 8u * (uint32_t)which; };

  // the first pass's images (forward, layer 0) start travelling now
  if (tid == 0 && LT > 0 && pair < n_pairs) tma_fetch_image(wbuf0, gimg(0, 0), bar(B_WIMG));

  // ---- issue (all 32 lanes of the warp that arrived last) ----
  // one chain pass of tile x from weight buffer (pass & 1): 12 MMAs (X_hi W_lo + X_lo W_hi + X_hi W_hi) into the tile's D.
  // `next` (tile 1 only): both tiles' previous pass has been consumed by every warp, so the other buffer is free for the
  // next pass's images.
  auto chain_issue = [&](int x, const float* next) {
    fence_after();
    const uint32_t b = pass & 1u;
    const uint32_t tb = uniform32(tbase) + (uint32_t)x * kTileCols;
    const uint32_t wb = wbuf0 + b * 2u * kImgBytes;
    const uint64_t bhi = uniform64(make_desc(wb)), blo = uniform64(make_desc(wb + kImgBytes));
    const uint32_t d = tb + cD;
    const uint32_t rbar = uniform32(bar(B_RES_CHAIN + x));
    const uint32_t wbar = uniform32(bar(B_WIMG + (int)b)), wph = uniform32((pass >> 1) & 1u);
    const uint32_t nbar = uniform32(bar(B_WIMG + (int)(b ^ 1u))), ndst = uniform32(wbuf0 + (b ^ 1u) * 2u * kImgBytes);
    if (elect_one()) {
      mbar_wait(wbar, wph);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) mma_tf32_ts(d, tb + cAhi + 8u * (uint32_t)ks, blo + (uint64_t)((2u * kLBO * (uint32_t)ks) >> 4), ks > 0 ? 1u : 0u);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) mma_tf32_ts(d, tb + cAlo + 8u * (uint32_t)ks, bhi + (uint64_t)((2u * kLBO * (uint32_t)ks) >> 4), 1u);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) mma_tf32_ts(d, tb + cAhi + 8u * (uint32_t)ks, bhi + (uint64_t)((2u * kLBO * (uint32_t)ks) >> 4), 1u);
      commit(rbar);
      if (next != nullptr) tma_fetch_image(ndst, next, nbar);
    }
    __syncwarp();
  };
  auto dw_issue = [&]() {
    fence_after();
    const uint32_t d = uniform32(tbase) + cDw;
    const uint64_t a0 = uniform64(make_desc_mn(smem_u32(img_base))), b0 = uniform64(make_desc_mn(smem_u32(img_base + kDwImgBytes)));
    const uint32_t dbar = uniform32(bar(B_DW_DONE));
    if (elect_one()) {
#ifndef CLB_PP_ABL_DW
#pragma unroll
      for (int ks = 0; ks < TR / 8; ++ks)
        mma_tf32_ss(d, a0 + (uint64_t)((2u * kDwSBO * (uint32_t)ks) >> 4), b0 + (uint64_t)((2u * kDwSBO * (uint32_t)ks) >> 4), kIdescDw, ks > 0 ? 1u : 0u);
#endif
      commit(dbar);
    }
    __syncwarp();
  };

  // ---- forward: consume the products of pass k-1 (k > 0), hand the operands of pass k (k < LT) over ----
  auto fwd_step = [&](auto X, int k, bool more) {
    constexpr int x = decltype(X)::value;
    Tile& t = T[x];
    const uint32_t tm = tm0 + (uint32_t)x * kTileCols;
    if (k > 0) {
      mbar_wait(bar(B_RES_CHAIN + x), (par >> x) & 1u); par ^= (1u << x);
      fence_after();
      uint32_t v[16];
      CLB_TMEM_LD16(tm + cD + col, v);
      wait_ld();
      const float* bk = bsm + (size_t)(k - 1) * WP + HW * hf;
#pragma unroll
      for (int j = 0; j < HW; ++j) { const float o = __uint_as_float(v[j]) + bk[j]; t.x[j] = fmaxf(o, kLeak * o); }
#ifndef CLB_PP_ABL_SCR
      if (train && k < LT) {                                   // a_k for the backward pass; a_LT stays in registers for the head
        float4* dst = scr0 + x * scr_tile + ((size_t)(k - 1) * NC + 4 * hf) * TR + rrow;
#pragma unroll
        for (int c = 0; c < 4; ++c) dst[(size_t)c * TR] = make_float4(t.x[4 * c], t.x[4 * c + 1], t.x[4 * c + 2], t.x[4 * c + 3]);
      }
#endif
    }
    if (k < LT) {
      uint32_t hi[16], lo[16];
      split16(t.x, hi, lo);
      CLB_TMEM_ST16(tm + cAhi + col, hi);
      CLB_TMEM_ST16(tm + cAlo + col, lo);
      wait_st();
      fence_before();
      if (arrive_last(cnt0 + 4u * (uint32_t)(C_CHAIN + x), lane, 7u)) {
        // the pass after this one: next forward layer, else the first dX pass, else the next pair's first layer
        const float* next = (x == 0) ? nullptr
                            : (k + 1 < LT) ? gimg(k + 1, 0) : (train && LT > 1) ? gimg(LT - 1, 1) : (more ? gimg(0, 0) : nullptr);
        chain_issue(x, next);
      }
      if (x == 1) pass += 1u;
    }
  };

  // collect the pending dW product (the PREVIOUS step's, either tile) into layer `layer`'s slot of the partial: D rows at lanes
  // (r % 16) + 32 (r / 16), see tc::collect_dw_red.  Frees the shared operand images and the accumulator for this step.
  auto collect_dw = [&](int layer) {
    mbar_wait(bar(B_DW_DONE), (par >> 2) & 1u); par ^= 4u;
    fence_after();
    uint32_t v0[16], v1[16];
    CLB_TMEM_LD16(tm0 + cDw + col, v0);
    CLB_TMEM_LD16(tm0 + cDw + 32 + col, v1);
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

  // my half of a_k of tile x, the input of layer k (k > 0: scratch slot k-1; k == 0: the metadata columns), prefetched into
  // `y` for the NEXT backward step (one buffer serves both tiles: the steps alternate)
  float y[16];
  auto load_act = [&](auto X, int k) {
    constexpr int x = decltype(X)::value;
    if (k > 0) {
      const float4* src = scr0 + x * scr_tile + ((size_t)(k - 1) * NC + 4 * hf) * TR + rrow;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
#ifdef CLB_PP_ABL_SCR
        const float4 v = make_float4(0.1f * c + 1e-3f * rrow, -0.2f, 0.3f + k, -0.4f);
#else
        const float4 v = __ldcg(src + (size_t)c * TR);
#endif
        y[4 * c] = v.x; y[4 * c + 1] = v.y; y[4 * c + 2] = v.z; y[4 * c + 3] = v.w;
      }
    } else {
      const int64_t row = tile_row(x);
      const bool inb = T[x].inb;
#pragma unroll
      for (int i = 0; i < HW; ++i) { const int f = HW * hf + i; y[i] = (inb && f < a.d) ? __ldcs(&a.meta[(size_t)f * a.n_rows + row]) : 0.f; }
    }
  };

  // ---- backward of layer k (k == LT: the Dense(2) head, dW only) ----
  auto bwd_step = [&](auto X, int k, bool more) {
    constexpr int x = decltype(X)::value;
    Tile& t = T[x];
    const uint32_t tm = tm0 + (uint32_t)x * kTileCols;
    char* const dw_a = img_base;
    char* const dw_b = dw_a + kDwImgBytes;
    float dp[HW];
    if (k == LT) {
#pragma unroll
      for (int j = 0; j < HW; ++j) { dp[j] = 0.f; y[j] = t.x[j]; }
      if (hf == 0) { dp[0] = t.dmu; dp[1] = t.drho; }
    } else {
      if (k == LT - 1) {                                       // delta a_LT comes from the head, not from a tensor product
#pragma unroll
        for (int i = 0; i < HW; ++i) {
          const float2 w = *reinterpret_cast<const float2*>(&Whead[(HW * hf + i) * 2]);
          dp[i] = w.x * t.dmu + w.y * t.drho;
        }
      } else {
        mbar_wait(bar(B_RES_CHAIN + x), (par >> x) & 1u); par ^= (1u << x);
        fence_after();
        uint32_t v[16];
        CLB_TMEM_LD16(tm + cD + col, v);
        wait_ld();
#pragma unroll
        for (int j = 0; j < HW; ++j) dp[j] = __uint_as_float(v[j]);
      }
#pragma unroll
      for (int j = 0; j < HW; ++j) dp[j] = ((t.mask >> j) & 1u) ? dp[j] : kLeak * dp[j];
    }
    const bool need_dx = k > 0 && k < LT;
    uint32_t hi[16], lo[16];
    split16(dp, hi, lo);
    if (need_dx) {                                             // the critical product first: delta a_k = delta p_k W_k^T
      CLB_TMEM_ST16(tm + cAhi + col, hi);
      CLB_TMEM_ST16(tm + cAlo + col, lo);
      wait_st();
      fence_before();
      if (arrive_last(cnt0 + 4u * (uint32_t)(C_CHAIN + x), lane, 7u)) {
        // the pass after this layer's dX: the next layer's dX (k - 1 >= 1), else the next pair's first forward layer
        const float* next = (x == 0) ? nullptr : (k > 1) ? gimg(k - 1, 1) : (more ? gimg(0, 0) : nullptr);
        chain_issue(x, next);
      }
      if (x == 1) pass += 1u;
    }
    // in the shadow of the previous step's dW product: the bias gradient (column sums of delta-p)
    bias_red16(dp, part32 + (size_t)k * PSLOT + WP * WP + HW * hf, lane, 16);
    // the previous step's dW (the other tile's): its operand images and the accumulator become free
    if (x == 1 || k < LT) collect_dw(x == 1 ? k : k + 1);
    swap_blocks(hi, sw); swap_blocks(lo, sw);
    dw_store_half(dw_b, rrow, hf, hi, lo, sw);
    unsigned m = 0u;                                           // sign mask of a_k: leaky' of the layer below
    {
      uint32_t a2[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) { a2[j] = __float_as_uint(y[j]); m |= (y[j] > 0.f ? 1u : 0u) << j; }
      swap_blocks(a2, sw);
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        hi[j] = a2[j];
        lo[j] = __float_as_uint(__uint_as_float(a2[j]) - __uint_as_float(a2[j] & 0xFFFFE000u));
      }
    }
    t.mask = m;
    dw_store_half(dw_a, rrow, hf, hi, lo, sw);
    fence_async_smem();
    fence_before();                                            // my tcgen05.ld of the shared accumulator precedes the next dW product
    if (arrive_last(cnt0 + 4u * (uint32_t)C_DW, lane, 7u)) dw_issue();
    if (k > 0 && k < LT && a.discard_scratch && (rrow & 1) == 0)
      discard_line(scr0 + x * scr_tile + ((size_t)(k - 1) * NC + 4 * hf + ((rrow & 7) >> 1)) * TR + (rrow - (rrow & 7)));
    // the next step's layer input: (A, k) -> (B, k) -> (A, k - 1); the head's input a_LT is still in registers (t.x)
    if (x == 0) { if (k < LT) load_act(TileIdx<1>{}, k); }
    else if (k > 0) load_act(TileIdx<0>{}, k - 1);
  };

  const TileIdx<0> A; const TileIdx<1> B;
  for (; pair < n_pairs; pair += gridDim.x) {
    const bool more = pair + gridDim.x < n_pairs;
#pragma unroll
    for (int x = 0; x < 2; ++x) {
      Tile& t = T[x];
      const int64_t row = tile_row(x);
      t.inb = row < a.n_rows;
      t.refl = t.inb ? __ldcs(&a.refl[row]) : -1;
      t.active = t.refl >= 0;
#pragma unroll
      for (int i = 0; i < HW; ++i) { const int f = HW * hf + i; t.x[i] = (t.inb && f < a.d) ? __ldcs(&a.meta[(size_t)f * a.n_rows + row]) : 0.f; }
    }
    for (int k = 0; k <= LT; ++k) { fwd_step(A, k, more); fwd_step(B, k, more); }
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
      const int other = hf ^ 1;                                // I finish tile `hf`; the other half's partial of tile `other` goes to its owner
      xch[other * TR + rrow] = hf ? make_float2(p0[0], p1[0]) : make_float2(p0[1], p1[1]);
      __syncthreads();
      const float2 o = xch[hf * TR + rrow];
      const float out0 = (hf ? p0[1] : p0[0]) + o.x + bsm[L * WP], out1 = (hf ? p1[1] : p1[0]) + o.y + bsm[L * WP + 1];
      const int64_t erow = tile_row(hf);                       // (no dynamic indexing of T: it must stay in registers)
      const bool einb = hf ? T[1].inb : T[0].inb, eact = hf ? T[1].active : T[0].active;
      const int erefl = hf ? T[1].refl : T[0].refl;
      float dmu, drho;
      obs_epilogue<LIK>(a, erow, einb, eact, erefl, lane, out0, out1, ev_f, ev_a, ev_b, ll_sum, dmu, drho);
      xch[(2 + hf) * TR + rrow] = make_float2(dmu, drho);
      __syncthreads();
#pragma unroll
      for (int x = 0; x < 2; ++x) { const float2 g = xch[(2 + x) * TR + rrow]; T[x].dmu = g.x; T[x].drho = g.y; }
    }
    if (!train) continue;
    // ---- backward ----
    // every step leaves behind the sign mask of its own input a_k (leaky' of the layer below) and the prefetched input of the next step
    for (int k = LT; k >= 0; --k) { bwd_step(A, k, more); bwd_step(B, k, more); }
    collect_dw(0);                                             // tile B's layer 0
  }
  // ---- flush: the log-likelihood sum ----
  ll_sum = warp_sum(ll_sum);
  if (lane == 0) red[warp] = ll_sum;
  fence_before();
  __syncthreads();
  if (tid == 0) {
    double t = 0.0;
    for (int i = 0; i < kThreadsPP / 32; ++i) t += red[i];
    flush_ll(a.ll_part, a.acc, t);
  }
  if (warp == 0) { fence_after(); tmem_dealloc(tbase); }
}

}  // namespace clb
