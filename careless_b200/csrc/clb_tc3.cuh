// clb_tc3.cuh -- k_obs_tc3: k_obs_tc2 software-pipelined ACROSS TILES (sm_100a, padded width 32, no image layers).
//
// k_obs_tc2 walks one 128-row tile at a time through forward, head / likelihood epilogue and backward.  Its forward pass is
// 28 % of the time although it holds only ~15 % of the instructions: a forward layer is a short burst of thread work (bias,
// LeakyReLU, operand split) followed by a full tensor-pipe round trip (store operands -> barrier -> issue -> 12 MMAs -> mbarrier
// -> tcgen05.ld) with nothing else to do in the CTA.  The backward pass has the opposite problem: between handing a layer's
// products over and collecting them the warps idle.
// Here the FORWARD PASS OF THE NEXT TILE (B) is folded into the BACKWARD PASS OF THE CURRENT TILE (A): backward step p of A
// (p = 0: the head's dW; p = 1 .. LT: hidden layer LT - p) carries forward stage p of B in the bubble after its hand-over:
//      A: operand split, tensor-memory / shared-memory stores        -> fence, __syncthreads()
//      warp 0 issues  B chain(p - 1)  [operands stored one step ago], then  A chain (delta-a = delta-p W^T);  warp 7: A's dW
//      A: bias-gradient shuffles
//      B: stage p = collect chain(p - 1), bias + LeakyReLU, activation scratch, split + tcgen05.st of the next operand
//      A: collect delta-a, leaky';  collect dW (tensor memory -> REDs into the FP32 partial)
// B's products are issued FIRST, so they are ready when the warps reach B's stage and A's chain finishes while they work on
// it; B's operands ride on A's fence + barrier: no additional CTA barrier, no additional tensor-pipe round trip.  A's backward
// has LT + 1 steps and B's forward LT + 1 stages, so they pair exactly; afterwards B's head / epilogue runs and B becomes A.
// Only the first tile of a CTA runs its forward pass alone.
// Resources (two CTAs per SM as before): tensor memory 256 columns = two operand / accumulator sets (A_hi, A_lo, D: 0 / 32 / 64
// and 96 / 128 / 224; a tile keeps its set from forward to backward) + the dW accumulator (160); shared memory + 17 KB for the
// second weight-image stream (forward and backward layers need different images at the same time); two activation scratch
// buffers per CTA (one fills while the other drains: the LIVE bytes stay one tile's worth, dead lines are discarded from L2);
// registers: A's layer input is remembered as a 16-bit sign mask, B has no register state between its stages.
// Included by clb_kernels.cuh (needs ObsArgs, obs_epilogue, bias_red16).
#pragma once

#ifndef CLB_TC3_EARLY
#define CLB_TC3_EARLY 0     // 1: B's forward stage runs before A's barrier and its chain is issued at THIS step's barrier (collected a step later)
#endif

namespace clb {
namespace tc3 {

using namespace tc;

constexpr int kThreads3 = 256;
constexpr int TR = 128;
__device__ __forceinline__ uint32_t col_hi(uint32_t s) { return s ? 96u : 0u; }
__device__ __forceinline__ uint32_t col_lo(uint32_t s) { return s ? 128u : 32u; }
__device__ __forceinline__ uint32_t col_d(uint32_t s) { return s ? 224u : 64u; }
// mbarriers
enum { B_CHAIN = 0 /* +set */, B_DW = 2, B_WB = 3 /* +buffer: backward weight images */, B_WF = 5 /* +buffer: forward */, N_BARS = 7 };

struct Smem3 {
  static size_t bytes(int n_layers) {
    return 2 * (size_t)kDwImgBytes + sizeof(float) * (64 + (size_t)n_layers * 32) + 64 * sizeof(double) + 8 * (size_t)kImgBytes
           + N_BARS * sizeof(uint64_t) + 64 + 4 * 128 * sizeof(float) + 128;
  }
};

// one weight-image stream (forward or backward): two buffers [hi | lo], pass counter (buffer = pass & 1, phase = (pass >> 1) & 1)
struct Stream {
  uint32_t img0;      // shared-memory address of buffer 0 (buffer 1 follows)
  uint32_t bar0;      // mbarrier of buffer 0 (buffer 1 follows)
  uint32_t pass;
};

// 12 MMAs of one chain pass on operand set s with the stream's current images, then the stream's next fetch.  Called by all
// 32 lanes of the issuing warp: the operands are made warp-uniform first (no R2UR waterfall per MMA, see clb_tc.cuh), one
// elected lane issues.
__device__ __forceinline__ void chain_mmas(uint32_t tbase, uint32_t s, const Stream& w, uint32_t res_bar, const float* next) {
  const uint32_t b = w.pass & 1u;
  const uint32_t wb = uniform32(w.img0 + b * 2u * kImgBytes), nwb = uniform32(w.img0 + (b ^ 1u) * 2u * kImgBytes);
  const uint64_t bhi = uniform64(make_desc(wb)), blo = uniform64(make_desc(wb + kImgBytes));
  const uint32_t tb = uniform32(tbase);
  const uint32_t d = tb + uniform32(col_d(s)), ahi = tb + uniform32(col_hi(s)), alo = tb + uniform32(col_lo(s));
  const uint32_t wbar = uniform32(w.bar0 + 8u * b), nbar = uniform32(w.bar0 + 8u * (b ^ 1u)), wph = uniform32((w.pass >> 1) & 1u);
  const uint32_t rbar = uniform32(res_bar);
  if (elect_one()) {
    mbar_wait(wbar, wph);
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) mma_tf32_ts(d, ahi + 8u * (uint32_t)ks, blo + (uint64_t)((2u * kLBO * (uint32_t)ks) >> 4), ks > 0 ? 1u : 0u);
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) mma_tf32_ts(d, alo + 8u * (uint32_t)ks, bhi + (uint64_t)((2u * kLBO * (uint32_t)ks) >> 4), 1u);
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) mma_tf32_ts(d, ahi + 8u * (uint32_t)ks, bhi + (uint64_t)((2u * kLBO * (uint32_t)ks) >> 4), 1u);
    commit(rbar);
    if (next != nullptr) tma_fetch_image(nwb, next, nbar);
  }
  __syncwarp();
}

}  // namespace tc3

template <int LIK>
__global__ void __launch_bounds__(tc3::kThreads3, 2) k_obs_tc3(ObsArgs a) {
  using namespace tc3;
  constexpr int WP = 32, NC = 8, HW = 16;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const int NL = a.lay.n_layers, L = NL - 1, LT = L;
  unsigned char* sp = smem_raw;
  char* dw_a = reinterpret_cast<char*>(sp);                 // dW A operand: [a_hi | a_lo]
  char* dw_b = dw_a + kDwImgBytes;                          // dW B operand: [delta-p_hi | delta-p_lo]
  sp += 2 * kDwImgBytes;
  float* Whead = reinterpret_cast<float*>(sp);              // [32][2]
  float* bsm = Whead + 64;                                  // [NL][32]
  double* red = reinterpret_cast<double*>(bsm + (size_t)NL * WP);
  char* w_img = reinterpret_cast<char*>(red + 64);          // backward stream [2][hi, lo], then forward stream [2][hi, lo]
  uint64_t* bars = reinterpret_cast<uint64_t*>(w_img + 8 * kImgBytes);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bars + N_BARS);
  float2* xch = reinterpret_cast<float2*>(slot + 4);        // [2][128]: head partial sums of hf = 1, then (dmu, drho)

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, rrow = tid & (TR - 1), hf = tid >> 7;
  const uint32_t col = (uint32_t)(HW * hf);
  if (tid == 0) for (int b = 0; b < N_BARS; ++b) mbar_init(smem_u32(bars + b), 1);
  if (warp == 0) tmem_alloc(smem_u32(slot));
  fence_before();
  for (int idx = tid; idx < WP * 2; idx += kThreads3) {
    const int i = idx / 2, j = idx % 2;
    Whead[idx] = (i < a.lay.in_dim[L] && j < a.lay.out_dim[L]) ? a.theta_mlp[a.lay.koff[L] + i * a.lay.out_dim[L] + j] : 0.f;
  }
  for (int idx = tid; idx < NL * WP; idx += kThreads3) {
    const int k = idx / WP, j = idx % WP;
    bsm[idx] = (j < a.lay.out_dim[k]) ? a.theta_mlp[a.lay.boff[k] + j] : 0.f;
  }
  fence_async_smem();
  __syncthreads();
  fence_after();
  const uint32_t tbase = *slot;
  const uint32_t row_addr = tbase + ((uint32_t)(32 * (warp & 3)) << 16);
  const uint32_t bar0 = smem_u32(bars);
  auto bar = [&](int which) -> uint32_t { return bar0 + 8u * (uint32_t)which; };
  Stream wb{smem_u32(w_img), bar(B_WB), 0u}, wf{smem_u32(w_img + 4 * kImgBytes), bar(B_WF), 0u};
  const uint64_t desc_dwa = make_desc_mn(smem_u32(dw_a)), desc_dwb = make_desc_mn(smem_u32(dw_b));
  uint32_t par = 0;                                         // phase bits: bit s = chain barrier of operand set s, bit 2 = dW barrier
  constexpr size_t IMGF = kImgBytes / 4;
  auto gimg = [&](int k, int dir) -> const float* { return a.wimg + ((size_t)(k * 2 + dir) * 2) * IMGF; };
  constexpr int PSLOT = WP * WP + WP, BOFF = WP * WP;
  float* part32 = a.partials32 + (size_t)(blockIdx.x % a.n_partials) * NL * PSLOT;
  float4* const scr0 = a.scratch + (size_t)blockIdx.x * 2 * (size_t)LT * NC * TR;      // two buffers [LT][8][TR]: tile parity
  const size_t scr_buf = (size_t)LT * NC * TR;
  double ll_sum = 0.0;
  float ev_f = 1.f, ev_a = 0.f, ev_b = 0.f;
  if (a.theta_lik != nullptr) { ev_f = softplusf(a.theta_lik[0]); ev_a = softplusf(a.theta_lik[1]); ev_b = softplusf(a.theta_lik[2]); }
  const int64_t n_tiles = (a.n_rows + TR - 1) / TR;
  const bool train = a.train_mlp != 0;
  const int sw = (rrow >> 2) & 1;                           // conflict-free image stores: see tc::dw_store_half
  const int64_t stride = gridDim.x;

  if (tid == 0 && blockIdx.x < n_tiles) {
    tma_fetch_image(wf.img0, gimg(0, 0), wf.bar0);                                   // forward layer 0
    if (train && LT > 1) tma_fetch_image(wb.img0, gimg(LT - 1, 1), wb.bar0);         // first dX pass
  }

  // ---- forward stage p of the tile on operand set s (rows: `frow`, in bounds: `finb`; activations to scratch buffer `fscr`) ----
  // p = 0: metadata -> operand of layer 0.  p >= 1: collect layer p - 1, bias + LeakyReLU -> a_p; p < LT: scratch slot p - 1 and the
  // operand of layer p; p == LT: a_LT stays in `h` for the head.  The caller publishes the operands (fence + barrier) and issues.
  auto fwd_stage = [&](int p, uint32_t s, int64_t frow, bool finb, float4* fscr, float (&h)[HW]) {
    if (p == 0) {
#pragma unroll
      for (int i = 0; i < HW; ++i) { const int f = HW * hf + i; h[i] = (finb && f < a.d) ? __ldcs(&a.meta[(size_t)f * a.n_rows + frow]) : 0.f; }
    } else {
      mbar_wait(bar(B_CHAIN + (int)s), (par >> s) & 1u); par ^= (1u << s);
      fence_after();
      uint32_t v[HW];
      CLB_TMEM_LD16(row_addr + col_d(s) + col, v);
      wait_ld();
      const float* bk = bsm + (size_t)(p - 1) * WP + HW * hf;
#pragma unroll
      for (int j = 0; j < HW; ++j) { const float o = __uint_as_float(v[j]) + bk[j]; h[j] = fmaxf(o, kLeak * o); }
      if (train && p < LT) {
#pragma unroll
        for (int c = 0; c < 4; ++c) fscr[((size_t)(p - 1) * NC + 4 * hf + c) * TR + rrow] = make_float4(h[4 * c], h[4 * c + 1], h[4 * c + 2], h[4 * c + 3]);
      }
    }
    if (p < LT) {
      uint32_t hi[HW], lo[HW];
      split16(h, hi, lo);
      CLB_TMEM_ST16(row_addr + col_hi(s) + col, hi);
      CLB_TMEM_ST16(row_addr + col_lo(s) + col, lo);
    }
  };
  // issue forward chain p of the tile on set s (warp 0, all lanes); `later` = a further tile will run a forward pass after this one
  auto fwd_issue = [&](int p, uint32_t s, bool later) {
    const float* next = (p + 1 < LT) ? gimg(p + 1, 0) : (later ? gimg(0, 0) : nullptr);
    chain_mmas(tbase, s, wf, bar(B_CHAIN + (int)s), next);
  };

  // ---- the forward pass of a tile alone (first tile of the CTA; every tile when nothing is trained) ----
  auto forward_alone = [&](uint32_t s, int64_t frow, bool finb, float4* fscr, float (&h)[HW], bool later) {
    for (int p = 0; p <= LT; ++p) {
      fwd_stage(p, s, frow, finb, fscr, h);
      if (p < LT) {
        wait_st();
        fence_before();
        __syncthreads();
        if (warp == 0) { fence_after(); fwd_issue(p, s, later); }
        wf.pass += 1u;
      }
    }
  };

  int64_t tile = blockIdx.x;
  uint32_t j = 0;                                           // tiles done by this CTA: operand set and scratch buffer = j & 1
  float h[HW];                                              // a_LT of the current tile (A)
  int64_t rowA = tile * TR + rrow;
  bool inbA = rowA < a.n_rows;
  int reflA = (tile < n_tiles && inbA) ? __ldcs(&a.refl[rowA]) : -1;
  if (tile < n_tiles) forward_alone(0u, rowA, inbA, scr0, h, train ? (tile + stride < n_tiles) : (tile + stride < n_tiles));

  for (; tile < n_tiles; tile += stride, ++j) {
    const uint32_t sA = j & 1u, sB = sA ^ 1u;
    float4* const scrA = scr0 + (size_t)sA * scr_buf;
    float4* const scrB = scr0 + (size_t)sB * scr_buf;
    // ---------------- head: partial dot products of both halves, epilogue in the hf = 0 threads ----------------
    float out0 = 0.f, out1 = 0.f;
#pragma unroll
    for (int i = 0; i < HW; ++i) {
      const float2 w = *reinterpret_cast<const float2*>(&Whead[(HW * hf + i) * 2]);
      out0 = fmaf(h[i], w.x, out0); out1 = fmaf(h[i], w.y, out1);
    }
    if (hf == 1) xch[rrow] = make_float2(out0, out1);
    __syncthreads();
    float dmu = 0.f, drho = 0.f;
    if (hf == 0) {
      const float2 o1 = xch[rrow];
      out0 += o1.x + bsm[L * WP]; out1 += o1.y + bsm[L * WP + 1];
      obs_epilogue<LIK>(a, rowA, inbA, reflA >= 0, reflA, lane, out0, out1, ev_f, ev_a, ev_b, ll_sum, dmu, drho);
      xch[TR + rrow] = make_float2(dmu, drho);
    }
    const int64_t tileB = tile + stride;
    const bool hasB = tileB < n_tiles;
    const int64_t rowB = tileB * TR + rrow;
    const bool inbB = hasB && rowB < a.n_rows;
    if (!train) {                                            // evaluation: forward passes only
      __syncthreads();
      if (hasB) forward_alone(sB, rowB, inbB, scrB, h, tileB + stride < n_tiles);
      rowA = rowB; inbA = inbB; reflA = inbB ? __ldcs(&a.refl[rowB]) : -1;
      continue;
    }
    const int reflB = inbB ? __ldcs(&a.refl[rowB]) : -1;
    const bool laterB = tileB + stride < n_tiles;            // a forward pass follows B's
    __syncthreads();
    { const float2 g = xch[TR + rrow]; dmu = g.x; drho = g.y; }

    // ---------------- backward of A, forward of B ----------------
    float dp[HW], nxt[HW];                                   // nxt: the layer input of the coming step (a_LT = h for the head's step)
    unsigned mask = 0u;                                      // signs of the current layer input: leaky' of the layer below
    auto load_act = [&](float (&dst)[HW], int k) {           // my half of a_k of tile A, the input of layer k
      if (k > 0) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float4 v = __ldcg(&scrA[((size_t)(k - 1) * NC + 4 * hf + c) * TR + rrow]);
          dst[4 * c] = v.x; dst[4 * c + 1] = v.y; dst[4 * c + 2] = v.z; dst[4 * c + 3] = v.w;
        }
      } else {
#pragma unroll
        for (int i = 0; i < HW; ++i) { const int f = HW * hf + i; dst[i] = (inbA && f < a.d) ? __ldcs(&a.meta[(size_t)f * a.n_rows + rowA]) : 0.f; }
      }
    };
#pragma unroll
    for (int i = 0; i < HW; ++i) nxt[i] = h[i];              // (h is free from here on: B's forward stages use it as their work array)
    for (int p = 0; p <= LT; ++p) {
      const int k = LT - p;                                  // p = 0: the head (k = LT, layer input a_LT = h); else hidden layer k
      const bool need_dx = p > 0 && k > 0;
      // ---- A: delta-p of this step ----
      if (p == 0) {
#pragma unroll
        for (int jj = 0; jj < HW; ++jj) dp[jj] = 0.f;
        if (hf == 0) { dp[0] = dmu; dp[1] = drho; }
      } else if (p == 1) {                                   // delta a_LT from the head, times leaky' of the last hidden layer
#pragma unroll
        for (int i = 0; i < HW; ++i) {
          const float2 w = *reinterpret_cast<const float2*>(&Whead[(HW * hf + i) * 2]);
          const float da = w.x * dmu + w.y * drho;
          dp[i] = ((mask >> i) & 1u) ? da : kLeak * da;
        }
      }                                                      // p >= 2: dp was collected (and masked) at the end of the previous step
      // ---- A: operands of the chain and of dW ----
      {
        uint32_t hi[HW], lo[HW];
        split16(dp, hi, lo);
        if (need_dx) {
          CLB_TMEM_ST16(row_addr + col_hi(sA) + col, hi);
          CLB_TMEM_ST16(row_addr + col_lo(sA) + col, lo);
        }
        swap_blocks(hi, sw); swap_blocks(lo, sw);
        dw_store_half(dw_b, rrow, hf, hi, lo, sw);
        uint32_t a2[HW];
        unsigned m = 0u;
#pragma unroll
        for (int i = 0; i < HW; ++i) {
          const float av = nxt[i];
          a2[i] = __float_as_uint(av); m |= (av > 0.f ? 1u : 0u) << i;
        }
        mask = m;
        swap_blocks(a2, sw);
#pragma unroll
        for (int i = 0; i < HW; ++i) {
          hi[i] = a2[i];
          lo[i] = __float_as_uint(__uint_as_float(a2[i]) - __uint_as_float(a2[i] & 0xFFFFE000u));
        }
        dw_store_half(dw_a, rrow, hf, hi, lo, sw);
      }
      if (k > 0) load_act(nxt, k - 1);                       // the next step's layer input, in flight during this step
#if CLB_TC3_EARLY
      // B: forward stage p BEFORE the barrier: it collects the chain issued at the previous step's barrier (a whole step ago) and its
      // operands are handed over at this step's barrier
      if (hasB) fwd_stage(p, sB, rowB, inbB, scrB, h);
#endif
      wait_st();
      fence_async_smem();
      fence_before();
      __syncthreads();
      // ---- issue: B's forward chain first (its operands were stored one step ago), then A's chain; warp 7: A's dW ----
      if (warp == 0) {
        fence_after();
#if CLB_TC3_EARLY
        if (need_dx) {
          const float* next = (k > 1) ? gimg(k - 1, 1) : (hasB ? gimg(LT - 1, 1) : nullptr);
          chain_mmas(tbase, sA, wb, bar(B_CHAIN + (int)sA), (LT > 1) ? next : nullptr);
        }
        if (hasB && p < LT) fwd_issue(p, sB, laterB);
#else
        if (hasB && p >= 1) fwd_issue(p - 1, sB, laterB);
        if (need_dx) {
          const float* next = (k > 1) ? gimg(k - 1, 1) : (hasB ? gimg(LT - 1, 1) : nullptr);
          chain_mmas(tbase, sA, wb, bar(B_CHAIN + (int)sA), (LT > 1) ? next : nullptr);
        }
#endif
      } else if (warp == 7) {
        fence_after();
        const uint32_t d = uniform32(tbase) + kColDw;
        const uint64_t a0 = uniform64(desc_dwa), b0 = uniform64(desc_dwb);
        const uint32_t dbar = uniform32(bar(B_DW));
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < TR / 8; ++ks)
            mma_tf32_ss(d, a0 + (uint64_t)((2u * kDwSBO * (uint32_t)ks) >> 4), b0 + (uint64_t)((2u * kDwSBO * (uint32_t)ks) >> 4), kIdescDw, ks > 0 ? 1u : 0u);
          commit(dbar);
        }
        __syncwarp();
      } else if (warp == 5 && a.discard_scratch && p >= 1 && k > 0) {   // this step's layer input a_k (slot k - 1) has been consumed by every warp
        const float4* dead = scrA + (size_t)(k - 1) * NC * TR;
#pragma unroll
        for (int i = 0; i < 4; ++i) asm volatile("discard.global.L2 [%0], 128;" :: "l"(reinterpret_cast<const char*>(dead) + (size_t)(32 * i + lane) * 128) : "memory");
      }
#if CLB_TC3_EARLY
      if (hasB && p < LT) wf.pass += 1u;
#else
      if (hasB && p >= 1) wf.pass += 1u;
#endif
      if (need_dx) wb.pass += 1u;
      // ---- A: bias gradient in the shadow of the tensor pipe ----
      bias_red16(dp, part32 + (size_t)k * PSLOT + BOFF + HW * hf, lane, 16);
#if !CLB_TC3_EARLY
      // ---- B: forward stage p ----
      if (hasB) fwd_stage(p, sB, rowB, inbB, scrB, h);
#endif
      // ---- A: delta-a of this layer -> delta-p of the next step ----
      if (need_dx) {
        mbar_wait(bar(B_CHAIN + (int)sA), (par >> sA) & 1u); par ^= (1u << sA);
        fence_after();
        uint32_t v[HW];
        CLB_TMEM_LD16(row_addr + col_d(sA) + col, v);
        wait_ld();
#pragma unroll
        for (int jj = 0; jj < HW; ++jj) { const float da = __uint_as_float(v[jj]); dp[jj] = ((mask >> jj) & 1u) ? da : kLeak * da; }
      }
      // ---- A: dW from tensor memory into the FP32 partial (frees the operand images and the accumulator) ----
      {
        mbar_wait(bar(B_DW), (par >> 2) & 1u); par ^= 4u;
        fence_after();
        const int q = warp & 3;
        uint32_t v0[HW], v1[HW];
        CLB_TMEM_LD16(row_addr + kColDw + col, v0);
        CLB_TMEM_LD16(row_addr + kColDw + 32 + col, v1);
        wait_ld();
        if (lane < 16) {
          const int i = (16 * q + lane) & 31;
          float4* dst = reinterpret_cast<float4*>(part32 + (size_t)k * PSLOT) + (4 * hf) * 32 + i;      // dw_slot32(i, 16 hf + 4 qq) / 4
#pragma unroll
          for (int qq = 0; qq < 4; ++qq)
            atomicAdd(dst + qq * 32, make_float4(__uint_as_float(v0[4 * qq]) + __uint_as_float(v1[4 * qq]),
                                                 __uint_as_float(v0[4 * qq + 1]) + __uint_as_float(v1[4 * qq + 1]),
                                                 __uint_as_float(v0[4 * qq + 2]) + __uint_as_float(v1[4 * qq + 2]),
                                                 __uint_as_float(v0[4 * qq + 3]) + __uint_as_float(v1[4 * qq + 3])));
        }
      }
    }
    // B becomes A (its a_LT is in h)
    rowA = rowB; inbA = inbB; reflA = reflB;
  }
  // ---- flush: the log-likelihood sum ----
  __syncthreads();
  ll_sum = warp_sum(ll_sum);
  if (lane == 0) red[warp] = ll_sum;
  fence_before();
  __syncthreads();
  if (tid == 0) {
    double t = 0.0;
    for (int i = 0; i < kThreads3 / 32; ++i) t += red[i];
    flush_ll(a.ll_part, a.acc, t);
  }
  if (warp == 0) { fence_after(); tmem_dealloc(tbase); }
}

}  // namespace clb
