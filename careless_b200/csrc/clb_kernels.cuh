// clb_kernels.cuh -- CUDA kernels of the ELBO gradient + Adam step (sm_100a).
//
// Step = k_refl_sample -> [k_dw_prior] -> [k_pack_images] -> k_obs_tc2 | k_obs | k_obs_tc16 -> k_reduce_partials*
//        -> k_refl_backward -> (all-reduce) -> k_var_sumsq -> k_pack_scalars -> (all-reduce) -> k_finalize -> k_adam
// See DESIGN.md for the data layout and the roofline of each kernel.
#pragma once
#include "clb_math.cuh"
#include "clb_tc.cuh"

#ifndef CLB_BWD_ORDER
// How the warps of k_obs_tc2 hand a pass over to the tensor pipe (all three are parity-green; B200 timings of the 10 M-observation
// step in profiles/README.md):  2 = __syncthreads() between the operand stores and the issue, warps 0 / 7 issue (16.8 ms, default);
// 0 = barrier-free: every warp arrives on a shared-memory counter, the LAST one issues chain + dW (17.8 ms);
// 1 = barrier-free, chain handed over before the dW operand images, dW collected one layer later (17.5 ms).
#define CLB_BWD_ORDER 2
#endif

namespace clb {

constexpr int kMaxLayers = 48;       // MLP layers incl. the Dense(2) head
constexpr int kObsThreads = 256;     // observations per CTA tile (one per thread)
constexpr int kMaxVars = 2 * kMaxLayers + 8;

// scalar accumulators (double) of one step
// Optional per-phase cycle accounting of k_obs (tools/phase_times.py builds with -DCLB_PHASE_TIMING): thread 0 of every
// CTA adds the cycles since its previous mark to g_phase[i].  Compiled out of the product library.
#ifdef CLB_PHASE_TIMING
__device__ unsigned long long g_phase[32];
__shared__ long long s_phase_last;
#define CLB_PH(i) do { if (threadIdx.x == 0) { const long long t_ = clock64(); atomicAdd(&g_phase[i], (unsigned long long)(t_ - s_phase_last)); s_phase_last = t_; } } while (0)
#define CLB_PH_START() do { if (threadIdx.x == 0) s_phase_last = clock64(); } while (0)
#else
#define CLB_PH(i) do { } while (0)
#define CLB_PH_START() do { } while (0)
#endif

enum { ACC_LOGQ_MINUS_LOGP = 0, ACC_LL = 1, ACC_SUMSQ_RAW = 2, ACC_SUMSQ_FILT = 3, ACC_NONFINITE = 4, ACC_COUNT = 8 };

// ---------------------------------------------------------------------------------------
// Per-reflection forward: reparameterised truncated-normal draw, log q, Wilson log p.
// One thread per (sample, reflection).  Initialises gz with the KL part of dL/dz.
// (variational.py:154, :123-139; surrogate_posteriors.py:50-53; wilson.py:50-57)
// ---------------------------------------------------------------------------------------
struct ReflArgs {
  const float* v_loc; const float* v_scale;       // (R)
  const uint8_t* centric; const float* eps_sigma; // (R): centric flag, epsilon*Sigma
  const int32_t* dw_parent;                       // (R) or null; -2 root, -1 absent, >=0 parent
  const uint32_t* refl_index;                     // (R) global index for the RNG
  const float* inj_u;                             // (S,R) or null
  float* z; float* gz;                            // (S,R)
  // (S,R) or null (forward-only calls): what the backward kernel needs of this draw, so that it does not have to repeat the
  // truncated-normal arithmetic -- {mu dz/dmu, (sigma - eps) dz/dsigma, cq mu dlogq/dmu, cq (sigma - eps) dlogq/dsigma}
  float4* bwd_coef;
  double* acc;
  double* kl_part;                                // deterministic mode: one partial per block (summed in a fixed order), else null
  int64_t R; int S; float eps; float cq;          // cq: KL coefficient per element
  uint64_t seed; uint32_t step;
};

// Four consecutive reflections per thread: v_loc / v_scale / eps_sigma / injected draws are read and z / gz written as one
// 16-byte access each (4 bytes for the centric flags), so a warp moves 512 contiguous bytes per instruction.  `vec` = the
// arrays are 16-byte aligned (R % 4 == 0 and aligned bases); otherwise the same code falls back to element accesses.
struct Vec4 { float v[4]; };
__device__ __forceinline__ Vec4 ld4(const float* p, int64_t i, int64_t n, bool vec, float fill = 0.f) {
  Vec4 o;
  if (vec) { const float4 t = *reinterpret_cast<const float4*>(p + i); o.v[0] = t.x; o.v[1] = t.y; o.v[2] = t.z; o.v[3] = t.w; }
  else {
#pragma unroll
    for (int j = 0; j < 4; ++j) o.v[j] = (i + j < n) ? p[i + j] : fill;
  }
  return o;
}
__device__ __forceinline__ void st4(float* p, int64_t i, int64_t n, bool vec, const Vec4& x) {
  if (vec) *reinterpret_cast<float4*>(p + i) = make_float4(x.v[0], x.v[1], x.v[2], x.v[3]);
  else {
#pragma unroll
    for (int j = 0; j < 4; ++j) if (i + j < n) p[i + j] = x.v[j];
  }
}

__global__ void __launch_bounds__(256) k_refl_sample(ReflArgs a, int vec) {
  const int64_t nq = (a.R + 3) >> 2;                    // groups of four reflections
  const int64_t gidx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double kl = 0.0;
  if (gidx < nq * a.S) {
    const int64_t r0 = (gidx % nq) << 2;
    const int s = (int)(gidx / nq);
    const bool v = vec != 0;
    const Vec4 vl = ld4(a.v_loc, r0, a.R, v), vs = ld4(a.v_scale, r0, a.R, v), es = ld4(a.eps_sigma, r0, a.R, v, 1.f);
    uint8_t cen[4] = {0, 0, 0, 0};
    if (v) { const uchar4 c4 = *reinterpret_cast<const uchar4*>(a.centric + r0); cen[0] = c4.x; cen[1] = c4.y; cen[2] = c4.z; cen[3] = c4.w; }
    else {
#pragma unroll
      for (int j = 0; j < 4; ++j) if (r0 + j < a.R) cen[j] = a.centric[r0 + j];
    }
    Vec4 u;
    if (a.inj_u) u = ld4(a.inj_u + (size_t)s * a.R, r0, a.R, v, 0.5f);
    else {
#pragma unroll
      for (int j = 0; j < 4; ++j) u.v[j] = (r0 + j < a.R) ? refl_uniform(a.seed, a.step, (uint32_t)s, a.refl_index[r0 + j]) : 0.5f;
    }
    Vec4 z, gz;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const bool live = r0 + j < a.R;
      const bool centric = cen[j] != 0;
      const TnSample t = tn_forward(vl.v[j], vs.v[j], centric ? 0.0f : 1e-32f /* manager.py:434 */, a.eps, u.v[j]);
      float g = t.dlogq_dz, term = t.logq;
      const bool wilson = (a.dw_parent == nullptr) || (live && a.dw_parent[r0 + j] == -2);
      if (wilson) {
        float lp, dlp;
        wilson_logp(t.z, centric, es.v[j], lp, dlp);
        term -= lp; g -= dlp;
      }
      z.v[j] = t.z; gz.v[j] = a.cq * g;
      if (live) kl += (double)term;
      if (a.bwd_coef != nullptr && live) {        // chain rule to the raw variables: d mu / d v_loc = mu, d sigma / d v_scale = sigma - eps
        const float dm = t.mu, ds = t.sigma - a.eps;
        a.bwd_coef[(size_t)s * a.R + r0 + j] = make_float4(dm * t.dz_dmu, ds * t.dz_dsigma, a.cq * dm * t.dlogq_dmu, a.cq * ds * t.dlogq_dsigma);
      }
    }
    st4(a.z + (size_t)s * a.R, r0, a.R, v, z);
    st4(a.gz + (size_t)s * a.R, r0, a.R, v, gz);
  }
  kl = warp_sum(kl);
  __shared__ double sm[8];
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = kl;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += sm[i];
    if (a.kl_part != nullptr) a.kl_part[blockIdx.x] = t; else atomicAdd(&a.acc[ACC_LOGQ_MINUS_LOGP], t);
  }
}

// DoubleWilson conditional prior for non-root entries (wilson.py:146-175): needs every z.
struct DwArgs {
  const float* z; float* gz;                      // (S,R)
  const uint8_t* centric; const float* eps_sigma; const int32_t* dw_parent; const int32_t* asu_id;
  const float* r_const;                           // (n_asu) fixed r, or
  const float* r_logit; float* g_r_logit;         // trainable logits + their gradient (optimize_r)
  double* acc;
  int64_t R; int S; float cq;
};

__global__ void __launch_bounds__(256) k_dw_prior(DwArgs a) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double kl = 0.0;
  if (idx < a.R * a.S) {
    const int64_t r = idx % a.R;
    const int64_t s = idx / a.R;
    const int parent = a.dw_parent[r];
    if (parent != -2) {
      const bool centric = a.centric[r] != 0;
      const int asu = a.asu_id[r];
      const float rr = a.r_logit ? sigmoidf(a.r_logit[asu]) : a.r_const[asu];
      const float zp = parent >= 0 ? a.z[s * a.R + parent] : 0.0f;
      const float loc = zp * rr;
      const float es = a.eps_sigma[r];
      const float s2 = (centric ? 1.0f : 0.5f) * es * (1.0f - rr * rr);
      float lp, dz, dloc, ds2;
      dw_child_logp(a.z[idx], loc, s2, centric, lp, dz, dloc, ds2);
      kl = -(double)lp;
      atomicAdd(&a.gz[idx], -a.cq * dz);
      if (parent >= 0) atomicAdd(&a.gz[s * a.R + parent], -a.cq * dloc * rr);
      if (a.r_logit) {
        const float ds2_dr = -(centric ? 2.0f : 1.0f) * es * rr;
        const float dlp_dr = dloc * zp + ds2 * ds2_dr;
        atomicAdd(&a.g_r_logit[asu], -a.cq * dlp_dr * rr * (1.0f - rr));
      }
    }
  }
  kl = warp_sum(kl);
  __shared__ double sm[8];
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = kl;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += sm[i];
    atomicAdd(&a.acc[ACC_LOGQ_MINUS_LOGP], t);
  }
}

// ---------------------------------------------------------------------------------------
// Observation kernel: scale MLP forward, scale sample, gather of z_f, likelihood, backward
// through everything, segmented reduction of dL/dz_f, per-CTA weight-gradient partials.
// (variational.py:156-171; scaling/nn.py:92-120; scaling/image.py:40-63; likelihoods/*.py)
// ---------------------------------------------------------------------------------------
struct MlpLayout {                 // flat (keras-order) parameter layout of the scale MLP
  int n_layers;                    // L + 1 (last = Dense(2) head)
  int in_dim[kMaxLayers], out_dim[kMaxLayers];
  int koff[kMaxLayers], boff[kMaxLayers];   // offsets of kernel / bias inside the MLP group
  int n_params;
};

struct ObsArgs {
  // rows (device layout: sorted + padded, SoA)
  const int32_t* refl; const int32_t* image; const int32_t* spot; const uint32_t* oidx;
  const float* meta;               // [d][Npad]
  const float* iobs; const float* sig;
  int64_t n_rows;                  // Npad
  int64_t n_rows_total;            // stride of injected eps (global N)
  int d;
  // model
  const float* theta_mlp; const float* theta_img;   // image scales (n_img-1) or null
  const float* wpack;              // [L][32][32] zero-padded FP32 copy of the hidden-layer kernels (TC kernels)
  const float* wimg;               // [L][fwd, bwd][hi, lo][kImgBytes] ready-made B operand images (k_obs_tc2, fetched by TMA)
  // image layers (scaling/image.py:66-125): per layer a kernel [n_img][W][W] (out, in) and a bias [n_img][W]
  int n_img_layers; int il_width; int il_n_images;
  const float* theta_il; float* g_il;   // parameter group and its gradient (null: frozen / eval)
  MlpLayout lay;
  const float* z; float* gz; int64_t R; int S;
  const float* inj_eps;            // (S, N_total) or null
  float* g_img;                    // gradient of image scales or null
  double* partials;                // [grid*KS][n_params], FP64 so that the cross-tile accumulation adds no FP32 rounding
  float* partials32;               // k_obs_tc2: the same memory as [grid][n_layers][32*32 + 32] FP32, accumulated with vector REDs
  float4* scratch;                 // [grid][L][WP/4][T]
  float* ipred_out;                // (S, N_total) original order, or null
  float* scale_mean_out; float* scale_std_out;   // (N_total) moments of the scale distribution, original order, or null
  double* acc;
  LikConst lik; float cl;          // likelihood coefficient (1/S or 1/(S*N))
  // Ev11 error model (likelihoods/mono.py:39-73): raw (pre-softplus) Sdfac, Sdadd, SdB or null; gradient or null
  const float* theta_lik; float* g_lik;
  int bijector; float shift; float eps;
  uint64_t seed; uint32_t step;
  int laue; int train_mlp;
  int n_partials;                  // tensor-core kernels: number of FP32 partial buffers the CTAs share (blockIdx % n_partials)
  // Deterministic mode (clb_config.deterministic): no floating-point atomics whose order could differ between runs.
  //  * dzf_rows (S, n_rows): every row's dL/dz_f is written out and k_gz_reduce sums each reflection's rows in a fixed order;
  //  * ll_part [grid]: per-CTA log-likelihood sums, added up in a fixed order by k_pack_scalars;
  //  * det: the CTA's partial buffer is exclusive (n_partials == grid) and every address receives REDs from ONE thread only
  //    (separate slots for the a_hi / a_lo rows of a kernel gradient and for every warp's bias sums), in program order.
  float* dzf_rows; double* ll_part; int det;
  int discard_scratch;             // 1: drop the activation scratch lines from L2 once the backward pass has consumed them (discard.global.L2)
  int bias_feat15;                 // k_obs_tc16 with max(metadata columns, width) <= 15: feature 15 of every layer input is padding, so the dW
                                   // operand image carries a constant 1 there and row 15 of the dW product IS the bias gradient (no shuffles)
};

// The activation scratch is written in the forward pass and read exactly once in the backward pass.  Without help every
// line is written back to DRAM when it is evicted (2.5 KB per observation and step); `discard.global.L2` tells the L2 that
// a line's contents are dead, so a line that is still resident when its reader is done never costs DRAM bandwidth.
// One lane per 128-byte line (8 consecutive rows x 16 B) issues it AFTER the warp has consumed the loaded registers.
__device__ __forceinline__ void discard_line(const void* p) {
  asm volatile("discard.global.L2 [%0], 128;" :: "l"(p) : "memory");
}

template <int WP> struct ObsSmem {
  static constexpr int T = kObsThreads;
  static constexpr int HS = T + 4;                    // padded stride of the transposed activation tile
  static size_t bytes(int n_layers, bool tensor_cores = false, int n_img_layers = 0) {
    return bytes_base(n_layers, tensor_cores) + sizeof(float) * (size_t)n_img_layers * (WP * WP + WP);
  }
  static size_t bytes_base(int n_layers, bool tensor_cores) {
    if (tensor_cores)     // dW operand images, compact head weights, biases, bias sums, reductions, chain images (128-thread CTA)
      return 2 * (size_t)tc::kDwImgBytes + sizeof(float) * (2 * WP + 2 * (size_t)n_layers * WP + (size_t)(tc::kThreads / 32) * WP)
             + 64 * sizeof(double) + 2 * (size_t)tc::kImgBytes + 128;
    return sizeof(float) * ((size_t)(WP == 64 ? 1 : n_layers) * WP * WP + (size_t)n_layers * WP   // W (width 64: only the head; the rest is streamed), b
                            + (size_t)n_layers * WP                                // bias-grad accumulators
                            + (size_t)WP * HS + (size_t)T * WP      // staged activation / delta tiles
                            + (size_t)T * 16                          // K-split reduction buffer
                            + (size_t)(T / 32) * WP)                  // per-warp bias-gradient sums
           + 64 * sizeof(double);
  }
};


// Bias gradient, step 1: column sums of dp over the 32 observations of a warp.  A recursive-halving exchange
// leaves one lane per column with the warp's sum (WP-1 shuffles), stored to bias_part[warp][column].
template <int WP>
__device__ __forceinline__ void bias_partial(const float (&dp)[WP], float* bias_part, int tid) {
  constexpr int NH = (WP >= 32) ? 5 : (WP == 16) ? 4 : 3;      // halving steps (at most 5: 32 lanes)
  float v[WP];
#pragma unroll
  for (int j = 0; j < WP; ++j) v[j] = dp[j];
  const int lane = tid & 31;
#pragma unroll
  for (int st = 0; st < NH; ++st) {
    const int bit = 16 >> st, half = WP >> (st + 1);
    const bool up = (lane & bit) != 0;
#pragma unroll
    for (int j = 0; j < half; ++j) {
      const float send = up ? v[j] : v[j + half];
      const float recv = __shfl_xor_sync(0xffffffffu, send, bit);
      v[j] = (up ? v[j + half] : v[j]) + recv;
    }
  }
#pragma unroll
  for (int st = NH; st < 5; ++st) v[0] += __shfl_xor_sync(0xffffffffu, v[0], 16 >> st);
  if constexpr (WP == 64) {                                      // five halvings leave two columns per lane: 2 lane, 2 lane + 1
    bias_part[(tid >> 5) * WP + 2 * lane] = v[0];
    bias_part[(tid >> 5) * WP + 2 * lane + 1] = v[1];
  } else {
    constexpr int SH = 5 - NH;                                   // lanes sharing a column
    if ((lane & ((1 << SH) - 1)) == 0) bias_part[(tid >> 5) * WP + (lane >> SH)] = v[0];
  }
}

// One layer's weight gradient over the CTA tile: dW[i][j] = sum_obs a[obs][i] * dp[obs][j].
// a is staged transposed (S_h[i][obs], padded stride), dp row-major with XOR-swizzled float4 chunks.
// Each thread accumulates a 4x4 patch of dW (rows pi + (WP/4) r, float4 column chunk pj) over its share
// of the tile's observations (K-split over KS4 thread groups): 8 LDS.128 feed 64 FFMA.  The K-split
// partial sums are combined through shared memory (Rbuf) and the first WP*WP/4 threads add the result to
// the CTA's FP64 running sum in its L2-resident partial buffer (`part`, 4 consecutive doubles each;
// exclusive ownership, no atomics).  The partial is fetched BEFORE the barriers so the L2 round trip
// overlaps the loop.
template <int WP>
__device__ __forceinline__ void stage_and_accumulate(const float (&ain)[WP], const float (&dp)[WP],
                                                     float* S_h, float4* S_d, float4* Rbuf, float* bias_part,
                                                     float* dbacc_k, double* part, int tid,
                                                     float* il_gk = nullptr, float* il_gb = nullptr, int il_w = 0) {
  constexpr int T = kObsThreads, HS = T + 4, NC = WP / 4;
  constexpr int Q = WP / 4;               // patch rows are strided by Q, patch columns are one float4 chunk
  constexpr int TPL4 = Q * Q;             // threads covering one WPxWP matrix with 4x4 patches
  constexpr int KS4 = T / TPL4;           // K (observation) split
  constexpr int OBS = T / KS4;            // observations per K-split group
  constexpr int NOUT4 = WP * WP / 4;      // float4 outputs of one layer
  constexpr int NOWN = (NOUT4 + T - 1) / T;   // float4 outputs owned by one thread (1 up to width 32, 4 at width 64): tid, tid + T, ...
  static_assert(OBS % 4 == 0 && OBS >= 4, "tile too small for the K split");
  const bool owner = tid < NOUT4;
  const bool to_image = il_w > 0;         // image layer: the tile's gradient goes to that image's slot (atomics)
  double2 p01[NOWN], p23[NOWN];
#pragma unroll
  for (int m = 0; m < NOWN; ++m) {
    p01[m] = p23[m] = make_double2(0.0, 0.0);
    if (owner && !to_image) {
      p01[m] = __ldcg(reinterpret_cast<const double2*>(part + (size_t)m * T * 4));
      p23[m] = __ldcg(reinterpret_cast<const double2*>(part + (size_t)m * T * 4) + 1);
    }
  }
  __syncthreads();                        // previous consumers of the staging / reduction buffers are done
#pragma unroll
  for (int i = 0; i < WP; ++i) S_h[i * HS + tid] = ain[i];
#pragma unroll
  for (int c = 0; c < NC; ++c)
    S_d[tid * NC + (c ^ (tid & (NC - 1)))] = make_float4(dp[4 * c], dp[4 * c + 1], dp[4 * c + 2], dp[4 * c + 3]);
  __syncthreads();
  const int pp = tid % TPL4, ks = tid / TPL4;
  const int pi = pp % Q, pj = pp / Q;
  float acc[4][4];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;
  const float* hrow = S_h + pi * HS + ks * OBS;
  const float4* drow = S_d + (size_t)(ks * OBS) * NC;
  const int o0 = ks * OBS;
  float4 hb[2][4], db[2][4];
  auto load = [&](int buf, int o) {       // o: offset inside this group's observation range
#pragma unroll
    for (int r = 0; r < 4; ++r) hb[buf][r] = *reinterpret_cast<const float4*>(hrow + r * Q * HS + o);
#pragma unroll
    for (int q = 0; q < 4; ++q) db[buf][q] = drow[(o + q) * NC + (pj ^ ((o0 + o + q) & (NC - 1)))];
  };
  load(0, 0);
#pragma unroll
  for (int st = 0; st < OBS / 4; ++st) {
    const int cur = st & 1;
    if (st + 1 < OBS / 4) load(cur ^ 1, 4 * (st + 1));
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const float hv[4] = {hb[cur][r].x, hb[cur][r].y, hb[cur][r].z, hb[cur][r].w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        acc[r][0] = fmaf(hv[q], db[cur][q].x, acc[r][0]); acc[r][1] = fmaf(hv[q], db[cur][q].y, acc[r][1]);
        acc[r][2] = fmaf(hv[q], db[cur][q].z, acc[r][2]); acc[r][3] = fmaf(hv[q], db[cur][q].w, acc[r][3]);
      }
    }
  }
  // K-split partial sums -> Rbuf[ks][r][pp]
#pragma unroll
  for (int r = 0; r < 4; ++r) Rbuf[(ks * 4 + r) * TPL4 + pp] = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
  bias_partial<WP>(dp, bias_part, tid);
  __syncthreads();
  if (tid < WP) {
    float sum = 0.f;
#pragma unroll
    for (int w = 0; w < T / 32; ++w) sum += bias_part[w * WP + tid];
    if (!to_image) dbacc_k[tid] += sum;     // column tid of this layer is owned by thread tid: no atomics
    else if (il_gb != nullptr && tid < il_w) atomicAdd(&il_gb[tid], sum);
  }
  if (owner) {
#pragma unroll
    for (int m = 0; m < NOWN; ++m) {
      const int o4 = tid + m * T;           // this output: dK[i = pi + Q r][j = 4 pj ..] with o4 = r*Q*Q + pj*Q + pi
      float4 t = Rbuf[o4];
#pragma unroll
      for (int k2 = 1; k2 < KS4; ++k2) {
        const float4 v = Rbuf[k2 * NOUT4 + o4];
        t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
      }
      if (!to_image) {
        double* pm = part + (size_t)m * T * 4;
        __stcg(reinterpret_cast<double2*>(pm), make_double2(p01[m].x + (double)t.x, p01[m].y + (double)t.y));
        __stcg(reinterpret_cast<double2*>(pm) + 1, make_double2(p23[m].x + (double)t.z, p23[m].y + (double)t.w));
      } else if (il_gk != nullptr) {        // image kernels are stored (out, in)
        const int r4 = o4 / TPL4, pj4 = (o4 % TPL4) / Q, pi4 = o4 % Q;
        const int i = pi4 + Q * r4;
        const float tv[4] = {t.x, t.y, t.z, t.w};
        if (i < il_w) {
#pragma unroll
          for (int c = 0; c < 4; ++c) { const int j = 4 * pj4 + c; if (j < il_w) atomicAdd(&il_gk[j * il_w + i], tv[c]); }
        }
      }
    }
  }
}

// Tensor-core version of one layer's backward (TC kernels: WP == 32, 128-thread CTAs): dW_k = a_k^T dp_k and,
// when need_dx, dp <- delta a_k = dp_k W_k^T, both on tcgen05 (clb_tc.cuh); the FP32 pipe only splits operands,
// folds the accumulator copies and adds the result to the CTA's FP64 partial (same layout as the FP32 path:
// float4 output o4 = r*64 + pj*8 + pi holds element (i = pi + 8 r, j = 4 pj ..); thread tid owns o4 = tid, tid+128).
__device__ __forceinline__ void tc_layer_backward(tc::Ctx& tcx, float (&dp)[32], const float (&ain)[32], const float (&w)[8],
                                                  bool need_dx, float* bias_part, float* dbacc_k, double* part, int tid,
                                                  float* il_gk = nullptr, float* il_gb = nullptr, int il_w = 0) {
  const bool to_image = il_w > 0;          // image layer: the tile's gradient goes to that image's slot (atomics)
  double2 pr[2][2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    pr[h][0] = pr[h][1] = make_double2(0.0, 0.0);
    if (!to_image) {
      pr[h][0] = __ldcg(reinterpret_cast<const double2*>(part + 512 * h));
      pr[h][1] = __ldcg(reinterpret_cast<const double2*>(part + 512 * h) + 1);
    }
  }
  CLB_PH(5);
  bias_partial<32>(dp, bias_part, tid);
  CLB_PH(6);
  tc::issue_backward(tcx, dp, ain, w, need_dx);
  CLB_PH(7);
  if (need_dx) tc::collect(tcx, dp);
  CLB_PH(8);
  tc::collect_dw(tcx);
  CLB_PH(9);
  __syncthreads();
  CLB_PH(10);
  {
    const int pj = (tid & 63) >> 3, pi = tid & 7;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int r = (tid >> 6) + 2 * h;
      const float* st = reinterpret_cast<const float*>(tcx.dw_a) + (size_t)(pi + 8 * r) * tc::kStageStride + 4 * pj;
      const float4 t0 = *reinterpret_cast<const float4*>(st);
      const float4 t1 = *reinterpret_cast<const float4*>(st + (size_t)32 * tc::kStageStride);
      if (!to_image) {
        __stcg(reinterpret_cast<double2*>(part + 512 * h), make_double2(pr[h][0].x + (double)(t0.x + t1.x), pr[h][0].y + (double)(t0.y + t1.y)));
        __stcg(reinterpret_cast<double2*>(part + 512 * h) + 1, make_double2(pr[h][1].x + (double)(t0.z + t1.z), pr[h][1].y + (double)(t0.w + t1.w)));
      } else if (il_gk != nullptr) {
        const int i = pi + 8 * r;
        const float tv[4] = {t0.x + t1.x, t0.y + t1.y, t0.z + t1.z, t0.w + t1.w};
        if (i < il_w) {
#pragma unroll
          for (int c = 0; c < 4; ++c) { const int j = 4 * pj + c; if (j < il_w) atomicAdd(&il_gk[j * il_w + i], tv[c]); }
        }
      }
    }
    if (tid < 32) {
      float sum = 0.f;
#pragma unroll
      for (int w2 = 0; w2 < tc::kThreads / 32; ++w2) sum += bias_part[w2 * 32 + tid];
      if (!to_image) dbacc_k[tid] += sum;
      else if (il_gb != nullptr && tid < il_w) atomicAdd(&il_gb[tid], sum);
    }
  }
  CLB_PH(11);
  __syncthreads();      // the stage aliases the operand image of the next layer
  CLB_PH(12);
}

// Padded per-CTA partial layout: [NL][WP*WP] kernel sums in patch order -- element (i, j) of layer k at
// k*WP*WP + partial_elem(i, j, WP) -- followed by [NL][WP] bias sums.
__host__ __device__ inline int partial_elem(int i, int j, int WP) {
  const int Q = WP / 4;
  return (((i / Q) * Q * Q + (j >> 2) * Q + (i % Q)) << 2) + (j & 3);
}
__host__ __device__ inline int partial_row_size(int n_layers, int WP) { return n_layers * (WP * WP + WP); }

// Everything between the scale network's two outputs and their gradients for one observation row: the scale sample,
// the gather of z_f, the (Laue: per-spot) likelihood, dL/dz_f reduced over runs of equal refl_id with one atomic per
// run, image-scale and error-model gradients.  All 32 lanes of the warp must call it (warp-level scans inside).
template <int LIK>
__device__ __forceinline__ void obs_epilogue(const ObsArgs& a, int64_t row, bool inb, bool active, int refl, int lane,
                                             float out0, float out1, float ev_f, float ev_a, float ev_b,
                                             double& ll_sum, float& dmu, float& drho) {
  float sig_s, dsig;
  if (a.bijector == 0) { dsig = expf(out1); sig_s = dsig + a.eps; }
  else { sig_s = softplusf(out1) + a.eps; dsig = sigmoidf(out1); }
  // (the observation rows are read once per step: streaming loads, so that they do not displace the activation scratch from the L2)
  const int img = (a.image != nullptr && inb) ? __ldcs(&a.image[row]) : 0;
  const float aimg = (a.theta_img != nullptr && img > 0) ? a.theta_img[img - 1] : 1.0f;
  const uint32_t oi = inb ? __ldcs(&a.oidx[row]) : 0u;
  const float iobs = inb ? __ldcs(&a.iobs[row]) : 0.f;
  const float sg = inb ? __ldcs(&a.sig[row]) : 1.f;
  if (a.scale_mean_out != nullptr && active) {     // variational.py:67-69: scale_dist.mean() / .stddev()
    a.scale_mean_out[oi] = aimg * (out0 + a.shift);
    a.scale_std_out[oi] = fabsf(aimg) * sig_s;
  }
  // runs of equal keys inside the warp (same for every MC sample)
  const WarpRuns refl_runs = warp_runs(active ? refl : -1 - lane, lane);
  WarpRuns spot_runs = refl_runs, img_runs = refl_runs;
  if (a.laue) spot_runs = warp_runs(inb ? __ldcs(&a.spot[row]) : -1, lane);   // padding rows carry spot -1
  const bool img_live = active && img > 0;
  if (a.g_img != nullptr) img_runs = warp_runs(img_live ? img : -1 - lane, lane);
  float d_aimg = 0.f;
  dmu = 0.f; drho = 0.f;
  float ev_gf = 0.f, ev_ga = 0.f, ev_gb = 0.f;
  for (int s = 0; s < a.S; ++s) {
    float e = 0.f;
    if (active) e = a.inj_eps ? a.inj_eps[(size_t)s * a.n_rows_total + oi] : obs_normal(a.seed, a.step, (uint32_t)s, oi);
    const float base = fmaf(sig_s, e, out0) + a.shift;
    const float zs = aimg * base;
    const float zf = active ? __ldg(&a.z[(size_t)s * a.R + refl]) : 0.f;
    const float ip = zs * zf * zf;
    if (a.ipred_out != nullptr && active) a.ipred_out[(size_t)s * a.n_rows_total + oi] = ip;
    float x = ip;
    bool eval = active;
    if (a.laue) {           // harmonic segment-sum within the warp (spots never straddle a warp)
      x = warp_segtotal(active ? ip : 0.f, spot_runs, lane);
      eval = active && spot_runs.tail;   // count each spot once
    }
    float ll = 0.f, g = 0.f;
    if (a.theta_lik == nullptr) {
      if (active) lik_eval<LIK>(x, iobs, sg, a.lik, ll, g);
    } else if (active) {
      float gf, ga, gb;
      ev11_eval<LIK>(x, iobs, sg, ev_f, ev_a, ev_b, a.lik, ll, g, gf, ga, gb);
      if (eval) { ev_gf += gf; ev_ga += ga; ev_gb += gb; }
    }
    if (eval) ll_sum += (double)ll;
    const float G = active ? a.cl * g : 0.f;
    const float d_zs = G * zf * zf;
    const float d_zf = G * zs * 2.0f * zf;
    // segmented reduction of dL/dz_f over runs of equal refl_id, one atomic per run
    if (a.dzf_rows != nullptr) {       // deterministic mode: k_gz_reduce adds the rows of a reflection in a fixed order
      if (active) a.dzf_rows[(size_t)s * a.n_rows + row] = d_zf;
    } else {
      const float tot = warp_segsum(d_zf, refl_runs, lane);
      if (active && refl_runs.tail) atomicAdd(&a.gz[(size_t)s * a.R + refl], tot);
    }
    const float d_base = aimg * d_zs;
    d_aimg += base * d_zs;
    dmu += d_base;
    drho += d_base * e * dsig;
  }
  if (a.g_img != nullptr) {
    const float tot = warp_segsum(d_aimg, img_runs, lane);
    if (img_live && img_runs.tail) atomicAdd(&a.g_img[img - 1], tot);
  }
  if (a.g_lik != nullptr) {      // d loss / d raw error-model parameters: one atomic per warp and parameter
    ev_gf = warp_sum(ev_gf); ev_ga = warp_sum(ev_ga); ev_gb = warp_sum(ev_gb);
    if (lane == 0) {
      atomicAdd(&a.g_lik[0], a.cl * ev_gf * sigmoidf(a.theta_lik[0]));
      atomicAdd(&a.g_lik[1], a.cl * ev_ga * sigmoidf(a.theta_lik[1]));
      atomicAdd(&a.g_lik[2], a.cl * ev_gb * sigmoidf(a.theta_lik[2]));
    }
  }
}

__device__ __forceinline__ void flush_ll(double* ll_part, double* acc, double t) {     // (pointers by value: a reference to the kernel's parameter struct would force a local copy)
  if (ll_part != nullptr) ll_part[blockIdx.x] = t; else atomicAdd(&acc[ACC_LL], t);
}

// dL/dz_f of every reflection from the per-row values, in the fixed order of the host-built CSR (deterministic mode).
__global__ void __launch_bounds__(256) k_gz_reduce(const float* dzf_rows, const int32_t* refl_ptr, const int32_t* refl_rows,
                                                   float* gz, int64_t R, int S, int64_t n_rows) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= R * S) return;
  const int64_t r = idx % R, s = idx / R;
  float sum = 0.f;
  for (int32_t i = refl_ptr[r]; i < refl_ptr[r + 1]; ++i) sum += dzf_rows[(size_t)s * n_rows + refl_rows[i]];
  gz[idx] += sum;
}

// Fixed-order sum of per-block partials (deterministic mode): out[q] += sum_b part[q * stride + b].
__global__ void k_sum_partials(const double* part, int n_blocks, int stride, int n_quantities, double* out) {
  const int q = threadIdx.x;
  if (q >= n_quantities) return;
  double t = 0.0;
  for (int b = 0; b < n_blocks; ++b) t += part[(size_t)q * stride + b];
  out[q] += t;
}

// TC = true (WP == 32 only): the forward and dX products of the hidden layers run on the tensor cores
// (tcgen05.mma kind::tf32, 3xTF32 error-compensated, operands/accumulators in tensor memory; clb_tc.cuh);
// TC = false: everything on the FP32 FMA pipe.
template <int WP, int LIK, bool TC>
__global__ void __launch_bounds__(TC ? tc::kThreads : kObsThreads, TC ? 2 : 1) k_obs(ObsArgs a) {
  static_assert(!TC || WP == 32, "the tensor-core path is built for the padded width 32");
  constexpr int T = TC ? tc::kThreads : kObsThreads, HS = T + 4, NC = WP / 4;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const int NL = a.lay.n_layers;          // incl. head
  const int L = NL - 1;                   // hidden layers
  constexpr int HSTR = TC ? 2 : WP;       // row stride of the head weights (compact [WP][2] in the TC kernels)
  unsigned char* sp = smem_raw;
  char* tc_dwa = nullptr; char* tc_dwb = nullptr;
  if constexpr (TC) { tc_dwa = reinterpret_cast<char*>(sp); tc_dwb = tc_dwa + tc::kDwImgBytes; sp += 2 * tc::kDwImgBytes; }
  // width 64 (FP32 path): the hidden-layer kernels do not fit in shared memory (L x 16 KB); they are read from the zero-padded
  // packed copy in global memory instead (k_pack_weights; every lane of a warp reads the same address: L1 broadcast)
  constexpr bool STREAM_W = !TC && WP == 64;
  float* Wsm = reinterpret_cast<float*>(sp);                // [L][WP][WP] hidden layers (FP32 path only), then the head [WP][HSTR]
  float* Whead = Wsm + ((TC || STREAM_W) ? 0 : (size_t)L * WP * WP);
  float* bsm = Whead + (size_t)WP * HSTR;                   // [NL][WP]
  float* dbacc = bsm + (size_t)NL * WP;                     // [NL][WP]
  const int K = a.n_img_layers;           // image layers sit between the hidden layers and the head
  const int LT = L + K;                   // layers of the chain
  float* Wimg = dbacc + (size_t)NL * WP;                    // [K][WP][WP] this tile's image-layer kernels as [in][out]
  float* bimg = Wimg + (size_t)K * WP * WP;                 // [K][WP]
  float* nxtp = bimg + (size_t)K * WP;
  float* S_h = nullptr; float4* S_d = nullptr; float4* Rbuf = nullptr;
  if constexpr (!TC) {
    S_h = nxtp;                                             // [WP][HS]
    S_d = reinterpret_cast<float4*>(S_h + (size_t)WP * HS); // [T][NC]
    Rbuf = S_d + (size_t)T * NC;                            // [KS4][4][TPL4] float4 = T*16 floats
    nxtp = reinterpret_cast<float*>(Rbuf) + (size_t)T * 16;
  }
  float* bias_part = nxtp;                                  // [T/32][WP]
  double* red = reinterpret_cast<double*>(bias_part + (size_t)(T / 32) * WP);
  char* tc_img = reinterpret_cast<char*>(red + 64);         // [2][kImgBytes] B operand images (TC only)
  uint64_t* tc_bar = reinterpret_cast<uint64_t*>(tc_img + 2 * tc::kImgBytes);   // [0] chain, [1] dW
  uint32_t* tc_slot = reinterpret_cast<uint32_t*>(tc_bar + 2);

  const int tid = threadIdx.x, lane = tid & 31;
  tc::Ctx tcx{};
  if constexpr (TC) {
    if (tid == 0) { tc::mbar_init(tc::smem_u32(tc_bar), 1); tc::mbar_init(tc::smem_u32(tc_bar + 1), 1); }
    if (tid < 32) tc::tmem_alloc(tc::smem_u32(tc_slot));
    tc::fence_before();
  }
  // ---- stage the weights (zero padded to WP x WP) ----
  if constexpr (!TC && !STREAM_W) {
    for (int idx = tid; idx < L * WP * WP; idx += T) {
      const int k = idx / (WP * WP), i = (idx / WP) % WP, j = idx % WP;
      float w = 0.f;
      if (i < a.lay.in_dim[k] && j < a.lay.out_dim[k]) w = a.theta_mlp[a.lay.koff[k] + i * a.lay.out_dim[k] + j];
      Wsm[idx] = w;
    }
  }
  for (int idx = tid; idx < WP * HSTR; idx += T) {
    const int i = idx / HSTR, j = idx % HSTR;
    Whead[idx] = (i < a.lay.in_dim[L] && j < a.lay.out_dim[L]) ? a.theta_mlp[a.lay.koff[L] + i * a.lay.out_dim[L] + j] : 0.f;
  }
  for (int idx = tid; idx < NL * WP; idx += T) {
    const int k = idx / WP, j = idx % WP;
    bsm[idx] = (j < a.lay.out_dim[k]) ? a.theta_mlp[a.lay.boff[k] + j] : 0.f;
    dbacc[idx] = 0.f;
  }
  __syncthreads();
  if constexpr (TC) {
    tc::fence_after();
    const uint32_t tbase = *tc_slot;
    const int warp = tid >> 5;
    tcx.row_addr = tbase + ((uint32_t)(32 * warp) << 16);
    tcx.mbar = tc::smem_u32(tc_bar);
    tcx.parity = 0;
    tcx.img_hi = tc_img; tcx.img_lo = tc_img + tc::kImgBytes;
    tcx.desc_hi = tc::make_desc(tc::smem_u32(tcx.img_hi)); tcx.desc_lo = tc::make_desc(tc::smem_u32(tcx.img_lo));
    tcx.tid = tid;
    tcx.base = tbase;
    tcx.mbar_dw = tc::smem_u32(tc_bar + 1); tcx.parity_dw = 0;
    tcx.dw_a = tc_dwa; tcx.dw_b = tc_dwb;
    tcx.desc_dwa = tc::make_desc_mn(tc::smem_u32(tc_dwa)); tcx.desc_dwb = tc::make_desc_mn(tc::smem_u32(tc_dwb));
  }

  const int PP = partial_row_size(NL, WP);
  double* part_rows = a.partials + (size_t)blockIdx.x * PP + (size_t)tid * 4;   // valid for tid < WP*WP/4
  float4* scr = a.scratch + (size_t)blockIdx.x * LT * NC * T;
  double ll_sum = 0.0;
  float ev_f = 1.f, ev_a = 0.f, ev_b = 0.f;      // Ev11: Sdfac, Sdadd, SdB = softplus(raw)
  if (a.theta_lik != nullptr) { ev_f = softplusf(a.theta_lik[0]); ev_a = softplusf(a.theta_lik[1]); ev_b = softplusf(a.theta_lik[2]); }
  const int64_t n_tiles = (a.n_rows + T - 1) / T;

  CLB_PH_START();
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t row = tile * T + tid;
    const bool inb = row < a.n_rows;
    const int refl = inb ? a.refl[row] : -1;
    const bool active = refl >= 0;
    // image layers: the host prep never lets an image straddle a tile, so the whole tile uses one image's weights
    const int timg = (K > 0) ? a.image[tile * T] : 0;
    if (K > 0) {
      __syncthreads();                     // the previous tile is done with Wimg
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
    for (int i = 0; i < WP; ++i) h[i] = (inb && i < a.d) ? a.meta[(size_t)i * a.n_rows + row] : 0.f;
    float wreg[8];                       // TC: this thread's share of the next pass's weights
    auto wsrc = [&](int k) -> const float* {      // FP32 weights of chain layer k as [in][out] (TC: 32x32 padded)
      if (k >= L) return Wimg + (size_t)(k - L) * WP * WP;
      if constexpr (TC || STREAM_W) return a.wpack + (size_t)k * WP * WP; else return Wsm + (size_t)k * WP * WP;
    };
    if constexpr (TC) { if (LT > 0) tc::load_w<false>(wsrc(0), tid, wreg); }
    for (int k = 0; k < LT; ++k) {
      const float* Wk = wsrc(k);
      const float* bk = (k >= L) ? bimg + (size_t)(k - L) * WP : bsm + (size_t)k * WP;
      float o[WP];
      if constexpr (TC) {
        CLB_PH(0);
        tc::issue<false>(tcx, h, wreg);
        CLB_PH(1);
        // prefetch the next pass's weights while the tensor cores work: next forward layer, or the first backward layer
        if (k + 1 < LT) tc::load_w<false>(wsrc(k + 1), tid, wreg);
        else if (a.train_mlp && LT > 1) tc::load_w<true>(wsrc(LT - 1), tid, wreg);
        tc::collect(tcx, o);
        CLB_PH(2);
#pragma unroll
        for (int j = 0; j < WP; ++j) o[j] += bk[j];
      } else {
#pragma unroll
        for (int j = 0; j < WP; ++j) o[j] = bk[j];
#pragma unroll
        for (int i = 0; i < WP; ++i) {
#pragma unroll
          for (int c = 0; c < NC; ++c) {
            const float4 w = *reinterpret_cast<const float4*>(&Wk[i * WP + 4 * c]);
            o[4 * c] = fmaf(h[i], w.x, o[4 * c]); o[4 * c + 1] = fmaf(h[i], w.y, o[4 * c + 1]);
            o[4 * c + 2] = fmaf(h[i], w.z, o[4 * c + 2]); o[4 * c + 3] = fmaf(h[i], w.w, o[4 * c + 3]);
          }
        }
      }
#pragma unroll
      for (int j = 0; j < WP; ++j) h[j] = o[j] > 0.f ? o[j] : kLeak * o[j];
      if (a.train_mlp) {
#pragma unroll
        for (int c = 0; c < NC; ++c) scr[((size_t)k * NC + c) * T + tid] = make_float4(h[4 * c], h[4 * c + 1], h[4 * c + 2], h[4 * c + 3]);
      }
    }
    CLB_PH(3);
    float out0, out1;
    {
      out0 = bsm[L * WP]; out1 = bsm[L * WP + 1];
#pragma unroll
      for (int i = 0; i < WP; ++i) {
        const float2 w = *reinterpret_cast<const float2*>(&Whead[i * HSTR]);
        out0 = fmaf(h[i], w.x, out0); out1 = fmaf(h[i], w.y, out1);
      }
    }
    // ---------------- scale sample, gather, likelihood ----------------
    float dmu, drho;
    obs_epilogue<LIK>(a, row, inb, active, refl, lane, out0, out1, ev_f, ev_a, ev_b, ll_sum, dmu, drho);
    CLB_PH(4);
    if (!a.train_mlp) continue;
    // ---------------- backward through the MLP ----------------
    float dp[WP], nxt[WP];
    // prefetch the input activations of the last hidden layer while the head is processed
    auto load_act = [&](float (&dst)[WP], int k) {     // a_k: input of hidden layer k
      if (k > 0) {
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          const float4 v = __ldcg(&scr[((size_t)(k - 1) * NC + c) * T + tid]);
          dst[4 * c] = v.x; dst[4 * c + 1] = v.y; dst[4 * c + 2] = v.z; dst[4 * c + 3] = v.w;
        }
      } else {
#pragma unroll
        for (int i = 0; i < WP; ++i) dst[i] = (inb && i < a.d) ? a.meta[(size_t)i * a.n_rows + row] : 0.f;
      }
    };
    if (LT > 0) load_act(nxt, LT - 1);
#pragma unroll
    for (int j = 0; j < WP; ++j) dp[j] = 0.f;
    dp[0] = dmu; dp[1] = drho;
    // head: dW_out = a_L^T [dmu, drho]
    if constexpr (TC) tc_layer_backward(tcx, dp, h, wreg, false, bias_part, dbacc + L * WP, part_rows + (size_t)L * WP * WP, tid);
    else stage_and_accumulate<WP>(h, dp, S_h, S_d, Rbuf, bias_part, dbacc + L * WP, part_rows + (size_t)L * WP * WP, tid);
    using mask_t = typename std::conditional<(WP > 32), unsigned long long, unsigned>::type;
    mask_t mask = 0;                           // sign bits of a_{k+1}: leaky'(pre-activation)
#pragma unroll
    for (int j = 0; j < WP; ++j) mask |= (mask_t)(h[j] > 0.f ? 1u : 0u) << j;
#pragma unroll
    for (int i = 0; i < WP; ++i) {
      const float2 w = *reinterpret_cast<const float2*>(&Whead[i * HSTR]);
      dp[i] = w.x * dmu + w.y * drho;            // delta a_L
    }
    for (int k = LT - 1; k >= 0; --k) {
      // image layers send their gradient to the tile's image slot; hidden layers to the CTA partial
      const bool is_il = k >= L;
      float* il_gk = nullptr; float* il_gb = nullptr;
      if (is_il && a.g_il != nullptr) {
        const int w = a.il_width;
        const size_t lstride = (size_t)a.il_n_images * w * (w + 1);
        il_gk = a.g_il + (size_t)(k - L) * lstride + (size_t)timg * w * w;
        il_gb = a.g_il + (size_t)(k - L) * lstride + (size_t)a.il_n_images * w * w + (size_t)timg * w;
      }
      const int il_w = is_il ? a.il_width : 0;
      float* dbk = dbacc + (size_t)(is_il ? L : k) * WP;            // unused for image layers
      double* partk = part_rows + (size_t)(is_il ? L : k) * WP * WP;  // unused for image layers
      // delta p_k = delta a_{k+1} * leaky'(a_{k+1});  sign(a) == sign(pre-activation)
#pragma unroll
      for (int j = 0; j < WP; ++j) dp[j] = ((mask >> j) & 1u) ? dp[j] : kLeak * dp[j];
      float ain[WP];
      mask = 0;
#pragma unroll
      for (int i = 0; i < WP; ++i) { ain[i] = nxt[i]; mask |= (mask_t)(ain[i] > 0.f ? 1u : 0u) << i; }
      if (k > 0) load_act(nxt, k - 1);           // in flight during this layer's dW loop
      if constexpr (TC) {
        // delta a_k = delta p_k W_k^T and dW_k = a_k^T delta p_k, both on the tensor cores
        float wcur[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) wcur[q] = wreg[q];
        if (k > 1) tc::load_w<true>(wsrc(k - 1), tid, wreg);     // next backward layer
        tc_layer_backward(tcx, dp, ain, wcur, k > 0, bias_part, dbk, partk, tid, il_gk, il_gb, il_w);
        continue;
      }
      stage_and_accumulate<WP>(ain, dp, S_h, S_d, Rbuf, bias_part, dbk, partk, tid, il_gk, il_gb, il_w);
      if (k > 0) {
        const float* Wk = wsrc(k);
        float da[WP];
#pragma unroll
        for (int i = 0; i < WP; ++i) {
          float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
          for (int c = 0; c < NC; ++c) {
            const float4 w = *reinterpret_cast<const float4*>(&Wk[i * WP + 4 * c]);
            s0 = fmaf(w.x, dp[4 * c], s0); s1 = fmaf(w.y, dp[4 * c + 1], s1);
            s2 = fmaf(w.z, dp[4 * c + 2], s2); s3 = fmaf(w.w, dp[4 * c + 3], s3);
          }
          da[i] = (s0 + s1) + (s2 + s3);
        }
#pragma unroll
        for (int i = 0; i < WP; ++i) dp[i] = da[i];
      }
    }
  }
  // ---- flush: bias gradients and the log-likelihood sum ----
  __syncthreads();
  if (a.train_mlp) {
    for (int idx = tid; idx < NL * WP; idx += T)
      a.partials[(size_t)blockIdx.x * PP + (size_t)NL * WP * WP + idx] += (double)dbacc[idx];
  }
  ll_sum = warp_sum(ll_sum);
  if (lane == 0) red[tid >> 5] = ll_sum;
  if constexpr (TC) tc::fence_before();
  __syncthreads();
  if (tid == 0) {
    double t = 0.0;
    for (int i = 0; i < T / 32; ++i) t += red[i];
    flush_ll(a.ll_part, a.acc, t);
  }
  if constexpr (TC) {
    if (tid < 32) tc::tmem_dealloc(*tc_slot);
  }
}


// ---------------------------------------------------------------------------------------
// k_obs_tc2: the tensor-core observation kernel with TWO THREADS PER ROW (padded width 32).
// Same math, same shared / tensor-memory images and the same global layouts (scratch, FP64 partials) as
// k_obs<32, LIK, true>; the 128 rows of a tile are shared by 256 threads: thread (row = tid % 128, hf = tid / 128)
// carries features [16 hf, 16 hf + 16) of its row through the chain, so every per-row phase (operand split, tensor-
// memory and shared-memory image stores, bias / LeakyReLU, scratch traffic) is half as long per thread and an SM runs
// 16 warps (two CTAs) instead of 8.  The head and the likelihood epilogue run in the hf = 0 threads.
// ---------------------------------------------------------------------------------------
struct ObsSmem2 {
  static size_t bytes(int n_layers, int n_img_layers) {
    return 2 * (size_t)tc::kDwImgBytes + ((CLB_BIAS_ONES || CLB_BIAS_COL) ? (size_t)tc::kDwLBO : 0) + sizeof(float) * (64 + (size_t)n_layers * 32 + (size_t)n_img_layers * (1024 + 32))
           + 64 * sizeof(double) + 4 * (size_t)tc::kImgBytes + 64 + 4 * 128 * sizeof(float) + 128;
  }
};

// Column sums of the warp's 32 rows of a 16-column block (recursive halving, 16 shuffles); the even lanes then add
// their column (lane >> 1) to dst[col] with one RED each.
__device__ __forceinline__ void bias_red16(const float (&dp)[16], float* dst, int lane, int limit) {
  float v[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] = dp[j];
#pragma unroll
  for (int st = 0; st < 4; ++st) {
    const int bit = 16 >> st, half = 16 >> (st + 1);
    const bool up = (lane & bit) != 0;
#pragma unroll
    for (int j = 0; j < half; ++j) {
      const float send = up ? v[j] : v[j + half];
      const float recv = __shfl_xor_sync(0xffffffffu, send, bit);
      v[j] = (up ? v[j + half] : v[j]) + recv;
    }
  }
  v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
  const int col = lane >> 1;
  if ((lane & 1) == 0 && dst != nullptr && col < limit) atomicAdd(&dst[col], v[0]);
}

// One layer's backward on the tensor cores, two threads per row: dW_k = a_k^T dp_k and, when need_dx, dp <- dp_k W_k^T.
// Order (see clb_tc.cuh, "barrier-free hand-over"): chain operands -> collect the PREVIOUS layer's dW (tensor memory straight
// into the CTA's FP32 partial in L2, vector REDs) -> this layer's dW operand images -> sign mask / bias gradient in the shadow
// of the tensor pipe -> collect the chain.  `wk` / `bk` = kernel [32][32] and bias [32] slots of this layer in the partial, or
// the image's gradient slots (il_w > 0: kernel stored (out, in), width il_w); `dead` = scratch slot of `ain` (or null).
__device__ __forceinline__ void tc_layer_backward2(tc::Ctx& tcx, float (&dp)[16], const float (&ain)[16], bool need_dx,
                                                   const float* build_from, char* img_base, const float* next_img,
                                                   float* wk, float* bk, int il_w, const float4* dead = nullptr) {
  const int lane = tcx.tid & 31;
#if CLB_BWD_ORDER == 2
  // round-1 structure: __syncthreads() between the operand stores and the issue (warps 0 and 7 issue)
  tc::issue_backward3(tcx, dp, ain, need_dx, build_from, img_base, next_img);
  if (dead != nullptr && (tcx.tid >> 5) == 5) {
#pragma unroll
    for (int i = 0; i < 4; ++i) asm volatile("discard.global.L2 [%0], 128;" :: "l"(reinterpret_cast<const char*>(dead) + (size_t)(32 * i + lane) * 128) : "memory");
  }
#else
  uint32_t hi[16], lo[16];
  tc::split16(dp, hi, lo);
#endif
#if CLB_BWD_ORDER == 1
  // chain first, dW images while it runs, dW collected one layer later
  if (need_dx) tc::chain_handover<true>(tcx, hi, lo, build_from, img_base, next_img, lane);
  if (tcx.dw_pending) tc::collect_dw_red(tcx, tcx.pend_wk, tcx.pend_ilw, tcx.pend_bk);     // frees the operand images and the accumulator
  tc::dw_handover(tcx, hi, lo, ain, dead, lane);
  tcx.dw_pending = true; tcx.pend_wk = wk; tcx.pend_ilw = il_w; tcx.pend_bk = bk;
#elif CLB_BWD_ORDER == 0
  // one hand-over per layer (chain operands + dW images); dW collected at the end of this layer
  tc::bwd_handover(tcx, hi, lo, ain, need_dx, build_from, img_base, next_img, dead, lane);
#endif
#if !CLB_BIAS_ONES && !CLB_BIAS_COL
  // what does not feed the tensor cores runs while they work: the bias gradient (column sums of dp)
  bias_red16(dp, bk != nullptr ? bk + 16 * tcx.hf : nullptr, lane, il_w > 0 ? il_w - 16 * tcx.hf : 16);
#endif
  if (need_dx) {
    tc::collect2(tcx, dp);               // delta a_k
    // delta p_{k-1} = delta a_k * leaky'(pre-activation of layer k-1); sign(a_k) == sign(pre-activation), and a_k is still in
    // registers (compare + predicated multiply per value; the round-1 kernel built and re-read a 16-bit mask instead)
#pragma unroll
    for (int j = 0; j < 16; ++j) dp[j] = ain[j] > 0.f ? dp[j] : kLeak * dp[j];
  }
#if CLB_BWD_ORDER != 1
  tc::collect_dw_red(tcx, wk, il_w, bk);
#endif
}

// IL = the model has image layers (their code is compiled out otherwise)
template <int LIK, bool IL>
__global__ void __launch_bounds__(tc::kThreads2, 2) k_obs_tc2(ObsArgs a) {
  constexpr int WP = 32, TR = tc::kThreads, T = tc::kThreads2, NC = WP / 4, HW = 16;   // TR rows per tile, T threads
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const int NL = a.lay.n_layers, L = NL - 1, K = IL ? a.n_img_layers : 0, LT = L + K;
  unsigned char* sp = smem_raw;
  char* tc_dwa = reinterpret_cast<char*>(sp);               // dW A operand: MN groups [a_hi | a_lo | ONES | (delta-p_hi: unused rows)]
  constexpr size_t kOnes = (CLB_BIAS_ONES || CLB_BIAS_COL) ? tc::kDwLBO : 0;  // 16 KB of 1.0f after the a images: rows 64..95 (CLB_BIAS_ONES) / columns 64..71 (CLB_BIAS_COL) of the dW product = column sums of delta-p
  float* ones = reinterpret_cast<float*>(tc_dwa + tc::kDwImgBytes);
  char* tc_dwb = tc_dwa + tc::kDwImgBytes + kOnes;          // dW B operand: [delta-p_hi | delta-p_lo]
  sp += 2 * tc::kDwImgBytes + kOnes;
  float* Whead = reinterpret_cast<float*>(sp);              // [32][2]
  float* bsm = Whead + 64;                                  // [NL][32]
  float* Wimg = bsm + (size_t)NL * WP;                      // [K][32][32] this tile's image-layer kernels as [in][out]
  float* bimg = Wimg + (size_t)K * WP * WP;                 // [K][32]
  double* red = reinterpret_cast<double*>(bimg + (size_t)K * WP);
  char* tc_img = reinterpret_cast<char*>(red + 64);         // [2 buffers][hi, lo][kImgBytes] chain B operand images
  uint64_t* tc_bar = reinterpret_cast<uint64_t*>(tc_img + 4 * tc::kImgBytes);   // [0] chain, [1] dW, [2..3] image buffers
  uint32_t* tc_slot = reinterpret_cast<uint32_t*>(tc_bar + 4);   // [0] tensor-memory base, [2] / [3] arrival counters (chain, dW)
  float2* xch = reinterpret_cast<float2*>(tc_slot + 4);     // [2][128]: head partial sums of hf = 1, then (dmu, drho)

  const int tid = threadIdx.x, lane = tid & 31, rrow = tid & (TR - 1), hf = tid >> 7;
  tc::Ctx tcx{};
  if (tid == 0) {
    tc::mbar_init(tc::smem_u32(tc_bar), 1); tc::mbar_init(tc::smem_u32(tc_bar + 1), 1);
    tc::mbar_init(tc::smem_u32(tc_bar + 2), 1); tc::mbar_init(tc::smem_u32(tc_bar + 3), 1);
    tc_slot[2] = 0u; tc_slot[3] = 0u;
  }
  if (tid < 32) tc::tmem_alloc(tc::smem_u32(tc_slot));
  tc::fence_before();
  for (int idx = tid; idx < WP * 2; idx += T) {
    const int i = idx / 2, j = idx % 2;
    Whead[idx] = (i < a.lay.in_dim[L] && j < a.lay.out_dim[L]) ? a.theta_mlp[a.lay.koff[L] + i * a.lay.out_dim[L] + j] : 0.f;
  }
  for (int idx = tid; idx < NL * WP; idx += T) {
    const int k = idx / WP, j = idx % WP;
    bsm[idx] = (j < a.lay.out_dim[k]) ? a.theta_mlp[a.lay.boff[k] + j] : 0.f;
  }
  for (int idx = tid; idx < (int)(kOnes / 16); idx += T) reinterpret_cast<float4*>(ones)[idx] = make_float4(1.f, 1.f, 1.f, 1.f);
  tc::fence_async_smem();
  __syncthreads();
  {
    tc::fence_after();
    const uint32_t tbase = *tc_slot;
    const int warp = tid >> 5;
    tcx.row_addr = tbase + ((uint32_t)(32 * (warp & 3)) << 16);
    tcx.mbar = tc::smem_u32(tc_bar); tcx.parity = 0;
    tcx.img_hi = tc_img; tcx.img_lo = tc_img + tc::kImgBytes;
    tcx.desc_hi = tc::make_desc(tc::smem_u32(tcx.img_hi)); tcx.desc_lo = tc::make_desc(tc::smem_u32(tcx.img_lo));
    tcx.tid = tid; tcx.base = tbase;
    tcx.mbar_dw = tc::smem_u32(tc_bar + 1); tcx.parity_dw = 0;
    tcx.dw_a = tc_dwa; tcx.dw_b = tc_dwb;
    tcx.desc_dwa = tc::make_desc_mn(tc::smem_u32(tc_dwa)); tcx.desc_dwb = tc::make_desc_mn(tc::smem_u32(tc_dwb));
    tcx.row = rrow; tcx.hf = hf; tcx.col = (uint32_t)(HW * hf);
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      tcx.wimg[b] = tc::smem_u32(tc_img + (size_t)b * 2 * tc::kImgBytes);
      tcx.wdesc_hi[b] = tc::make_desc(tcx.wimg[b]); tcx.wdesc_lo[b] = tc::make_desc(tcx.wimg[b] + tc::kImgBytes);
      tcx.wbar[b] = tc::smem_u32(tc_bar + 2 + b);
    }
    tcx.pass = 0; tcx.wphase = 0;
    tcx.cnt_chain = tc::smem_u32(tc_slot + 2); tcx.cnt_dw = tc::smem_u32(tc_slot + 3); tcx.n_warps_m1 = T / 32 - 1;
    tcx.dw_pending = false; tcx.pend_wk = nullptr; tcx.pend_ilw = 0; tcx.pend_bk = nullptr;
    tcx.lo_off = a.det ? WP * WP : 0;
  }
  // ready-made images of hidden layer k in global memory: dir 0 = forward (B[n][k] = W[k][n]), 1 = backward; null for
  // image layers, whose per-tile kernels are turned into images by the threads themselves
  constexpr size_t IMGF = tc::kImgBytes / 4;
  auto gimg = [&](int k, int dir) -> const float* { return (!IL || k < L) ? a.wimg + ((size_t)(k * 2 + dir) * 2) * IMGF : nullptr; };
  // one layer of the CTA's FP32 partial: kernel [32][32], bias [32]; deterministic mode: kernel from a_hi rows, kernel from a_lo
  // rows, bias per row quarter [4][32] (see ObsArgs::det)
  const int PSLOT = a.det ? tc::kPslotDet : WP * WP + WP;
  const int BOFF = a.det ? 2 * WP * WP + 32 * ((tid >> 5) & 3) : WP * WP;      // this warp's bias slot inside a layer's slot
  float* part32 = a.partials32 + (size_t)(blockIdx.x % a.n_partials) * NL * PSLOT;     // shared by a few CTAs (REDs): small L2 footprint
  float4* scr = a.scratch + (size_t)blockIdx.x * LT * NC * TR;
  double ll_sum = 0.0;
  float ev_f = 1.f, ev_a = 0.f, ev_b = 0.f;
  if (a.theta_lik != nullptr) { ev_f = softplusf(a.theta_lik[0]); ev_a = softplusf(a.theta_lik[1]); ev_b = softplusf(a.theta_lik[2]); }
  const int64_t n_tiles = (a.n_rows + TR - 1) / TR;
  auto wsrc = [&](int k) -> const float* { return (k >= L) ? Wimg + (size_t)(k - L) * WP * WP : a.wpack + (size_t)k * 1024; };

  // the first pass's images (forward, layer 0) start travelling now
  if (tid == 0 && LT > 0 && blockIdx.x < n_tiles && gimg(0, 0) != nullptr) tc::tma_fetch_image(tcx.wimg[0], gimg(0, 0), tcx.wbar[0]);
  CLB_PH_START();
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const bool more_tiles = tile + gridDim.x < n_tiles;
    const int64_t row = tile * TR + rrow;
    const bool inb = row < a.n_rows;
    const int refl = inb ? __ldcs(&a.refl[row]) : -1;
    const bool active = refl >= 0;
    const int timg = (IL && K > 0) ? a.image[tile * TR] : 0;
    if (IL && K > 0) {
      __syncthreads();
      const int w = a.il_width;
      const size_t lstride = (size_t)a.il_n_images * w * (w + 1);
      for (int idx = tid; idx < K * WP * WP; idx += T) {
        const int l = idx / (WP * WP), i = (idx / WP) % WP, j = idx % WP;
        Wimg[idx] = (i < w && j < w) ? a.theta_il[l * lstride + ((size_t)timg * w + j) * w + i] : 0.f;
      }
      for (int idx = tid; idx < K * WP; idx += T) {
        const int l = idx / WP, j = idx % WP;
        bimg[idx] = (j < w) ? a.theta_il[l * lstride + (size_t)a.il_n_images * w * w + (size_t)timg * w + j] : 0.f;
      }
      __syncthreads();
    }
    // ---------------- forward: my 16 features ----------------
    float h[HW];
#pragma unroll
    for (int i = 0; i < HW; ++i) { const int f = HW * hf + i; h[i] = (inb && f < a.d) ? __ldcs(&a.meta[(size_t)f * a.n_rows + row]) : 0.f; }
    for (int k = 0; k < LT; ++k) {
      const float* bk = ((IL && k >= L) ? bimg + (size_t)(k - L) * WP : bsm + (size_t)k * WP) + HW * hf;
      float o[HW];
      CLB_PH(0);
      // the pass after this one: next forward layer, else the first dX pass, else the next tile's first layer
      const float* next = (k + 1 < LT) ? gimg(k + 1, 0) : (a.train_mlp && LT > 1) ? gimg(LT - 1, 1) : (more_tiles ? gimg(0, 0) : nullptr);
#if CLB_BWD_ORDER == 2
      tc::issue3(tcx, h, (IL && k >= L) ? wsrc(k) : nullptr, tc_img, next, CLB_BIAS_IN_MMA ? bk : nullptr);
#else
      tc::issue4(tcx, h, (IL && k >= L) ? wsrc(k) : nullptr, tc_img, next, lane);
#endif
      CLB_PH(1);
      tc::collect2(tcx, o);
      CLB_PH(2);
#pragma unroll
      for (int j = 0; j < HW; ++j) { const float v = (CLB_BIAS_IN_MMA && CLB_BWD_ORDER == 2) ? o[j] : o[j] + bk[j]; h[j] = fmaxf(v, kLeak * v); }
#ifndef CLB_ABL_SCR
      if (a.train_mlp && k + 1 < LT) {          // the last layer's output stays in registers (h) for the head
#pragma unroll
        for (int c = 0; c < 4; ++c) scr[((size_t)k * NC + 4 * hf + c) * TR + rrow] = make_float4(h[4 * c], h[4 * c + 1], h[4 * c + 2], h[4 * c + 3]);
      }
#endif
    }
    CLB_PH(3);
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
      obs_epilogue<LIK>(a, row, inb, active, refl, lane, out0, out1, ev_f, ev_a, ev_b, ll_sum, dmu, drho);
      xch[TR + rrow] = make_float2(dmu, drho);
    }
    CLB_PH(4);
    if (!a.train_mlp) { __syncthreads(); continue; }
    __syncthreads();
    { const float2 g = xch[TR + rrow]; dmu = g.x; drho = g.y; }
    // ---------------- backward ----------------
    float dp[HW], nxt[HW];
    auto load_act = [&](float (&dst)[HW], int k) {           // my half of a_k, the input of chain layer k
      if (k > 0) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
#ifdef CLB_ABL_SCR
          const float4 v = make_float4(dmu + c, drho, dmu - c, drho + k);
#else
          const float4 v = __ldcg(&scr[((size_t)(k - 1) * NC + 4 * hf + c) * TR + rrow]);
#endif
          dst[4 * c] = v.x; dst[4 * c + 1] = v.y; dst[4 * c + 2] = v.z; dst[4 * c + 3] = v.w;
        }
      } else {
#pragma unroll
        for (int i = 0; i < HW; ++i) { const int f = HW * hf + i; dst[i] = (inb && f < a.d) ? __ldcs(&a.meta[(size_t)f * a.n_rows + row]) : 0.f; }
      }
    };
    if (LT > 0) load_act(nxt, LT - 1);
#pragma unroll
    for (int j = 0; j < HW; ++j) dp[j] = 0.f;
    if (hf == 0) { dp[0] = dmu; dp[1] = drho; }
    // head: dW_out = a_L^T [dmu, drho]
    tc_layer_backward2(tcx, dp, h, false, nullptr, tc_img, nullptr, part32 + (size_t)L * PSLOT, part32 + (size_t)L * PSLOT + BOFF, 0);
#pragma unroll
    for (int i = 0; i < HW; ++i) {       // delta a_LT from the head, times leaky' of the last hidden layer (sign of its output h)
      const float2 w = *reinterpret_cast<const float2*>(&Whead[(HW * hf + i) * 2]);
      const float da = w.x * dmu + w.y * drho;
      dp[i] = h[i] > 0.f ? da : kLeak * da;
    }
    for (int k = LT - 1; k >= 0; --k) {
      const bool is_il = IL && k >= L;
      float* il_gk = nullptr; float* il_gb = nullptr;
      if (is_il && a.g_il != nullptr) {
        const int w = a.il_width;
        const size_t lstride = (size_t)a.il_n_images * w * (w + 1);
        il_gk = a.g_il + (size_t)(k - L) * lstride + (size_t)timg * w * w;
        il_gb = a.g_il + (size_t)(k - L) * lstride + (size_t)a.il_n_images * w * w + (size_t)timg * w;
      }
      const int il_w = is_il ? a.il_width : 0;
      float* wk = is_il ? il_gk : part32 + (size_t)k * PSLOT;
      float* bk2 = is_il ? il_gb : part32 + (size_t)k * PSLOT + BOFF;
      float ain[HW];
#pragma unroll
      for (int i = 0; i < HW; ++i) ain[i] = nxt[i];
      if (k > 0) load_act(nxt, k - 1);
      // the pass after this layer's dX: the next layer's dX (k - 1 >= 1), else the next tile's first forward layer
      const float* next = (k > 1) ? gimg(k - 1, 1) : (more_tiles ? gimg(0, 0) : nullptr);
      const float4* dead = (a.discard_scratch && k > 0) ? scr + (size_t)(k - 1) * NC * TR : nullptr;      // the 16 KB slot of a_k
      tc_layer_backward2(tcx, dp, ain, k > 0, is_il ? wsrc(k) : nullptr, tc_img, next, wk, bk2, il_w, dead);
    }
    if (tcx.dw_pending) { tc::collect_dw_red(tcx, tcx.pend_wk, tcx.pend_ilw, tcx.pend_bk); tcx.dw_pending = false; }
  }
  // ---- flush: the log-likelihood sum ----
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
  if (tid < 32) tc::tmem_dealloc(*tc_slot);
}

// Ev11 on the empty Laue slots: each contributes logpdf(0; I_k, sigma'(0; sigma_k)) per MC sample, and because sigma'
// depends on the error-model parameters, also a gradient (the reference differentiates through the padded entries of
// likelihoods/laue.py:13-35 like through any other).  `mult` = number of MC samples.
template <int LIK>
__global__ void __launch_bounds__(256) k_ev11_empty(const float* iobs, const float* sig, int64_t n, const float* theta_lik, float* g_lik,
                                                    double* acc, LikConst lik, float cl, float mult) {
  const float f = softplusf(theta_lik[0]), av = softplusf(theta_lik[1]), bv = softplusf(theta_lik[2]);
  double ll_sum = 0.0;
  float sgf = 0.f, sga = 0.f, sgb = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float ll, g, gf, ga, gb;
    ev11_eval<LIK>(0.f, iobs[i], sig[i], f, av, bv, lik, ll, g, gf, ga, gb);
    ll_sum += (double)ll; sgf += gf; sga += ga; sgb += gb;
  }
  ll_sum = warp_sum(ll_sum); sgf = warp_sum(sgf); sga = warp_sum(sga); sgb = warp_sum(sgb);
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&acc[ACC_LL], ll_sum * (double)mult);
    if (g_lik != nullptr) {
      atomicAdd(&g_lik[0], cl * mult * sgf * sigmoidf(theta_lik[0]));
      atomicAdd(&g_lik[1], cl * mult * sga * sigmoidf(theta_lik[1]));
      atomicAdd(&g_lik[2], cl * mult * sgb * sigmoidf(theta_lik[2]));
    }
  }
}

// Merged results per reflection (io/manager.py:188-197, :209): F = <z>, SigF = sd(z) of the truncated-normal
// surrogate on [low, 1e10] ([3P] tfd.TruncatedNormal mean/variance), I = SigF^2 + F^2, <F^4> on [low, inf)
// (surrogate_posteriors.py:55-72 closed form == scipy truncnorm.moment(4)), SigI = sqrt(max((1e-5 I)^2, <F^4> - I^2)).
// FP64 arithmetic (R is small), FP32 outputs like the reference.
__global__ void __launch_bounds__(256) k_results(const float* v_loc, const float* v_scale, const uint8_t* centric, int64_t R, float eps,
                                                 float* F, float* SigF, float* I, float* SigI) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  const double mu = exp((double)v_loc[r]), sg = exp((double)v_scale[r]) + (double)eps;
  const double low = centric[r] ? 0.0 : 1e-32, high = 1e10;
  const double al = (low - mu) / sg, be = (high - mu) / sg;
  const double isq = 0.3989422804014327;
  const double pa = isq * exp(-0.5 * al * al), pb = isq * exp(-0.5 * be * be);
  const double Z = normcdf(-al) - normcdf(-be);
  const double bpb = (pb == 0.0) ? 0.0 : be * pb;
  const double d1 = (pa - pb) / Z;
  const double mean = mu + sg * d1;
  const double var = sg * sg * (1.0 + (al * pa - bpb) / Z - d1 * d1);
  const double a = low;
  const double aterm = (a * a * a + a * a * mu + a * mu * mu + sg * sg * (3.0 * a + 5.0 * mu) + mu * mu * mu) * pa;
  const double m4 = mu * mu * mu * mu + 6.0 * mu * mu * sg * sg + 3.0 * sg * sg * sg * sg + sg * aterm / normcdf(-al);
  const double sd = sqrt(fmax(var, 0.0));
  const double inten = sd * sd + mean * mean;
  const double ivar = fmax((inten * 1e-5) * (inten * 1e-5), m4 - inten * inten);
  F[r] = (float)mean; SigF[r] = (float)sd; I[r] = (float)inten; SigI[r] = (float)sqrt(ivar);
}

__global__ void __launch_bounds__(256) k_count_obs(const int32_t* refl, int64_t n_rows, float* N) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_rows && refl[i] >= 0) atomicAdd(&N[refl[i]], 1.0f);
}

// Zero-padded [L][32][32] FP32 copy of the hidden-layer kernels for the tensor-core kernels (they read 4 KB per
// pass through L1/L2 instead of keeping 80 KB of weights in shared memory, which makes room for two CTAs per SM).
__global__ void __launch_bounds__(256) k_pack_weights(const float* theta_mlp, MlpLayout lay, float* wpack, int WP) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int L = lay.n_layers - 1;
  if (idx >= L * WP * WP) return;
  const int k = idx / (WP * WP), i = (idx / WP) % WP, j = idx % WP;
  wpack[idx] = (i < lay.in_dim[k] && j < lay.out_dim[k]) ? theta_mlp[lay.koff[k] + i * lay.out_dim[k] + j] : 0.f;
}

// Ready-made B operand images of every hidden layer for k_obs_tc2 (fetched by TMA): per layer [fwd, bwd][hi, lo] in
// the canonical K-major no-swizzle UMMA layout with the padded LBO of clb_tc.cuh; forward B[n][k] = W[k][n] (n = out,
// k = in), backward B[n][k] = W[n][k] (n = in, k = out).  hi = tf32(w), lo = w - hi.
__global__ void __launch_bounds__(256) k_pack_images(const float* theta_mlp, MlpLayout lay, float* wimg) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int L = lay.n_layers - 1;
  if (idx >= L * 2 * 1024) return;
  const int layer = idx >> 11, dir = (idx >> 10) & 1, n = (idx >> 5) & 31, k = idx & 31;
  const int i = dir ? n : k, j = dir ? k : n;          // W[i = in][j = out]
  const float w = (i < lay.in_dim[layer] && j < lay.out_dim[layer]) ? theta_mlp[lay.koff[layer] + i * lay.out_dim[layer] + j] : 0.f;
  const float hi = tc::tf32_rna(w);
  const size_t IMGF = tc::kImgBytes / 4;
  float* base = wimg + ((size_t)(layer * 2 + dir) * 2) * IMGF;
  const uint32_t off = ((k >> 2) * tc::kLBO + (n >> 3) * tc::kSBO + (n & 7) * 16 + (k & 3) * 4) / 4;
  base[off] = hi;
  base[IMGF + off] = w - hi;
}

// k_obs_tc2's partials: [rows][n_layers][32*32 + 32] FP32 (kernel [in][out] padded to 32 x 32, then the bias), summed over
// the CTAs in FP64 in a fixed order.
__global__ void __launch_bounds__(256) k_reduce_partials32(const float* partials, int rows, MlpLayout lay, float* grad, int det, int transposed = 0) {
  // 32 parameters per block, the rows split into 8 contiguous chunks (one per warp) that are summed in parallel and combined in a
  // fixed order: the same deterministic result whatever the timing, 8x shorter dependent-load chains (the one-thread-per-parameter
  // version took 52 us for 74 rows -- a sixth of a small problem's step)
  __shared__ double part[8][32];
  const int pl = threadIdx.x & 31, g = threadIdx.x >> 5;
  const int p = blockIdx.x * 32 + pl;
  const bool live = p < lay.n_params;
  const int PSLOT = det ? tc::kPslotDet : 32 * 32 + 32, PP = lay.n_layers * PSLOT;
  const int BIAS = det ? 2048 : 1024;
  int extra = 0, n_extra = 0;            // deterministic layout: further slots to add (a_lo rows; the other row quarters' bias sums)
  int src = -1;
  for (int k = 0; live && k < lay.n_layers; ++k) {
    const int nk = lay.in_dim[k] * lay.out_dim[k];
    if (p >= lay.koff[k] && p < lay.koff[k] + nk) {
      const int i = (p - lay.koff[k]) / lay.out_dim[k], j = (p - lay.koff[k]) % lay.out_dim[k];
      src = k * PSLOT + (transposed ? tc::dw_slot32(j, i) : tc::dw_slot32(i, j));     // CLB_BIAS_COL: k_obs_tc2 stores the slot transposed
      if (det) { extra = 1024; n_extra = 1; }
      break;
    }
    if (p >= lay.boff[k] && p < lay.boff[k] + lay.out_dim[k]) { src = k * PSLOT + BIAS + (p - lay.boff[k]); if (det) { extra = 32; n_extra = 3; } break; }
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

// Sum the per-CTA partial weight gradients (padded layout, see partial_row_size) into the flat keras-order
// gradient: grad[p] = sum_rows partials[row][src(p)]  (fixed order => deterministic).
__global__ void __launch_bounds__(256) k_reduce_partials(const double* partials, int rows, MlpLayout lay, int WP, float* grad) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= lay.n_params) return;
  const int PP = partial_row_size(lay.n_layers, WP);
  int src = -1;
  for (int k = 0; k < lay.n_layers; ++k) {
    const int nk = lay.in_dim[k] * lay.out_dim[k];
    if (p >= lay.koff[k] && p < lay.koff[k] + nk) {
      const int i = (p - lay.koff[k]) / lay.out_dim[k], j = (p - lay.koff[k]) % lay.out_dim[k];
      src = k * WP * WP + partial_elem(i, j, WP);
      break;
    }
    if (p >= lay.boff[k] && p < lay.boff[k] + lay.out_dim[k]) {
      src = lay.n_layers * WP * WP + k * WP + (p - lay.boff[k]);
      break;
    }
  }
  double s = 0.0;
  for (int r = 0; r < rows; ++r) s += partials[(size_t)r * PP + src];
  grad[p] = (float)s;
}

// ---------------------------------------------------------------------------------------
// Per-reflection backward: chain dL/dz to (v_loc, v_scale).  One thread per reflection.
// ---------------------------------------------------------------------------------------
struct ReflBwdArgs {
  const float* v_loc; const float* v_scale;
  const float4* bwd_coef; const float* gz;        // (S,R): written by k_refl_sample; dL/dz after the observation kernel
  float* g_loc; float* g_scale;
  int64_t R; int S;
  // fused tail of the per-reflection chain (rows A8 / A9 on the surrogate slice)
  double* var_sums;             // [0..1] raw / filtered sum of squares of g_loc, [2..3] of g_scale (null: do not accumulate)
  double* ss_part;              // deterministic mode: [4][gridDim.x] per-block partials instead of atomics on var_sums
  // Adam on (v_loc, v_scale) inside this kernel: legal whenever no NORM-based clipping is configured, because then the update
  // of an element depends on nothing but its own gradient ([3P] tf_keras Adam.update_step; non-finite elements -> 0,
  // variational.py:208).  alpha = lr sqrt(1 - b2^t) / (1 - b1^t) comes from the host; stop_step as in k_adam.
  float* theta_loc; float* theta_scale; float* m_loc; float* m_scale; float* v2_loc; float* v2_scale;   // null: no fused update
  float alpha, beta1, beta2, adam_eps, clipvalue;
  const int* stop_step; int step_index;
};

// Chain dL/dz to (v_loc, v_scale): g_loc = sum_s gz A + C, g_scale = sum_s gz B + D with the four coefficients the forward
// kernel left behind -- no transcendental arithmetic here, the kernel streams {gz, coefficients, parameters, Adam moments} and
// is bound by HBM bandwidth.  Four reflections per thread, 16-byte accesses (see k_refl_sample).
__global__ void __launch_bounds__(256) k_refl_backward(ReflBwdArgs a, int vec) {
  const int64_t nq = (a.R + 3) >> 2;
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double ss[4] = {0.0, 0.0, 0.0, 0.0};                   // raw / filtered sums of squares of g_loc, g_scale
  if (q < nq) {
    const int64_t r0 = q << 2;
    const bool v = vec != 0;
    Vec4 gl, gs;
#pragma unroll
    for (int j = 0; j < 4; ++j) { gl.v[j] = 0.f; gs.v[j] = 0.f; }
    for (int s = 0; s < a.S; ++s) {
      const Vec4 g = ld4(a.gz + (size_t)s * a.R, r0, a.R, v);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (r0 + j < a.R) {
          const float4 c = __ldcs(&a.bwd_coef[(size_t)s * a.R + r0 + j]);
          gl.v[j] += fmaf(g.v[j], c.x, c.z);
          gs.v[j] += fmaf(g.v[j], c.y, c.w);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (r0 + j < a.R) {
        const double a2 = (double)gl.v[j] * (double)gl.v[j], b2 = (double)gs.v[j] * (double)gs.v[j];
        ss[0] += a2; ss[2] += b2;
        if (isfinite(gl.v[j])) ss[1] += a2;
        if (isfinite(gs.v[j])) ss[3] += b2;
      }
    }
    st4(a.g_loc, r0, a.R, v, gl);
    st4(a.g_scale, r0, a.R, v, gs);
    if (a.theta_loc != nullptr && a.step_index <= *a.stop_step) {
      const Vec4 vl = ld4(a.v_loc, r0, a.R, v), vs = ld4(a.v_scale, r0, a.R, v);
      auto adam = [&](float* theta, float* m, float* v2, const Vec4& par, const Vec4& grad) {
        Vec4 mm = ld4(m, r0, a.R, v), vv = ld4(v2, r0, a.R, v), th;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float g = isfinite(grad.v[j]) ? grad.v[j] : 0.f;
          if (a.clipvalue > 0.f) g = fminf(fmaxf(g, -a.clipvalue), a.clipvalue);
          mm.v[j] += (g - mm.v[j]) * (1.0f - a.beta1);
          vv.v[j] += (g * g - vv.v[j]) * (1.0f - a.beta2);
          th.v[j] = par.v[j] - a.alpha * mm.v[j] / (sqrtf(vv.v[j]) + a.adam_eps);
        }
        st4(m, r0, a.R, v, mm); st4(v2, r0, a.R, v, vv); st4(theta, r0, a.R, v, th);
      };
      adam(a.theta_loc, a.m_loc, a.v2_loc, vl, gl);
      adam(a.theta_scale, a.m_scale, a.v2_scale, vs, gs);
    }
  }
  if (a.var_sums != nullptr) {
    __shared__ double sm[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i) { ss[i] = warp_sum(ss[i]); if ((threadIdx.x & 31) == 0) sm[i][threadIdx.x >> 5] = ss[i]; }
    __syncthreads();
    if (threadIdx.x < 4) {
      double t = 0.0;
      for (int i = 0; i < 8; ++i) t += sm[threadIdx.x][i];
      if (a.ss_part != nullptr) a.ss_part[(size_t)threadIdx.x * gridDim.x + blockIdx.x] = t;
      else if (t != 0.0) atomicAdd(&a.var_sums[threadIdx.x], t);      // NaN != 0 is true, so NaN propagates
    }
  }
}

// ---------------------------------------------------------------------------------------
// Optimiser: per-variable sums of squares, finalize (norms, clip factors, metrics), Adam.
// ---------------------------------------------------------------------------------------
struct VarTable {
  int n_vars;
  int64_t off[kMaxVars], size[kMaxVars];
  int trainable[kMaxVars];
  int replicated[kMaxVars];     // 1: gradient is all-reduced across ranks (count its norm once)
};

// grid = (chunks, n_vars).  sums[2*v] = raw sum of squares (NaN/inf propagate), sums[2*v+1] = filtered.
// which: 0 = every variable, 1 = rank-local variables only, 2 = replicated variables only (in-library exchange: the local
// sums travel with the all-reduce, the replicated ones are taken from the reduced gradient afterwards, identically on every rank).
__global__ void __launch_bounds__(256) k_var_sumsq(const float* grad, VarTable vt, double* sums, int which, int skip_surrogate) {
  const int v = blockIdx.y;
  if (!vt.trainable[v]) return;
  if ((which == 1 && vt.replicated[v]) || (which == 2 && !vt.replicated[v])) return;
  if (skip_surrogate && v < 2) return;          // the surrogate's sums were accumulated by k_refl_backward
  const int64_t n = vt.size[v];
  const float* g = grad + vt.off[v];
  double raw = 0.0, filt = 0.0;
  auto add = [&](float x) { const double x2 = (double)x * (double)x; raw += x2; if (isfinite(x)) filt += x2; };
  if ((reinterpret_cast<uintptr_t>(g) & 15) == 0) {       // 16-byte accesses; the (< 4) tail elements by the first threads
    const int64_t n4 = n >> 2;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
      const float4 x = reinterpret_cast<const float4*>(g)[i];
      add(x.x); add(x.y); add(x.z); add(x.w);
    }
    const int64_t t = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) add(g[t]);
  } else {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) add(g[i]);
  }
  raw = warp_sum(raw); filt = warp_sum(filt);
  __shared__ double sm[2][8];
  if ((threadIdx.x & 31) == 0) { sm[0][threadIdx.x >> 5] = raw; sm[1][threadIdx.x >> 5] = filt; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0;
    for (int i = 0; i < 8; ++i) { a += sm[0][i]; b += sm[1][i]; }
    if (a != 0.0) atomicAdd(&sums[2 * v], a);      // NaN != 0 is true, so NaN propagates
    if (b != 0.0) atomicAdd(&sums[2 * v + 1], b);
  }
}

struct FinalizeArgs {
  const double* acc;            // local accumulators
  double* red;                  // reduce scalars: [0]=logq-logp sum, [1]=ll sum, [2..]=per-var sums (2 each)
  VarTable vt;
  double* metrics;              // [4] loss nll kl gradnorm for this step
  float* var_scale;             // per-variable gradient scale from clipping
  float* adam_alpha;            // [1]
  int* stop_step; int step;
  double kl_div, kl_coef, ll_div;   // kl = sum/kl_div ; loss = kl_coef*kl - ll/ll_div
  float clipnorm, global_clipnorm;
  float lr, beta1, beta2; int64_t t;   // t = step index (1-based) for bias correction
  float alpha;                  // lr sqrt(1 - beta2^t) / (1 - beta1^t), computed on the host (the same value k_refl_backward uses)
};

// Packs the local scalars into the reduce buffer (so one all-reduce covers them).
// fused != 0 (in-library exchange): the replicated variables' slots are zeroed here and filled after the all-reduce.
__global__ void k_pack_scalars(const double* acc, const double* var_sums, VarTable vt, double* red, int rank,
                               double ll_const, int fused) {
  const int i = threadIdx.x;
  if (i == 0) red[0] = acc[ACC_LOGQ_MINUS_LOGP];
  if (i == 1) red[1] = acc[ACC_LL] + ll_const;   // + constant log-density of this rank's empty Laue slots
  if (i < vt.n_vars) {
    // replicated variables hold the same (all-reduced) gradient on every rank: count them once
    const bool mine = fused ? !vt.replicated[i] : (!vt.replicated[i] || rank == 0);
    red[2 + 2 * i] = mine ? var_sums[2 * i] : 0.0;
    red[3 + 2 * i] = mine ? var_sums[2 * i + 1] : 0.0;
  }
}

__global__ void k_finalize(FinalizeArgs a) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const double kl = a.red[0] / a.kl_div;
  const double ll = a.red[1] / a.ll_div;
  double raw = 0.0, filt = 0.0;
  for (int v = 0; v < a.vt.n_vars; ++v) {
    if (!a.vt.trainable[v]) continue;
    raw += a.red[2 + 2 * v];
    filt += a.red[3 + 2 * v];
  }
  const double gn = sqrt(raw);
  a.metrics[0] = a.kl_coef * kl - ll;
  a.metrics[1] = -ll;
  a.metrics[2] = kl;
  a.metrics[3] = gn;
  const float gscale = (a.global_clipnorm > 0.f) ? (float)(a.global_clipnorm / fmax(sqrt(filt), (double)a.global_clipnorm)) : 1.0f;
  for (int v = 0; v < a.vt.n_vars; ++v) {
    float sc = 1.0f;
    if (a.clipnorm > 0.f) sc = (float)(a.clipnorm / fmax(sqrt(a.red[3 + 2 * v]), (double)a.clipnorm));
    a.var_scale[v] = sc * gscale;
  }
  a.adam_alpha[0] = a.alpha;
  if (!isfinite(gn) && a.step < *a.stop_step) *a.stop_step = a.step;
}

// grid = (chunks, n_vars).  [3P] tf_keras Adam.update_step; non-finite gradient elements -> 0 (variational.py:208).
// 16-byte accesses when the variable's slice is aligned (m, v, theta, grad share the offset).  skip_surrogate: variables 0 / 1
// (v_loc, v_scale) were already updated inside k_refl_backward.
__global__ void __launch_bounds__(256) k_adam(float* theta, float* m, float* v, const float* grad, VarTable vt,
                                              const float* var_scale, const float* adam_alpha,
                                              float clipvalue, float beta1, float beta2, float adam_eps,
                                              const int* stop_step, int step, int skip_surrogate) {
  const int var = blockIdx.y;
  if (!vt.trainable[var]) return;
  if (skip_surrogate && var < 2) return;
  // variational.py:271-274: the loop breaks AFTER the step whose norm was non-finite has been
  // applied (with the filtered gradient); later enqueued steps must not touch the state.
  if (step > *stop_step) return;
  const float sc = var_scale[var];
  const float alpha = adam_alpha[0];
  const int64_t n = vt.size[var], off = vt.off[var];
  auto upd = [&](float g, float& mi, float& vi, float& th) {
    g = isfinite(g) ? g * sc : 0.f;
    if (clipvalue > 0.f) g = fminf(fmaxf(g, -clipvalue), clipvalue);
    mi += (g - mi) * (1.0f - beta1);
    vi += (g * g - vi) * (1.0f - beta2);
    th -= alpha * mi / (sqrtf(vi) + adam_eps);
  };
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
  if ((off & 3) == 0) {
    const int64_t n4 = n >> 2;
    float4* t4 = reinterpret_cast<float4*>(theta + off); float4* m4 = reinterpret_cast<float4*>(m + off);
    float4* v4 = reinterpret_cast<float4*>(v + off); const float4* g4 = reinterpret_cast<const float4*>(grad + off);
    for (int64_t i = tid; i < n4; i += stride) {
      const float4 g = g4[i];
      float4 mi = m4[i], vi = v4[i], th = t4[i];
      upd(g.x, mi.x, vi.x, th.x); upd(g.y, mi.y, vi.y, th.y); upd(g.z, mi.z, vi.z, th.z); upd(g.w, mi.w, vi.w, th.w);
      m4[i] = mi; v4[i] = vi; t4[i] = th;
    }
    const int64_t t = (n4 << 2) + tid;
    if (t < n) upd(grad[off + t], m[off + t], v[off + t], theta[off + t]);
  } else {
    for (int64_t i = tid; i < n; i += stride) upd(grad[off + i], m[off + i], v[off + i], theta[off + i]);
  }
}

}  // namespace clb

#include "clb_tc16.cuh"
#include "clb_pp.cuh"
#include "clb_tc3.cuh"
