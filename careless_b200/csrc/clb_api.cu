// clb_api.cu -- C-ABI of libcareless_b200.so (see include/careless_b200.h).
//
// Host side of the B200-native ELBO gradient + Adam step: device memory, the host prep that
// turns the reference's input tuple (careless/models/base.py:22-31) into the sorted / padded
// SoA device layout, kernel launches, and the parameter / optimiser state.
#include "../../include/careless_b200.h"
#include "clb_kernels.cuh"
#include "clb_prep.cuh"

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <cstring>
#include <dlfcn.h>
#include <string>
#include <vector>

using namespace clb;

namespace {

std::string g_create_error;

struct DevBuf {
  void* p = nullptr; size_t bytes = 0;
  ~DevBuf() { if (p) cudaFree(p); }
  cudaError_t alloc(size_t n) {
    if (p && n <= bytes) return cudaSuccess;
    if (p) { cudaFree(p); p = nullptr; bytes = 0; }
    if (n == 0) return cudaSuccess;
    cudaError_t e = cudaMalloc(&p, n);
    if (e == cudaSuccess) bytes = n;
    return e;
  }
  void free() { if (p) { cudaFree(p); p = nullptr; bytes = 0; } }
  template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct PinBuf {
  void* p = nullptr; size_t bytes = 0;
  ~PinBuf() { if (p) cudaFreeHost(p); }
  cudaError_t alloc(size_t n) {
    if (p && n <= bytes) return cudaSuccess;
    if (p) { cudaFreeHost(p); p = nullptr; bytes = 0; }
    if (n == 0) return cudaSuccess;
    cudaError_t e = cudaMallocHost(&p, n);
    if (e == cudaSuccess) bytes = n;
    return e;
  }
  template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

// ---- NCCL, bound at run time (dlopen) -----------------------------------------------------------------------------
// The library has no link-time dependency on NCCL: a single-GPU user (and the CPU-side build check) never needs it, and a
// process that has already loaded an NCCL (torch bundles its own libnccl.so.2) gets THAT copy back from dlopen instead of
// a second one.  Only the few entry points of the step's exchange are bound; types follow nccl.h (2.x ABI).
struct NcclUniqueId { char internal[128]; };
typedef struct ncclComm* NcclComm;
enum { kNcclSum = 0, kNcclFloat32 = 7, kNcclFloat64 = 8 };
struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(NcclUniqueId*) = nullptr;
  int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
  int (*CommDestroy)(NcclComm) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  std::string err;
  bool load() {
    if (lib) return true;
    const char* names[] = {getenv("CLB_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      if (!n || !n[0]) continue;
      lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (lib) break;
    }
    if (!lib) { err = std::string("cannot load NCCL (libnccl.so.2): ") + (dlerror() ? dlerror() : "?"); return false; }
    auto sym = [&](const char* n) { void* p = dlsym(lib, n); if (!p) err = std::string("NCCL symbol missing: ") + n; return p; };
    GetUniqueId = reinterpret_cast<decltype(GetUniqueId)>(sym("ncclGetUniqueId"));
    CommInitRank = reinterpret_cast<decltype(CommInitRank)>(sym("ncclCommInitRank"));
    CommDestroy = reinterpret_cast<decltype(CommDestroy)>(sym("ncclCommDestroy"));
    AllReduce = reinterpret_cast<decltype(AllReduce)>(sym("ncclAllReduce"));
    GroupStart = reinterpret_cast<decltype(GroupStart)>(sym("ncclGroupStart"));
    GroupEnd = reinterpret_cast<decltype(GroupEnd)>(sym("ncclGroupEnd"));
    GetErrorString = reinterpret_cast<decltype(GetErrorString)>(sym("ncclGetErrorString"));
    if (!GetUniqueId || !CommInitRank || !CommDestroy || !AllReduce || !GroupStart || !GroupEnd || !GetErrorString) { lib = nullptr; return false; }
    return true;
  }
};
NcclApi g_nccl;

}  // namespace

struct clb_handle {
  clb_config cfg{};
  std::string err;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  int n_sms = 0;

  // sizes
  int64_t R = 0; int S = 1;
  int64_t n_rows_raw = 0, n_rows = 0 /*padded*/, n_rows_total = 0;
  int WP = 0, NL = 0, KS = 1, grid_obs = 0, n_partials = 0;
  size_t smem_obs = 0;
  MlpLayout lay{};
  VarTable vt{};
  int var_group[kMaxVars]{};
  int64_t goff[CLB_N_GROUPS]{}, gsize[CLB_N_GROUPS]{};
  int gtrain[CLB_N_GROUPS]{};
  int64_t P = 0;           // total parameters
  int64_t adam_t = 0;
  uint32_t step_counter = 0;   // RNG step index
  bool have_obs = false, have_prior = false, in_step = false;
  bool eval_mode = false;    // clb_eval: forward only (no gradients, no Adam)
  bool adam_fused = false;   // this step's surrogate Adam update ran inside k_refl_backward
  bool no_fused_adam = false;   // CLB_FUSED_ADAM=0: keep the surrogate's update in k_adam (A/B and tests)
  bool use_tc = false;       // tensor-core (tcgen05) path of k_obs: padded width 32, unless CLB_NO_TC=1
  bool use_tc2 = false;      // ... with two threads per observation row (k_obs_tc2), unless CLB_TC_ONE_THREAD_PER_ROW=1
  bool use_pp = false;       // width 32 without image layers: warp-specialised two-tile ping-pong kernel (k_obs_pp), unless CLB_PP=0
  bool use_tc3 = false;      // width 32 without image layers: k_obs_tc2 software-pipelined across tiles (k_obs_tc3), CLB_TC3=1
  bool use_tc16 = false;     // narrow MLPs (padded width <= 16, with or without image layers) on the tensor cores (k_obs_tc16), unless CLB_TC16=0
  bool discard_scratch = true;   // consumed activation-scratch lines are discarded from L2 (no DRAM write-back), unless CLB_DISCARD=0
  bool debug_sync = false;   // CLB_DEBUG_SYNC=1: synchronise + log after every kernel launch
  int order = CLB_ORDER_REFL;
  double ll_const = 0.0;   // per-sample constant log-likelihood of empty Laue slots
  DevBuf empty_slots;      // Ev11: (I_k) then (sigma_k) of the empty Laue slots
  int64_t n_empty = 0;

  // device state
  DevBuf theta, m, v, grad;
  DevBuf centric, eps_sigma, dw_parent, asu_id, r_const, refl_index;
  DevBuf z, gz, bwd_coef;
  // deterministic mode
  bool det = false;
  DevBuf dzf_rows, refl_ptr, refl_rows, ll_part, kl_part, ss_part;
  int kl_blocks = 0, ss_blocks = 0;
  DevBuf rows;             // one allocation holding all row arrays
  PinBuf rows_host;        // pinned mirror (for re-upload)
  bool rows_host_valid = false;   // device-side prep: the mirror is filled (device -> host) only when a re-upload is asked for
  bool device_prep = true;        // rows are sorted / padded / gathered on the GPU (clb_prep.cuh), unless CLB_DEVICE_PREP=0
  bool bias_feat15 = false;       // k_obs_tc16, max(n_meta, mlp_width) <= 15: bias gradients as row 15 of the dW products (unless CLB_BIAS_FEAT=0)
  double prep_ms = 0.0;           // wall time of the last clb_set_observations row preparation
  size_t rows_bytes = 0;
  // double-buffered input pipeline (clb_prefetch_observations): the next step's rows travel on a copy stream into the
  // other buffer while the current step computes; step_begin switches buffers once the copy has landed
  DevBuf rows_alt;
  size_t row_off[7] = {0, 0, 0, 0, 0, 0, 0};   // refl, image, spot, oidx, meta, iobs, sig
  bool has_img_rows = false, has_spot_rows = false, pending_swap = false, alt_used = false;
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_copy_done = nullptr, ev_rows_free[2] = {nullptr, nullptr};   // [b]: the last kernel reading buffer b is done
  int cur_rows = 0;          // which of the two events belongs to the buffer `rows` currently points to
  int32_t *d_refl = nullptr, *d_image = nullptr, *d_spot = nullptr; uint32_t* d_oidx = nullptr;
  float *d_meta = nullptr, *d_iobs = nullptr, *d_sig = nullptr;
  DevBuf partials, wpack, wimg;     // partials: [weight-gradient partials | activation scratch]
  size_t partial_bytes = 0; float4* scratch_ptr = nullptr;
  int obs_threads = kObsThreads;   // rows per CTA tile: 256 (FP32 kernels) or 128 (tensor-core kernels, 2 CTAs per SM)
  DevBuf acc, var_sums, red, metrics, var_scale, adam_alpha, stop_step;
  DevBuf inj_u, inj_eps, ipred, scale_mom, results;
  bool want_scale_moments = false;
  bool want_ipred = false;
  int metrics_cap = 0;

  // timing
  bool timing = false;
  std::vector<cudaEvent_t> ev;   // pairs
  size_t ev_used = 0;
  double obs_ms = 0.0; int64_t obs_launches = 0, total_launches = 0;

  // in-library exchange step (clb_comm_init): one grouped NCCL all-reduce per step on the handle's stream
  NcclComm comm = nullptr;
  int64_t exchanges = 0;

  ~clb_handle() {
    if (comm && g_nccl.CommDestroy) g_nccl.CommDestroy(comm);
    for (auto e : ev) cudaEventDestroy(e);
    if (ev_copy_done) cudaEventDestroy(ev_copy_done);
    for (auto e : ev_rows_free) if (e) cudaEventDestroy(e);
    if (copy_stream) cudaStreamDestroy(copy_stream);
    if (own_stream && stream) cudaStreamDestroy(stream);
  }
};

namespace {

int fail(clb_handle* h, int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
  if (h) h->err = buf; else g_create_error = buf;
  return code;
}

#define CLB_CUDA(h, expr)                                                                      \
  do { cudaError_t e_ = (expr);                                                                \
       if (e_ != cudaSuccess) return fail(h, CLB_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,     \
                                          cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)

#define CLB_LAUNCHED(h)                                                                        \
  do { CLB_CUDA(h, cudaGetLastError()); (h)->total_launches++;                                 \
       if ((h)->debug_sync) { fprintf(stderr, "[clb] launch %lld at %s:%d ...", (long long)(h)->total_launches, __FILE__, __LINE__); \
                              fflush(stderr); CLB_CUDA(h, cudaStreamSynchronize((h)->stream)); fprintf(stderr, " done\n"); } } while (0)

int round_width(int w) { return w <= 8 ? 8 : w <= 16 ? 16 : w <= 32 ? 32 : w <= 64 ? 64 : -1; }

template <int WP, int LIK, bool TC> cudaError_t launch_obs(clb_handle* h, const ObsArgs& a) {
  cudaError_t e = cudaFuncSetAttribute(k_obs<WP, LIK, TC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_obs);
  if (e != cudaSuccess) return e;
  k_obs<WP, LIK, TC><<<h->grid_obs, TC ? tc::kThreads : kObsThreads, h->smem_obs, h->stream>>>(a);
  return cudaGetLastError();
}

template <int LIK, bool IL> cudaError_t launch_obs_tc2(clb_handle* h, const ObsArgs& a) {
  cudaError_t e = cudaFuncSetAttribute(k_obs_tc2<LIK, IL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_obs);
  if (e != cudaSuccess) return e;
  k_obs_tc2<LIK, IL><<<h->grid_obs, tc::kThreads2, h->smem_obs, h->stream>>>(a);
  return cudaGetLastError();
}

cudaError_t dispatch_obs(clb_handle* h, const ObsArgs& a) {
  const int lik = h->cfg.likelihood;
  if (h->use_tc16) {
    const bool il = h->cfg.image_layers > 0;
    auto kern = il ? (lik ? k_obs_tc16<1, true> : k_obs_tc16<0, true>) : (lik ? k_obs_tc16<1, false> : k_obs_tc16<0, false>);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_obs);
    if (e != cudaSuccess) return e;
    kern<<<h->grid_obs, tc16::kRows, h->smem_obs, h->stream>>>(a);
    return cudaGetLastError();
  }
  if (h->use_pp) {
    auto kern = lik ? k_obs_pp<1> : k_obs_pp<0>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_obs);
    if (e != cudaSuccess) return e;
    kern<<<h->grid_obs, pp::kThreadsPP, h->smem_obs, h->stream>>>(a);
    return cudaGetLastError();
  }
  if (h->use_tc3) {
    auto kern = lik ? k_obs_tc3<1> : k_obs_tc3<0>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_obs);
    if (e != cudaSuccess) return e;
    kern<<<h->grid_obs, tc3::kThreads3, h->smem_obs, h->stream>>>(a);
    return cudaGetLastError();
  }
  if (h->use_tc2) {
    if (h->cfg.image_layers > 0) return lik ? launch_obs_tc2<1, true>(h, a) : launch_obs_tc2<0, true>(h, a);
    return lik ? launch_obs_tc2<1, false>(h, a) : launch_obs_tc2<0, false>(h, a);
  }
  switch (h->WP) {
    case 8:  return lik ? launch_obs<8, 1, false>(h, a) : launch_obs<8, 0, false>(h, a);
    case 16: return lik ? launch_obs<16, 1, false>(h, a) : launch_obs<16, 0, false>(h, a);
    case 32:
      if (h->use_tc) return lik ? launch_obs<32, 1, true>(h, a) : launch_obs<32, 0, true>(h, a);
      return lik ? launch_obs<32, 1, false>(h, a) : launch_obs<32, 0, false>(h, a);
    case 64: return lik ? launch_obs<64, 1, false>(h, a) : launch_obs<64, 0, false>(h, a);     // FP32-FMA only (no tensor-core tiling at this width yet)
  }
  return cudaErrorInvalidValue;
}


void build_layout(clb_handle* h) {
  const clb_config& c = h->cfg;
  MlpLayout& l = h->lay;
  int off = 0, fan_in = c.n_meta;
  l.n_layers = c.mlp_layers + 1;
  for (int k = 0; k < c.mlp_layers; ++k) {
    l.in_dim[k] = fan_in; l.out_dim[k] = c.mlp_width;
    l.koff[k] = off; off += fan_in * c.mlp_width;
    l.boff[k] = off; off += c.mlp_width;
    fan_in = c.mlp_width;
  }
  const int k = c.mlp_layers;
  l.in_dim[k] = fan_in; l.out_dim[k] = 2;
  l.koff[k] = off; off += fan_in * 2;
  l.boff[k] = off; off += 2;
  l.n_params = off;
}

void build_vars(clb_handle* h) {
  const clb_config& c = h->cfg;
  h->gsize[CLB_GROUP_SF_LOC] = h->R;
  h->gsize[CLB_GROUP_SF_SCALE] = h->R;
  h->gsize[CLB_GROUP_MLP] = h->lay.n_params;
  h->gsize[CLB_GROUP_IMAGE_SCALES] = c.image_scales ? std::max(0, c.n_images - 1) : 0;
  h->gsize[CLB_GROUP_DW_R] = (c.prior == CLB_PRIOR_DOUBLE_WILSON && c.optimize_dw_r) ? c.n_asu : 0;
  h->gsize[CLB_GROUP_IMAGE_LAYERS] = (int64_t)c.image_layers * c.n_images * c.mlp_width * (c.mlp_width + 1);
  h->gsize[CLB_GROUP_LIKELIHOOD] = c.refine_uncertainties ? 3 : 0;
  int64_t off = 0;
  for (int g = 0; g < CLB_N_GROUPS; ++g) { h->goff[g] = off; off += h->gsize[g]; h->gtrain[g] = 1; }
  h->P = off;
  VarTable& vt = h->vt;
  int n = 0;
  auto add = [&](int group, int64_t o, int64_t sz, int repl) {
    vt.off[n] = o; vt.size[n] = sz; vt.trainable[n] = 1; vt.replicated[n] = repl; h->var_group[n] = group; ++n;
  };
  add(CLB_GROUP_SF_LOC, h->goff[CLB_GROUP_SF_LOC], h->R, 0);
  add(CLB_GROUP_SF_SCALE, h->goff[CLB_GROUP_SF_SCALE], h->R, 0);
  for (int k = 0; k < h->lay.n_layers; ++k) {
    add(CLB_GROUP_MLP, h->goff[CLB_GROUP_MLP] + h->lay.koff[k], (int64_t)h->lay.in_dim[k] * h->lay.out_dim[k], 1);
    add(CLB_GROUP_MLP, h->goff[CLB_GROUP_MLP] + h->lay.boff[k], h->lay.out_dim[k], 1);
  }
  if (h->gsize[CLB_GROUP_IMAGE_SCALES] > 0) add(CLB_GROUP_IMAGE_SCALES, h->goff[CLB_GROUP_IMAGE_SCALES], h->gsize[CLB_GROUP_IMAGE_SCALES], 1);
  if (h->gsize[CLB_GROUP_DW_R] > 0) add(CLB_GROUP_DW_R, h->goff[CLB_GROUP_DW_R], h->gsize[CLB_GROUP_DW_R], 1);
  for (int l = 0; l < c.image_layers; ++l) {       // keras variables of each ImageLayer: kernel, bias
    const int64_t w = c.mlp_width, stride = (int64_t)c.n_images * w * (w + 1);
    add(CLB_GROUP_IMAGE_LAYERS, h->goff[CLB_GROUP_IMAGE_LAYERS] + l * stride, (int64_t)c.n_images * w * w, 1);
    add(CLB_GROUP_IMAGE_LAYERS, h->goff[CLB_GROUP_IMAGE_LAYERS] + l * stride + (int64_t)c.n_images * w * w, (int64_t)c.n_images * w, 1);
  }
  for (int64_t i = 0; i < h->gsize[CLB_GROUP_LIKELIHOOD]; ++i)    // Sdfac, Sdadd, SdB: three scalar keras variables
    add(CLB_GROUP_LIKELIHOOD, h->goff[CLB_GROUP_LIKELIHOOD] + i, 1, 1);
  vt.n_vars = n;
}

void refresh_trainable(clb_handle* h) {
  for (int v = 0; v < h->vt.n_vars; ++v) h->vt.trainable[v] = h->gtrain[h->var_group[v]];
}

double lgamma_d(double x) { return std::lgamma(x); }

// [3P] tf_keras Adam: alpha_t = lr sqrt(1 - beta2^t) / (1 - beta1^t); computed once per step on the host and used by every kernel
float adam_alpha_host(const clb_config& c, int64_t t) {
  return (float)((double)c.learning_rate * std::sqrt(1.0 - std::pow((double)c.beta_2, (double)t)) / (1.0 - std::pow((double)c.beta_1, (double)t)));
}

// log-density of x=0 under the likelihood with (loc, scale): the empty Laue slots (laue.py:23-25)
double lik_logpdf_zero(const clb_config& c, double loc, double scale) {
  const double t = (0.0 - loc) / scale;
  if (c.likelihood == CLB_LIK_NORMAL) return -0.5 * t * t - std::log(scale) - 0.5 * std::log(2.0 * M_PI);
  const double v = c.dof;
  return lgamma_d(0.5 * (v + 1.0)) - lgamma_d(0.5 * v) - 0.5 * std::log(v * M_PI) - std::log(scale)
         - 0.5 * (v + 1.0) * std::log1p(t * t / v);
}

// ---------------------------------------------------------------------------------------
// Host prep: reference input tuple (base.py:22-31) -> sorted / padded SoA rows.
//   mono: stable counting sort by refl_id  (segmented reduction of dL/dz_f inside warps)
//   Laue: stable counting sort by harmonic_id; spots padded so none straddles a 32-row warp chunk
// ---------------------------------------------------------------------------------------
struct RowPlan {
  int order = CLB_ORDER_REFL;
  int64_t npad = 0;
  double ll_const = 0.0;              // sum over empty Laue slots of logpdf(0; I_k, sigma_k)
  std::vector<float> empty;           // (I_k, sigma_k) of the empty Laue slots: first all I, then all sigma (Ev11 needs them on the device)
  std::vector<int32_t> perm;          // sorted position -> original row
  std::vector<int64_t> pos;           // sorted position -> padded row
};

// Argument checks and the resolution of CLB_ORDER_AUTO shared by the host and the device preparation.
int resolve_order(std::string& err, int64_t n, int laue, const int64_t* refl_id, const int64_t* image_id, const float* metadata,
                  const float* iobs, const float* sig, const int64_t* harmonic_id, int& order, int image_tile) {
  char buf[256];
  if (n <= 0 || !refl_id || !metadata || !iobs || !sig) { err = "clb_set_observations: null/empty input"; return 1; }
  if (laue && !harmonic_id) { err = "Laue model needs harmonic_id"; return 1; }
  if (n >= ((int64_t)1 << 31) - 64) { snprintf(buf, sizeof buf, "n_rows %lld exceeds 2^31 per handle", (long long)n); err = buf; return 1; }
  if (order == CLB_ORDER_AUTO) order = laue ? CLB_ORDER_SPOT : (image_tile > 0 ? CLB_ORDER_IMAGE : CLB_ORDER_REFL);
  if (image_tile > 0 && !image_id) { err = "image layers need image_id"; return 1; }
  if (image_tile > 0 && order != CLB_ORDER_SPOT && order != CLB_ORDER_IMAGE) { err = "image layers need image-major rows (CLB_ORDER_IMAGE or CLB_ORDER_SPOT)"; return 1; }
  if (laue && order != CLB_ORDER_SPOT) { err = "Laue rows must use CLB_ORDER_SPOT"; return 1; }
  if (!laue && order == CLB_ORDER_SPOT) { err = "CLB_ORDER_SPOT needs a Laue model"; return 1; }
  if (order == CLB_ORDER_IMAGE && !image_id) { err = "CLB_ORDER_IMAGE needs image_id"; return 1; }
  return 0;
}

// The first problem of input row i, in the order the checks are made (empty string: the row is fine).
std::string check_row(int64_t i, int64_t n, int64_t n_total, int64_t R, int n_images, int laue, const int64_t* refl_id,
                      const int64_t* image_id, const int64_t* harmonic_id, const int64_t* obs_index) {
  char buf[256]; buf[0] = 0;
  if (refl_id[i] < 0 || refl_id[i] >= R) snprintf(buf, sizeof buf, "refl_id[%lld]=%lld outside [0,%lld)", (long long)i, (long long)refl_id[i], (long long)R);
  else if (n_images > 0 && image_id && (image_id[i] < 0 || image_id[i] >= n_images)) snprintf(buf, sizeof buf, "image_id[%lld]=%lld outside [0,%lld)", (long long)i, (long long)image_id[i], (long long)n_images);
  else if (obs_index && (obs_index[i] < 0 || obs_index[i] >= n_total)) snprintf(buf, sizeof buf, "obs_index[%lld]=%lld outside [0,%lld)", (long long)i, (long long)obs_index[i], (long long)n_total);
  else if (laue && (harmonic_id[i] < 0 || harmonic_id[i] >= n)) snprintf(buf, sizeof buf, "harmonic_id[%lld]=%lld outside [0,%lld)", (long long)i, (long long)harmonic_id[i], (long long)n);
  return buf;
}

// Start position of every key's run of rows in the padded layout (kpos), from the run lengths.  The padding rules are a
// sequential recurrence over KEYS: a Laue spot must not straddle a 32-row warp chunk, an image (image layers) starts a new tile.
// `len(k)` = rows of key k, `first_image(k)` = image of the key's first row (Laue + image layers).  Returns the padded row count
// in npad (before the final rounding to 32); SPOT order also collects the empty slots and their constant log-likelihood.
template <class Len, class FirstImage>
int plan_key_positions(std::string& err, RowPlan& plan, int order, int64_t n_keys, int image_tile, int likelihood, float dof,
                       const float* iobs, const float* sig, Len len_of, FirstImage first_image, uint32_t* kpos, int64_t& npad) {
  char buf[256];
  int64_t p = 0, prev_img = -1;
  if (order == CLB_ORDER_SPOT) {
    clb_config lc{}; lc.likelihood = likelihood; lc.dof = dof;
    std::vector<float> empty_i, empty_s;
    for (int64_t k = 0; k < n_keys; ++k) {
      const int64_t len = len_of(k);
      kpos[k] = (uint32_t)p;
      if (len == 0) { plan.ll_const += lik_logpdf_zero(lc, iobs[k], sig[k]); empty_i.push_back(iobs[k]); empty_s.push_back(sig[k]); continue; }
      if (len > 32) { snprintf(buf, sizeof buf, "spot %lld has %lld harmonics; at most 32 are supported", (long long)k, (long long)len); err = buf; return 1; }
      if (image_tile > 0) {                          // harmonic ids are image-major (formatter.py:617)
        const int64_t img = first_image(k);
        if (img < prev_img) { snprintf(buf, sizeof buf, "image layers need harmonic_id to be image-major (spot %lld)", (long long)k); err = buf; return 1; }
        if (img != prev_img) p = (p + image_tile - 1) / image_tile * image_tile;
        prev_img = img;
      }
      if ((p & 31) + len > 32) p = (p + 31) & ~(int64_t)31;
      kpos[k] = (uint32_t)p;
      p += len;
    }
    plan.empty = empty_i;
    plan.empty.insert(plan.empty.end(), empty_s.begin(), empty_s.end());
  } else {                                           // CLB_ORDER_IMAGE with image layers: every image starts a new tile
    for (int64_t k = 0; k < n_keys; ++k) {
      const int64_t len = len_of(k);
      if (len > 0) p = (p + image_tile - 1) / image_tile * image_tile;
      kpos[k] = (uint32_t)p;
      p += len;
    }
  }
  npad = p;
  return 0;
}

int plan_rows(std::string& err, RowPlan& plan, int64_t n, int64_t n_total, int64_t R, int n_images, int laue,
              int likelihood, float dof, const int64_t* refl_id, const int64_t* image_id, const float* metadata,
              const float* iobs, const float* sig, const int64_t* harmonic_id, const int64_t* obs_index, int order,
              int image_tile = 0) {
  // image_tile > 0 (image layers): rows are image-major and no image may straddle a tile of image_tile rows
  if (resolve_order(err, n, laue, refl_id, image_id, metadata, iobs, sig, harmonic_id, order, image_tile)) return 1;
  plan.order = order;
  for (int64_t i = 0; i < n; ++i) {
    const bool ok = !(refl_id[i] < 0 || refl_id[i] >= R) && !(n_images > 0 && image_id && (image_id[i] < 0 || image_id[i] >= n_images)) &&
                    !(obs_index && (obs_index[i] < 0 || obs_index[i] >= n_total)) && !(laue && (harmonic_id[i] < 0 || harmonic_id[i] >= n));
    if (!ok) { err = check_row(i, n, n_total, R, n_images, laue, refl_id, image_id, harmonic_id, obs_index); return 1; }
  }
  std::vector<int32_t> key;
  int64_t n_keys = 0;
  if (order == CLB_ORDER_REFL) { key.resize(n); for (int64_t i = 0; i < n; ++i) key[i] = (int32_t)refl_id[i]; n_keys = R; }
  else if (order == CLB_ORDER_SPOT) { key.resize(n); for (int64_t i = 0; i < n; ++i) key[i] = (int32_t)harmonic_id[i]; n_keys = n; }
  else if (order == CLB_ORDER_IMAGE) {
    int64_t mx = 0; for (int64_t i = 0; i < n; ++i) mx = std::max(mx, image_id[i]);
    key.resize(n); for (int64_t i = 0; i < n; ++i) key[i] = (int32_t)image_id[i]; n_keys = mx + 1;
  }
  plan.perm.resize(n);
  std::vector<int64_t> count;
  if (n_keys > 0) {
    count.assign((size_t)n_keys + 1, 0);
    for (int64_t i = 0; i < n; ++i) count[key[i] + 1]++;
    for (int64_t k = 0; k < n_keys; ++k) count[k + 1] += count[k];
    std::vector<int64_t> cur(count.begin(), count.end() - 1);
    for (int64_t i = 0; i < n; ++i) plan.perm[cur[key[i]]++] = (int32_t)i;
  } else {
    for (int64_t i = 0; i < n; ++i) plan.perm[i] = (int32_t)i;
  }
  plan.pos.resize(n);
  plan.ll_const = 0.0;
  int64_t npad = n;
  if (order == CLB_ORDER_SPOT || (order == CLB_ORDER_IMAGE && image_tile > 0)) {
    std::vector<uint32_t> kpos((size_t)n_keys);
    if (plan_key_positions(err, plan, order, n_keys, image_tile, likelihood, dof, iobs, sig,
                           [&](int64_t k) { return count[k + 1] - count[k]; },
                           [&](int64_t k) { return image_id[plan.perm[count[k]]]; }, kpos.data(), npad)) return 1;
    for (int64_t k = 0; k < n_keys; ++k)
      for (int64_t j = count[k]; j < count[k + 1]; ++j) plan.pos[j] = (int64_t)kpos[k] + (j - count[k]);
  } else {
    for (int64_t i = 0; i < n; ++i) plan.pos[i] = i;
  }
  plan.npad = (npad + 31) & ~(int64_t)31;
  return 0;
}

void fill_rows(const RowPlan& plan, int64_t n, int d, const int64_t* refl_id, const int64_t* image_id,
               const float* metadata, const float* iobs, const float* sig, const int64_t* harmonic_id,
               const int64_t* obs_index, int32_t* p_refl, int32_t* p_img, int32_t* p_spot, uint32_t* p_oidx,
               float* p_meta, float* p_iobs, float* p_sig) {
  const int64_t npad = plan.npad;
  for (int64_t i = 0; i < npad; ++i) { p_refl[i] = -1; p_oidx[i] = 0; p_iobs[i] = 0.f; p_sig[i] = 1.f; }
  if (p_img) std::fill(p_img, p_img + npad, 0);
  if (p_spot) std::fill(p_spot, p_spot + npad, -1);
  std::fill(p_meta, p_meta + (size_t)npad * d, 0.f);
  for (int64_t sidx = 0; sidx < n; ++sidx) {
    const int64_t i = plan.perm[sidx], p = plan.pos[sidx];
    p_refl[p] = (int32_t)refl_id[i];
    if (p_img) p_img[p] = (int32_t)image_id[i];
    p_oidx[p] = (uint32_t)(obs_index ? obs_index[i] : i);
    for (int j = 0; j < d; ++j) p_meta[(size_t)j * npad + p] = metadata[(size_t)i * d + j];
    if (p_spot) {
      const int64_t k = harmonic_id[i];
      p_spot[p] = (int32_t)k; p_iobs[p] = iobs[k]; p_sig[p] = sig[k];     // formatter.py:637-640: spot k's value sits at index k
    } else { p_iobs[p] = iobs[i]; p_sig[p] = sig[i]; }
  }
}

// Exclusive prefix sum of m uint32 on the device, in place (recursive over 2 048-element tiles; tmp: m / 2047 + 16 elements).
cudaError_t device_exclusive_scan(uint32_t* data, int64_t m, uint32_t* tmp, cudaStream_t st) {
  if (m <= 0) return cudaSuccess;
  const int64_t nt = (m + prep::kScanTile - 1) / prep::kScanTile;
  prep::k_scan_tile<<<(unsigned)nt, prep::kScanThreads, 0, st>>>(data, m, tmp);
  if (nt > 1) {
    cudaError_t e = device_exclusive_scan(tmp, nt, tmp + nt, st);
    if (e != cudaSuccess) return e;
    prep::k_scan_add<<<(unsigned)((m + 255) / 256), 256, 0, st>>>(data, m, tmp);
  }
  return cudaGetLastError();
}

struct RowLayout { size_t bytes = 0, o_refl = 0, o_img = 0, o_spot = 0, o_oidx = 0, o_meta = 0, o_iobs = 0, o_sig = 0; };
RowLayout carve_rows(int64_t npad, int d, bool has_img, bool has_spot) {
  RowLayout L;
  auto carve = [&](size_t nbytes) { size_t o = L.bytes; L.bytes += (nbytes + 255) & ~(size_t)255; return o; };
  L.o_refl = carve(sizeof(int32_t) * npad);
  L.o_img = has_img ? carve(sizeof(int32_t) * npad) : 0;
  L.o_spot = has_spot ? carve(sizeof(int32_t) * npad) : 0;
  L.o_oidx = carve(sizeof(uint32_t) * npad);
  L.o_meta = carve(sizeof(float) * npad * d);
  L.o_iobs = carve(sizeof(float) * npad);
  L.o_sig = carve(sizeof(float) * npad);
  return L;
}

// Device-side version of plan_rows + fill_rows (clb_prep.cuh): the raw tuple is copied to the GPU, checked, sorted (stable LSD
// radix sort), padded and gathered there; only the O(n_keys) padding recurrence of Laue spots / image tiles runs on the host.
// On success h->rows holds the rows (layout L) and plan carries order, npad, ll_const and the empty Laue slots.
int prep_rows_device(clb_handle* h, RowPlan& plan, RowLayout& L, int64_t n, int64_t n_total, int n_images, int image_tile,
                     const int64_t* refl_id, const int64_t* image_id, const float* metadata, const float* iobs, const float* sig,
                     const int64_t* harmonic_id, const int64_t* obs_index, int order) {
  const clb_config& c = h->cfg;
  std::string err;
  if (resolve_order(err, n, c.laue, refl_id, image_id, metadata, iobs, sig, harmonic_id, order, image_tile)) return fail(h, CLB_ERR_INVALID, "%s", err.c_str());
  plan.order = order; plan.ll_const = 0.0;
  cudaStream_t st = h->stream;
  const int d = c.n_meta;
  const bool has_img = image_id != nullptr, has_spot = c.laue != 0;
  const int64_t n_keys = order == CLB_ORDER_REFL ? h->R : order == CLB_ORDER_SPOT ? n : (int64_t)n_images;
  const bool need_runs = order == CLB_ORDER_SPOT || (order == CLB_ORDER_IMAGE && image_tile > 0);
  // ---- raw tuple -> device ----
  DevBuf r_refl, r_img, r_harm, r_oidx, r_meta, r_iobs, r_sig, key[2], idx[2], G, tmp, count, off, kposb, posb, bad, kimgb;
  auto up = [&](DevBuf& b, const void* src, size_t bytes) -> cudaError_t {
    cudaError_t e = b.alloc(bytes);
    return e != cudaSuccess ? e : cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyHostToDevice, st);
  };
  CLB_CUDA(h, up(r_refl, refl_id, sizeof(int64_t) * n));
  if (has_img) CLB_CUDA(h, up(r_img, image_id, sizeof(int64_t) * n));
  if (has_spot) CLB_CUDA(h, up(r_harm, harmonic_id, sizeof(int64_t) * n));
  if (obs_index) CLB_CUDA(h, up(r_oidx, obs_index, sizeof(int64_t) * n));
  CLB_CUDA(h, up(r_meta, metadata, sizeof(float) * (size_t)n * d));
  CLB_CUDA(h, up(r_iobs, iobs, sizeof(float) * n));
  CLB_CUDA(h, up(r_sig, sig, sizeof(float) * n));
  // ---- checks, keys, run lengths ----
  CLB_CUDA(h, key[0].alloc(sizeof(uint32_t) * n)); CLB_CUDA(h, key[1].alloc(sizeof(uint32_t) * n));
  CLB_CUDA(h, idx[0].alloc(sizeof(uint32_t) * n)); CLB_CUDA(h, idx[1].alloc(sizeof(uint32_t) * n));
  CLB_CUDA(h, bad.alloc(sizeof(unsigned long long)));
  CLB_CUDA(h, cudaMemsetAsync(bad.p, 0xff, sizeof(unsigned long long), st));
  if (need_runs) { CLB_CUDA(h, count.alloc(sizeof(uint32_t) * n_keys)); CLB_CUDA(h, cudaMemsetAsync(count.p, 0, sizeof(uint32_t) * n_keys, st)); }
  const unsigned gkeys = (unsigned)std::min<int64_t>((n + 255) / 256, (int64_t)h->n_sms * 16);
  prep::k_prep_keys<<<gkeys, 256, 0, st>>>(n, r_refl.as<int64_t>(), has_img ? r_img.as<int64_t>() : nullptr, has_spot ? r_harm.as<int64_t>() : nullptr,
                                           obs_index ? r_oidx.as<int64_t>() : nullptr, h->R, (int64_t)n_images, n_total, c.laue, order,
                                           key[0].as<uint32_t>(), bad.as<unsigned long long>(), need_runs ? count.as<uint32_t>() : nullptr);
  CLB_LAUNCHED(h);
  unsigned long long first_bad = prep::kNoBadRow;
  CLB_CUDA(h, cudaMemcpyAsync(&first_bad, bad.p, sizeof first_bad, cudaMemcpyDeviceToHost, st));
  CLB_CUDA(h, cudaStreamSynchronize(st));
  if (first_bad != prep::kNoBadRow)
    return fail(h, CLB_ERR_INVALID, "%s", check_row((int64_t)first_bad, n, n_total, h->R, n_images, c.laue, refl_id, image_id, harmonic_id, obs_index).c_str());
  // ---- stable LSD radix sort of (key, row) ----
  int bits = 1; while (bits < 32 && ((int64_t)1 << bits) < n_keys) ++bits;
  const int passes = (bits + 7) / 8;
  const int nblocks = (int)((n + prep::kRsTile - 1) / prep::kRsTile);
  const int64_t gsize = 256 * (int64_t)nblocks;
  CLB_CUDA(h, G.alloc(sizeof(uint32_t) * gsize));
  CLB_CUDA(h, tmp.alloc(sizeof(uint32_t) * (std::max<int64_t>(gsize, n_keys) / (prep::kScanTile - 1) + 64)));
  int cur = 0;
  for (int pass = 0; pass < passes; ++pass) {
    prep::k_rs_hist<<<nblocks, prep::kRsThreads, 0, st>>>(key[cur].as<uint32_t>(), n, 8 * pass, G.as<uint32_t>(), nblocks);
    CLB_LAUNCHED(h);
    CLB_CUDA(h, device_exclusive_scan(G.as<uint32_t>(), gsize, tmp.as<uint32_t>(), st));
    prep::k_rs_scatter<<<nblocks, prep::kRsThreads, 0, st>>>(key[cur].as<uint32_t>(), pass == 0 ? nullptr : idx[cur].as<uint32_t>(), n, 8 * pass,
                                                             G.as<uint32_t>(), nblocks, key[cur ^ 1].as<uint32_t>(), idx[cur ^ 1].as<uint32_t>());
    CLB_LAUNCHED(h);
    cur ^= 1;
  }
  const uint32_t* skey = key[cur].as<uint32_t>();
  const uint32_t* perm = idx[cur].as<uint32_t>();
  // ---- padded positions ----
  int64_t npad = n;
  const uint32_t* pos = nullptr;
  if (need_runs) {
    CLB_CUDA(h, off.alloc(sizeof(uint32_t) * n_keys));
    CLB_CUDA(h, cudaMemcpyAsync(off.p, count.p, sizeof(uint32_t) * n_keys, cudaMemcpyDeviceToDevice, st));
    CLB_CUDA(h, device_exclusive_scan(off.as<uint32_t>(), n_keys, tmp.as<uint32_t>(), st));
    std::vector<uint32_t> cnt_h((size_t)n_keys), kpos_h((size_t)n_keys);
    std::vector<int32_t> kimg_h;
    CLB_CUDA(h, cudaMemcpyAsync(cnt_h.data(), count.p, sizeof(uint32_t) * n_keys, cudaMemcpyDeviceToHost, st));
    if (order == CLB_ORDER_SPOT && image_tile > 0) {
      kimg_h.resize((size_t)n_keys);
      CLB_CUDA(h, kimgb.alloc(sizeof(int32_t) * n_keys));
      prep::k_prep_first_image<<<(unsigned)((n_keys + 255) / 256), 256, 0, st>>>(n_keys, count.as<uint32_t>(), off.as<uint32_t>(), perm, r_img.as<int64_t>(), kimgb.as<int32_t>());
      CLB_LAUNCHED(h);
      CLB_CUDA(h, cudaMemcpyAsync(kimg_h.data(), kimgb.p, sizeof(int32_t) * n_keys, cudaMemcpyDeviceToHost, st));
    }
    CLB_CUDA(h, cudaStreamSynchronize(st));
    if (plan_key_positions(err, plan, order, n_keys, image_tile, c.likelihood, c.dof, iobs, sig,
                           [&](int64_t k) { return (int64_t)cnt_h[k]; }, [&](int64_t k) { return (int64_t)kimg_h[k]; }, kpos_h.data(), npad))
      return fail(h, CLB_ERR_INVALID, "%s", err.c_str());
    CLB_CUDA(h, kposb.alloc(sizeof(uint32_t) * n_keys)); CLB_CUDA(h, posb.alloc(sizeof(uint32_t) * n));
    CLB_CUDA(h, cudaMemcpyAsync(kposb.p, kpos_h.data(), sizeof(uint32_t) * n_keys, cudaMemcpyHostToDevice, st));
    prep::k_prep_pos<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, skey, off.as<uint32_t>(), kposb.as<uint32_t>(), posb.as<uint32_t>());
    CLB_LAUNCHED(h);
    CLB_CUDA(h, cudaStreamSynchronize(st));          // kpos_h goes out of scope
    pos = posb.as<uint32_t>();
  }
  npad = (npad + 31) & ~(int64_t)31;
  plan.npad = npad;
  // ---- the rows ----
  L = carve_rows(npad, d, has_img, has_spot);
  CLB_CUDA(h, h->rows.alloc(L.bytes));
  char* db = h->rows.as<char>();
  int32_t* o_img = has_img ? reinterpret_cast<int32_t*>(db + L.o_img) : nullptr;
  int32_t* o_spot = has_spot ? reinterpret_cast<int32_t*>(db + L.o_spot) : nullptr;
  prep::k_fill_defaults<<<(unsigned)((npad + 255) / 256), 256, 0, st>>>(npad, d, reinterpret_cast<int32_t*>(db + L.o_refl), o_img, o_spot,
      reinterpret_cast<uint32_t*>(db + L.o_oidx), reinterpret_cast<float*>(db + L.o_meta), reinterpret_cast<float*>(db + L.o_iobs), reinterpret_cast<float*>(db + L.o_sig));
  CLB_LAUNCHED(h);
  prep::k_fill_gather<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, npad, d, perm, pos, r_refl.as<int64_t>(), has_img ? r_img.as<int64_t>() : nullptr,
      has_spot ? r_harm.as<int64_t>() : nullptr, obs_index ? r_oidx.as<int64_t>() : nullptr, r_meta.as<float>(), r_iobs.as<float>(), r_sig.as<float>(),
      reinterpret_cast<int32_t*>(db + L.o_refl), o_img, o_spot, reinterpret_cast<uint32_t*>(db + L.o_oidx), reinterpret_cast<float*>(db + L.o_meta),
      reinterpret_cast<float*>(db + L.o_iobs), reinterpret_cast<float*>(db + L.o_sig));
  CLB_LAUNCHED(h);
  CLB_CUDA(h, cudaStreamSynchronize(st));            // the temporaries are freed on return
  return CLB_OK;
}

// The pinned host mirror of the prepared rows (re-upload / prefetch pipeline): after a device-side preparation it is filled on demand.
int ensure_rows_host(clb_handle* h) {
  if (h->rows_host_valid) return CLB_OK;
  CLB_CUDA(h, h->rows_host.alloc(h->rows_bytes));
  CLB_CUDA(h, cudaMemcpyAsync(h->rows_host.p, h->rows.p, h->rows_bytes, cudaMemcpyDeviceToHost, h->stream));
  CLB_CUDA(h, cudaStreamSynchronize(h->stream));
  h->rows_host_valid = true;
  return CLB_OK;
}

}  // namespace

extern "C" {

int clb_abi_version(void) { return CLB_ABI_VERSION; }

const char* clb_last_error(const clb_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int clb_create(const clb_config* cfg, clb_handle** out) {
  if (!cfg || !out) return fail(nullptr, CLB_ERR_INVALID, "clb_create: null argument");
  *out = nullptr;
  if (cfg->abi_version != CLB_ABI_VERSION) return fail(nullptr, CLB_ERR_INVALID, "ABI version mismatch: got %d, library is %d", cfg->abi_version, CLB_ABI_VERSION);
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 || cfg->device >= ndev)
    return fail(nullptr, CLB_ERR_NO_DEVICE, "no CUDA device %d (found %d); careless_b200 has no CPU fallback", cfg->device, ndev);
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, cfg->device) != cudaSuccess) return fail(nullptr, CLB_ERR_CUDA, "cudaGetDeviceProperties failed");
  if (prop.major != 10) return fail(nullptr, CLB_ERR_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a (B200) only", cfg->device, prop.major, prop.minor);
  if (cfg->n_refl <= 0 || cfg->n_meta <= 0 || cfg->mlp_layers < 0 || cfg->mlp_width <= 0 || cfg->mc_samples <= 0)
    return fail(nullptr, CLB_ERR_INVALID, "invalid sizes: n_refl=%lld n_meta=%d mlp_layers=%d mlp_width=%d mc_samples=%d",
                (long long)cfg->n_refl, cfg->n_meta, cfg->mlp_layers, cfg->mlp_width, cfg->mc_samples);
  if (cfg->mlp_layers + 1 > kMaxLayers) return fail(nullptr, CLB_ERR_INVALID, "mlp_layers %d exceeds the supported maximum %d", cfg->mlp_layers, kMaxLayers - 1);
  if (cfg->likelihood == CLB_LIK_STUDENTT && !(cfg->dof > 0.f)) return fail(nullptr, CLB_ERR_INVALID, "student-t likelihood needs dof > 0");
  if (cfg->image_scales && cfg->n_images <= 0) return fail(nullptr, CLB_ERR_INVALID, "image scales need n_images > 0");
  if (cfg->image_layers < 0 || (cfg->image_layers > 0 && cfg->n_images <= 0)) return fail(nullptr, CLB_ERR_INVALID, "image layers need n_images > 0");
  if (cfg->image_layers > 0 && cfg->mlp_layers == 0 && cfg->n_meta != cfg->mlp_width)
    return fail(nullptr, CLB_ERR_INVALID, "image layers without MLP layers need n_meta == mlp_width");
  if (cfg->mlp_layers + cfg->image_layers + 1 > kMaxLayers) return fail(nullptr, CLB_ERR_INVALID, "too many layers");
  if (cfg->prior == CLB_PRIOR_DOUBLE_WILSON && cfg->n_asu <= 0) return fail(nullptr, CLB_ERR_INVALID, "DoubleWilson needs n_asu > 0");
  if (cfg->deterministic && (cfg->image_scales || cfg->image_layers > 0 || cfg->refine_uncertainties || cfg->prior != CLB_PRIOR_WILSON))
    return fail(nullptr, CLB_ERR_INVALID, "deterministic mode supports MLPScaler models with the Wilson prior (no image scales / image layers / "
                                          "refined uncertainties / DoubleWilson: their gradients are accumulated with atomics)");
  const int wmax = std::max(std::max(cfg->n_meta, cfg->mlp_width), 2);
  const int WP = round_width(wmax);
  if (WP < 0) return fail(nullptr, CLB_ERR_INVALID, "max(n_meta, mlp_width) = %d exceeds the supported width 64", wmax);

  clb_handle* h = new clb_handle();
  h->cfg = *cfg;
  if (h->cfg.n_refl_total <= 0) h->cfg.n_refl_total = h->cfg.n_refl;
  if (h->cfg.world_size <= 0) { h->cfg.world_size = 1; h->cfg.rank = 0; }
  { const char* dbg = getenv("CLB_DEBUG_SYNC"); h->debug_sync = dbg && dbg[0] == '1'; }
  { const char* dis = getenv("CLB_DISCARD"); h->discard_scratch = !(dis && dis[0] == '0'); }
  { const char* dp = getenv("CLB_DEVICE_PREP"); h->device_prep = !(dp && dp[0] == '0'); }
  { const char* fa = getenv("CLB_FUSED_ADAM"); h->no_fused_adam = fa && fa[0] == '0'; }
  h->det = cfg->deterministic != 0;
  h->R = cfg->n_refl; h->S = cfg->mc_samples; h->WP = WP; h->KS = 1;
  auto bail = [&](int code) { g_create_error = h->err; delete h; return code; };
#define CREATE_CUDA(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) { fail(h, CLB_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(e_)); return bail(CLB_ERR_CUDA); } } while (0)
  CREATE_CUDA(cudaSetDevice(cfg->device));
  h->n_sms = prop.multiProcessorCount;
  if (cfg->stream) h->stream = (cudaStream_t)cfg->stream;
  else { CREATE_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)); h->own_stream = true; }
  build_layout(h);
  build_vars(h);
  h->NL = h->lay.n_layers;
  { const char* no_tc = getenv("CLB_NO_TC"); h->use_tc = (WP == 32) && (cfg->mlp_layers + cfg->image_layers) > 0 && !(no_tc && no_tc[0] == '1'); }
  { const char* one = getenv("CLB_TC_ONE_THREAD_PER_ROW"); h->use_tc2 = h->use_tc && !(one && one[0] == '1'); }
  { const char* t16 = getenv("CLB_TC16"); const char* no_tc = getenv("CLB_NO_TC");
    h->use_tc16 = (WP <= 16) && cfg->mlp_layers > 0 && cfg->n_meta <= 16 &&
                  !(t16 && t16[0] == '0') && !(no_tc && no_tc[0] == '1'); }
  // k_obs_pp (one CTA per SM, two tiles ping-ponging through a dedicated issuer warp) is parity-green but measured slower than
  // k_obs_tc2 on B200 (18.7 vs 16.8 ms: two warps per scheduler cannot cover the ALU / TMEM latencies); opt-in with CLB_PP=1
  { const char* ppe = getenv("CLB_PP"); h->use_pp = h->use_tc2 && cfg->image_layers == 0 && cfg->mlp_layers > 0 && (ppe && ppe[0] == '1') && !cfg->deterministic; }
  { const char* bf = getenv("CLB_BIAS_FEAT"); h->bias_feat15 = h->use_tc16 && std::max(cfg->n_meta, cfg->mlp_width) <= 15 && !(bf && bf[0] == '0'); }
  { const char* t3 = getenv("CLB_TC3"); h->use_tc3 = h->use_tc2 && !h->use_pp && cfg->image_layers == 0 && cfg->mlp_layers > 0 && (t3 && t3[0] == '1') && !cfg->deterministic; }
  h->obs_threads = (h->use_tc || h->use_tc16) ? tc::kThreads : kObsThreads;      // = observation rows per CTA tile
  switch (WP) {
    case 8: h->smem_obs = ObsSmem<8>::bytes(h->NL, false, cfg->image_layers); break;
    case 16: h->smem_obs = ObsSmem<16>::bytes(h->NL, false, cfg->image_layers); break;
    case 64: h->smem_obs = ObsSmem<64>::bytes(h->NL, false, cfg->image_layers); break;
    default: h->smem_obs = ObsSmem<32>::bytes(h->NL, h->use_tc, cfg->image_layers); break;
  }
  if (h->use_tc2) h->smem_obs = ObsSmem2::bytes(h->NL, cfg->image_layers);
  if (h->use_tc16) h->smem_obs = ObsSmem16::bytes(h->NL, cfg->image_layers);
  if (h->use_pp) h->smem_obs = pp::SmemPP::bytes(h->NL);
  if (h->use_tc3) h->smem_obs = tc3::Smem3::bytes(h->NL);
  if (h->smem_obs > (size_t)prop.sharedMemPerBlockOptin) {
    fail(h, CLB_ERR_INVALID, "scale MLP (%d layers x padded width %d) needs %zu B of shared memory; the device offers %zu",
         cfg->mlp_layers, WP, h->smem_obs, (size_t)prop.sharedMemPerBlockOptin);
    return bail(CLB_ERR_INVALID);
  }
  const size_t pb = sizeof(float) * (size_t)h->P;
  CREATE_CUDA(h->theta.alloc(pb)); CREATE_CUDA(h->m.alloc(pb)); CREATE_CUDA(h->v.alloc(pb)); CREATE_CUDA(h->grad.alloc(pb));
  CREATE_CUDA(cudaMemsetAsync(h->m.p, 0, pb, h->stream));
  CREATE_CUDA(cudaMemsetAsync(h->v.p, 0, pb, h->stream));
  CREATE_CUDA(cudaMemsetAsync(h->grad.p, 0, pb, h->stream));
  CREATE_CUDA(h->z.alloc(sizeof(float) * h->R * h->S));
  CREATE_CUDA(h->gz.alloc(sizeof(float) * h->R * h->S));
  CREATE_CUDA(h->acc.alloc(sizeof(double) * ACC_COUNT));
  CREATE_CUDA(h->var_sums.alloc(sizeof(double) * 2 * kMaxVars));
  CREATE_CUDA(h->red.alloc(sizeof(double) * (2 + 2 * kMaxVars)));
  CREATE_CUDA(h->var_scale.alloc(sizeof(float) * kMaxVars));
  CREATE_CUDA(h->adam_alpha.alloc(sizeof(float) * 4));
  CREATE_CUDA(h->stop_step.alloc(sizeof(int) * 4));
  // reference initial values: identity kernels, zero biases (nn.py:55-79), image scales 1 (image.py:21)
  std::vector<float> init((size_t)h->P, 0.f);
  float* mlp = init.data() + h->goff[CLB_GROUP_MLP];
  for (int k = 0; k < h->lay.n_layers; ++k)
    for (int i = 0; i < std::min(h->lay.in_dim[k], h->lay.out_dim[k]); ++i) mlp[h->lay.koff[k] + i * h->lay.out_dim[k] + i] = 1.f;
  for (int64_t i = 0; i < h->gsize[CLB_GROUP_IMAGE_SCALES]; ++i) init[h->goff[CLB_GROUP_IMAGE_SCALES] + i] = 1.f;
  for (int l = 0; l < cfg->image_layers; ++l) {            // image.py:73-80: eye(units, in) for every image
    const int64_t w = cfg->mlp_width, stride = (int64_t)cfg->n_images * w * (w + 1);
    float* kern = init.data() + h->goff[CLB_GROUP_IMAGE_LAYERS] + l * stride;
    for (int64_t im = 0; im < cfg->n_images; ++im) for (int64_t j = 0; j < w; ++j) kern[(im * w + j) * w + j] = 1.f;
  }
  for (int64_t i = 0; i < h->gsize[CLB_GROUP_LIKELIHOOD]; ++i)       // softplus^-1(1): TransformedVariable(1., Softplus) (mono.py:42-44)
    init[h->goff[CLB_GROUP_LIKELIHOOD] + i] = 0.54132485f;
  CREATE_CUDA(cudaMemcpyAsync(h->theta.p, init.data(), pb, cudaMemcpyHostToDevice, h->stream));
  const int big = 0x7fffffff;
  CREATE_CUDA(cudaMemcpyAsync(h->stop_step.p, &big, sizeof(int), cudaMemcpyHostToDevice, h->stream));
  CREATE_CUDA(cudaStreamSynchronize(h->stream));
#undef CREATE_CUDA
  *out = h;
  return CLB_OK;
}

void clb_destroy(clb_handle* h) {
  if (!h) return;
  cudaSetDevice(h->cfg.device);
  cudaStreamSynchronize(h->stream);
  delete h;
}

int clb_synchronize(clb_handle* h) {
  if (!h) return CLB_ERR_INVALID;
  CLB_CUDA(h, cudaStreamSynchronize(h->stream));
  return CLB_OK;
}

static void point_rows(clb_handle* h, char* db) {
  h->d_refl = reinterpret_cast<int32_t*>(db + h->row_off[0]);
  h->d_image = h->has_img_rows ? reinterpret_cast<int32_t*>(db + h->row_off[1]) : nullptr;
  h->d_spot = h->has_spot_rows ? reinterpret_cast<int32_t*>(db + h->row_off[2]) : nullptr;
  h->d_oidx = reinterpret_cast<uint32_t*>(db + h->row_off[3]);
  h->d_meta = reinterpret_cast<float*>(db + h->row_off[4]);
  h->d_iobs = reinterpret_cast<float*>(db + h->row_off[5]);
  h->d_sig = reinterpret_cast<float*>(db + h->row_off[6]);
}

int clb_set_observations(clb_handle* h, int64_t n, int64_t n_total, const int64_t* refl_id, const int64_t* image_id,
                         const float* metadata, const float* iobs, const float* sig,
                         const int64_t* harmonic_id, const int64_t* obs_index, int32_t order) {
  if (!h) return CLB_ERR_INVALID;
  const clb_config& c = h->cfg;
  if ((c.image_scales || c.image_layers > 0) && !image_id) return fail(h, CLB_ERR_INVALID, "image scales / image layers need image_id");
  if (n_total <= 0) n_total = n;
  RowPlan plan;
  std::string err;
  CLB_CUDA(h, cudaSetDevice(c.device));
  const int d = c.n_meta;
  const bool has_img = image_id != nullptr;
  const bool has_spot = c.laue != 0;
  const int n_images_chk = (c.image_scales || c.image_layers > 0) ? c.n_images : 0;
  const int image_tile = c.image_layers > 0 ? h->obs_threads : 0;
  RowLayout L;
  if (h->pending_swap || h->alt_used) { CLB_CUDA(h, cudaStreamSynchronize(h->copy_stream)); h->pending_swap = false; h->alt_used = false; }
  h->rows_alt.free();
  // device-side preparation (default): everything but the deterministic mode (which also wants the CSR of row positions per
  // reflection, built from the host permutation) and a caller-forced image order without a known image count
  const bool on_device = h->device_prep && !h->det && n > 0 && !(order == CLB_ORDER_IMAGE && n_images_chk == 0) &&
                         order >= CLB_ORDER_AUTO && order <= CLB_ORDER_IMAGE;        // CLB_ORDER_NONE (rows kept as given): nothing to sort
  const auto t_prep0 = std::chrono::steady_clock::now();
  if (on_device) {
    const int rc = prep_rows_device(h, plan, L, n, n_total, n_images_chk, image_tile, refl_id, image_id, metadata, iobs, sig, harmonic_id, obs_index, order);
    if (rc != CLB_OK) return rc;
    h->rows_host_valid = false;
  } else {
    if (plan_rows(err, plan, n, n_total, h->R, n_images_chk, c.laue, c.likelihood, c.dof,
                  refl_id, image_id, metadata, iobs, sig, harmonic_id, obs_index, order, image_tile))
      return fail(h, CLB_ERR_INVALID, "%s", err.c_str());
    L = carve_rows(plan.npad, d, has_img, has_spot);
    CLB_CUDA(h, h->rows_host.alloc(L.bytes));
    CLB_CUDA(h, h->rows.alloc(L.bytes));
    char* hb = h->rows_host.as<char>();
    fill_rows(plan, n, d, refl_id, image_id, metadata, iobs, sig, harmonic_id, obs_index,
              reinterpret_cast<int32_t*>(hb + L.o_refl), has_img ? reinterpret_cast<int32_t*>(hb + L.o_img) : nullptr,
              has_spot ? reinterpret_cast<int32_t*>(hb + L.o_spot) : nullptr, reinterpret_cast<uint32_t*>(hb + L.o_oidx),
              reinterpret_cast<float*>(hb + L.o_meta), reinterpret_cast<float*>(hb + L.o_iobs), reinterpret_cast<float*>(hb + L.o_sig));
    h->rows_host_valid = true;
  }
  h->prep_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_prep0).count();
  const int64_t npad = plan.npad;
  const size_t bytes = L.bytes;
  const size_t o_refl = L.o_refl, o_img = L.o_img, o_spot = L.o_spot, o_oidx = L.o_oidx, o_meta = L.o_meta, o_iobs = L.o_iobs, o_sig = L.o_sig;
  h->ll_const = plan.ll_const;
  h->n_empty = 0;
  if (c.refine_uncertainties) {      // the empty slots' log-density depends on the error-model parameters: evaluated on the device
    h->ll_const = 0.0;
    h->n_empty = (int64_t)plan.empty.size() / 2;
    if (h->n_empty > 0) {
      CLB_CUDA(h, h->empty_slots.alloc(sizeof(float) * plan.empty.size()));
      CLB_CUDA(h, cudaMemcpyAsync(h->empty_slots.p, plan.empty.data(), sizeof(float) * plan.empty.size(), cudaMemcpyHostToDevice, h->stream));
      CLB_CUDA(h, cudaStreamSynchronize(h->stream));
    }
  }
  h->rows_bytes = bytes;
  h->row_off[0] = o_refl; h->row_off[1] = o_img; h->row_off[2] = o_spot; h->row_off[3] = o_oidx;
  h->row_off[4] = o_meta; h->row_off[5] = o_iobs; h->row_off[6] = o_sig;
  h->has_img_rows = has_img; h->has_spot_rows = has_spot;
  point_rows(h, h->rows.as<char>());
  h->n_rows_raw = n; h->n_rows = npad; h->n_rows_total = n_total; h->order = plan.order;

  // ---- launch geometry + per-CTA buffers of the observation kernel ----
  const int64_t n_tiles = (npad + h->obs_threads - 1) / h->obs_threads;
  h->grid_obs = (int)std::min<int64_t>(n_tiles, (int64_t)h->n_sms * (h->use_tc16 ? 4 : h->use_tc ? 2 : 1));
  if (h->use_pp) h->grid_obs = (int)std::min<int64_t>((n_tiles + 1) / 2, (int64_t)h->n_sms * pp::kCtasPerSM);     // two CTAs per SM, two tiles in flight each
  {
    // One allocation [weight-gradient partials | activation scratch] so that a single L2 access-policy window can cover
    // both: the partials are read-modify-written once per tile and layer and must not be evicted by the scratch
    // streaming through the same cache; whatever persisting capacity is left keeps the most recently written
    // activations on chip until the backward pass reads them (and discards them).  CLB_L2_WINDOW=0 no window, 1 partials
    // only, 2 both (default).  Measured on B200 (10 M observations, ncu dram bytes per launch of k_obs_tc2, discard on):
    // no window 7.96 GB, partials only 5.30 GB, partials + scratch 1.98 GB (5.6x the algorithmic 0.356 GB); time unchanged.
    // tensor-core kernels: the CTAs accumulate with REDs, so several of them can share one partial buffer -- a quarter as
    // many buffers as CTAs keeps the L2 footprint (and the persisting carve-out) small without measurable contention
    h->n_partials = (h->use_tc16 || h->use_tc2) ? std::max(1, h->grid_obs / 4) : h->grid_obs;
    if (h->det) h->n_partials = h->grid_obs;          // exclusive buffers: one writer per address
    const size_t pbytes = h->use_tc16 ? sizeof(float) * (size_t)h->n_partials * h->NL * (h->det ? tc16::PSLOT16_DET : tc16::PSLOT16)
                          : h->use_tc2 ? sizeof(float) * (size_t)h->n_partials * h->NL * (h->det ? tc::kPslotDet : 32 * 32 + 32)
                                       : sizeof(double) * (size_t)h->grid_obs * h->KS * partial_row_size(h->NL, h->WP);
    const size_t pb = (pbytes + 255) & ~(size_t)255;
    const size_t sbytes = sizeof(float4) * (size_t)h->grid_obs * ((h->use_pp || h->use_tc3) ? 2 : 1) * std::max(1, c.mlp_layers + c.image_layers) * ((h->use_tc16 ? 16 : h->WP) / 4) * h->obs_threads;
    CLB_CUDA(h, h->partials.alloc(pb + sbytes));
    h->partial_bytes = pbytes;
    h->scratch_ptr = reinterpret_cast<float4*>(h->partials.as<char>() + pb);
    const char* mode_s = getenv("CLB_L2_WINDOW");
    const int mode = mode_s ? atoi(mode_s) : 2;
    int max_persist = 0, max_window = 0;
    cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, c.device);
    cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, c.device);
    if (mode > 0 && max_persist > 0 && max_window > 0) {
      const size_t span = std::min<size_t>(mode >= 2 ? pb + sbytes : pb, (size_t)max_window);
      const size_t want = std::min<size_t>(span, (size_t)max_persist);
      cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want);
      cudaStreamAttrValue attr{};
      attr.accessPolicyWindow.base_ptr = h->partials.p;
      attr.accessPolicyWindow.num_bytes = span;
      attr.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)want / (double)span);
      attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
      attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
      if (cudaStreamSetAttribute(h->stream, cudaStreamAttributeAccessPolicyWindow, &attr) != cudaSuccess) cudaGetLastError();
    }
  }
  if (h->use_tc || h->WP == 64) CLB_CUDA(h, h->wpack.alloc(sizeof(float) * (size_t)std::max(1, c.mlp_layers) * h->WP * h->WP));
  if (h->use_tc2 || h->use_tc16) {     // [L][fwd, bwd][hi, lo][image bytes]; the padding bytes of the images stay zero
    const size_t nb = (size_t)std::max(1, c.mlp_layers) * (h->use_tc16 ? 6 * (size_t)tc16::kImg16 : 4 * (size_t)tc::kImgBytes);
    CLB_CUDA(h, h->wimg.alloc(nb));
    CLB_CUDA(h, cudaMemsetAsync(h->wimg.p, 0, nb, h->stream));
  }
  if (h->det) {
    // CSR of the padded row positions of every reflection, in row order: the fixed summation order of k_gz_reduce
    std::vector<int32_t> ptr((size_t)h->R + 1, 0), rows((size_t)n);
    for (int64_t sidx = 0; sidx < n; ++sidx) ptr[(size_t)refl_id[plan.perm[sidx]] + 1]++;
    for (int64_t r = 0; r < h->R; ++r) ptr[r + 1] += ptr[r];
    {
      std::vector<int32_t> cur(ptr.begin(), ptr.end() - 1);
      std::vector<std::pair<int64_t, int64_t>> order((size_t)n);      // (padded position, reflection): ascending position within a reflection
      for (int64_t sidx = 0; sidx < n; ++sidx) order[sidx] = {plan.pos[sidx], refl_id[plan.perm[sidx]]};
      std::sort(order.begin(), order.end());
      for (const auto& pr : order) rows[cur[pr.second]++] = (int32_t)pr.first;
    }
    CLB_CUDA(h, h->refl_ptr.alloc(sizeof(int32_t) * ptr.size())); CLB_CUDA(h, h->refl_rows.alloc(sizeof(int32_t) * rows.size()));
    CLB_CUDA(h, cudaMemcpyAsync(h->refl_ptr.p, ptr.data(), sizeof(int32_t) * ptr.size(), cudaMemcpyHostToDevice, h->stream));
    CLB_CUDA(h, cudaMemcpyAsync(h->refl_rows.p, rows.data(), sizeof(int32_t) * rows.size(), cudaMemcpyHostToDevice, h->stream));
    CLB_CUDA(h, h->dzf_rows.alloc(sizeof(float) * (size_t)npad * h->S));
    CLB_CUDA(h, h->ll_part.alloc(sizeof(double) * h->grid_obs));
    CLB_CUDA(h, cudaStreamSynchronize(h->stream));
  }
  h->have_obs = true;
  return on_device ? CLB_OK : clb_upload_observations(h);
}

// Host prep only (no CUDA): the sorted / padded SoA device layout of clb_set_observations, written to
// caller arrays of `capacity` rows.  *n_padded receives the padded row count; when capacity is too
// small nothing else is written.  Used by the CPU test-suite to check the layout bit-exactly.
int clb_prepare_rows(int64_t n, int64_t n_refl, int32_t n_meta, int32_t n_images, int32_t laue, int32_t likelihood, float dof,
                     const int64_t* refl_id, const int64_t* image_id, const float* metadata, const float* iobs,
                     const float* sig, const int64_t* harmonic_id, const int64_t* obs_index, int32_t order, int32_t image_tile,
                     int64_t capacity, int64_t* n_padded, int32_t* refl_out, int32_t* image_out, int32_t* spot_out,
                     uint32_t* oidx_out, float* meta_out, float* iobs_out, float* sig_out, double* ll_const) {
  RowPlan plan;
  std::string err;
  if (plan_rows(err, plan, n, n, n_refl, n_images, laue, likelihood, dof, refl_id, image_id, metadata, iobs, sig,
                harmonic_id, obs_index, order, image_tile))
    return fail(nullptr, CLB_ERR_INVALID, "%s", err.c_str());
  if (n_padded) *n_padded = plan.npad;
  if (ll_const) *ll_const = plan.ll_const;
  if (capacity < plan.npad) return CLB_OK;
  if (!refl_out || !oidx_out || !meta_out || !iobs_out || !sig_out) return fail(nullptr, CLB_ERR_INVALID, "clb_prepare_rows: null output");
  fill_rows(plan, n, n_meta, refl_id, image_id, metadata, iobs, sig, harmonic_id, obs_index,
            refl_out, image_id ? image_out : nullptr, laue ? spot_out : nullptr, oidx_out, meta_out, iobs_out, sig_out);
  return CLB_OK;
}

// Start the host -> device copy of the prepared rows into the OTHER device buffer on a copy stream and return at once;
// the next clb_step / clb_step_begin / clb_eval waits for the copy (device side) and switches to that buffer.  With one
// call per step this is a prefetching input pipeline: step t computes while the rows of step t + 1 arrive.
int clb_prefetch_observations(clb_handle* h) {
  if (!h || !h->have_obs) return fail(h, CLB_ERR_STATE, "clb_prefetch_observations before clb_set_observations");
  if (h->pending_swap) return fail(h, CLB_ERR_STATE, "clb_prefetch_observations: the previous prefetch has not been consumed by a step yet");
  CLB_CUDA(h, cudaSetDevice(h->cfg.device));
  if (!h->copy_stream) {
    CLB_CUDA(h, cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    CLB_CUDA(h, cudaEventCreateWithFlags(&h->ev_copy_done, cudaEventDisableTiming));
    CLB_CUDA(h, cudaEventCreateWithFlags(&h->ev_rows_free[0], cudaEventDisableTiming));
    CLB_CUDA(h, cudaEventCreateWithFlags(&h->ev_rows_free[1], cudaEventDisableTiming));
  }
  { const int rc = ensure_rows_host(h); if (rc != CLB_OK) return rc; }
  if (!h->rows_alt.p) CLB_CUDA(h, h->rows_alt.alloc(h->rows_bytes));
  // the target buffer was last read by the step before the current one; wait until that step's kernels are done
  if (h->alt_used) CLB_CUDA(h, cudaStreamWaitEvent(h->copy_stream, h->ev_rows_free[h->cur_rows ^ 1], 0));
  CLB_CUDA(h, cudaMemcpyAsync(h->rows_alt.p, h->rows_host.p, h->rows_bytes, cudaMemcpyHostToDevice, h->copy_stream));
  CLB_CUDA(h, cudaEventRecord(h->ev_copy_done, h->copy_stream));
  h->pending_swap = true;
  return CLB_OK;
}

int clb_upload_observations(clb_handle* h) {
  if (!h || !h->have_obs) return fail(h, CLB_ERR_STATE, "clb_upload_observations before clb_set_observations");
  CLB_CUDA(h, cudaSetDevice(h->cfg.device));
  { const int rc = ensure_rows_host(h); if (rc != CLB_OK) return rc; }
  CLB_CUDA(h, cudaMemcpyAsync(h->rows.p, h->rows_host.p, h->rows_bytes, cudaMemcpyHostToDevice, h->stream));
  return CLB_OK;
}

int clb_download_rows(clb_handle* h, int64_t capacity, int64_t* n_padded, int32_t* refl_out, int32_t* image_out, int32_t* spot_out,
                      uint32_t* oidx_out, float* meta_out, float* iobs_out, float* sig_out, double* ll_const, double* prep_ms) {
  if (!h || !h->have_obs) return fail(h, CLB_ERR_STATE, "clb_download_rows before clb_set_observations");
  if (n_padded) *n_padded = h->n_rows;
  if (ll_const) *ll_const = h->ll_const;
  if (prep_ms) *prep_ms = h->prep_ms;
  if (capacity < h->n_rows) return CLB_OK;
  if (!refl_out || !oidx_out || !meta_out || !iobs_out || !sig_out) return fail(h, CLB_ERR_INVALID, "clb_download_rows: null output");
  CLB_CUDA(h, cudaSetDevice(h->cfg.device));
  const size_t np = (size_t)h->n_rows;
  auto down = [&](void* dst, const void* src, size_t bytes) { return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, h->stream); };
  CLB_CUDA(h, down(refl_out, h->d_refl, sizeof(int32_t) * np));
  if (image_out && h->d_image) CLB_CUDA(h, down(image_out, h->d_image, sizeof(int32_t) * np));
  if (spot_out && h->d_spot) CLB_CUDA(h, down(spot_out, h->d_spot, sizeof(int32_t) * np));
  CLB_CUDA(h, down(oidx_out, h->d_oidx, sizeof(uint32_t) * np));
  CLB_CUDA(h, down(meta_out, h->d_meta, sizeof(float) * np * h->cfg.n_meta));
  CLB_CUDA(h, down(iobs_out, h->d_iobs, sizeof(float) * np));
  CLB_CUDA(h, down(sig_out, h->d_sig, sizeof(float) * np));
  CLB_CUDA(h, cudaStreamSynchronize(h->stream));
  return CLB_OK;
}

int clb_set_prior(clb_handle* h, const uint8_t* centric, const float* mult, const float* sigma,
                  const int32_t* dw_parent, const int32_t* asu_id, const float* r,
                  const int64_t* refl_index, float init_scale) {
  if (!h || !centric || !mult) return fail(h, CLB_ERR_INVALID, "clb_set_prior: null input");
  const clb_config& c = h->cfg;
  const bool dw = c.prior == CLB_PRIOR_DOUBLE_WILSON;
  if (dw && (!dw_parent || !asu_id || !r)) return fail(h, CLB_ERR_INVALID, "DoubleWilson needs dw_parent, asu_id and r");
  CLB_CUDA(h, cudaSetDevice(c.device));
  const int64_t R = h->R;
  std::vector<float> es((size_t)R);
  std::vector<uint32_t> ridx((size_t)R);
  for (int64_t i = 0; i < R; ++i) {
    es[i] = mult[i] * (sigma ? sigma[i] : 1.0f);
    if (!(es[i] > 0.f)) return fail(h, CLB_ERR_INVALID, "multiplicity*sigma must be positive (entry %lld)", (long long)i);
    ridx[i] = (uint32_t)(refl_index ? refl_index[i] : i);
  }
  if (dw) {
    for (int i = 0; i < c.n_asu; ++i)
      if (r[i] >= 1.f || r[i] <= -1.f) return fail(h, CLB_ERR_INVALID, "double-wilson r value %g outside of allowed range (-1, 1)", r[i]);   // manager.py:415-419
    for (int64_t i = 0; i < R; ++i) {
      if (dw_parent[i] < -2 || dw_parent[i] >= R) return fail(h, CLB_ERR_INVALID, "dw_parent[%lld] out of range", (long long)i);
      if (asu_id[i] < 0 || asu_id[i] >= c.n_asu) return fail(h, CLB_ERR_INVALID, "asu_id[%lld] out of range", (long long)i);
    }
  }
  CLB_CUDA(h, h->centric.alloc(R)); CLB_CUDA(h, h->eps_sigma.alloc(sizeof(float) * R)); CLB_CUDA(h, h->refl_index.alloc(sizeof(uint32_t) * R));
  CLB_CUDA(h, cudaMemcpyAsync(h->centric.p, centric, R, cudaMemcpyHostToDevice, h->stream));
  CLB_CUDA(h, cudaMemcpyAsync(h->eps_sigma.p, es.data(), sizeof(float) * R, cudaMemcpyHostToDevice, h->stream));
  CLB_CUDA(h, cudaMemcpyAsync(h->refl_index.p, ridx.data(), sizeof(uint32_t) * R, cudaMemcpyHostToDevice, h->stream));
  if (dw) {
    CLB_CUDA(h, h->dw_parent.alloc(sizeof(int32_t) * R)); CLB_CUDA(h, h->asu_id.alloc(sizeof(int32_t) * R)); CLB_CUDA(h, h->r_const.alloc(sizeof(float) * c.n_asu));
    CLB_CUDA(h, cudaMemcpyAsync(h->dw_parent.p, dw_parent, sizeof(int32_t) * R, cudaMemcpyHostToDevice, h->stream));
    CLB_CUDA(h, cudaMemcpyAsync(h->asu_id.p, asu_id, sizeof(int32_t) * R, cudaMemcpyHostToDevice, h->stream));
    CLB_CUDA(h, cudaMemcpyAsync(h->r_const.p, r, sizeof(float) * c.n_asu, cudaMemcpyHostToDevice, h->stream));
    if (c.optimize_dw_r) {
      std::vector<float> logit((size_t)c.n_asu);
      for (int i = 0; i < c.n_asu; ++i) logit[i] = std::log(r[i]) - std::log1p(-r[i]);     // tfb.Sigmoid inverse, wilson.py:105-110
      CLB_CUDA(h, cudaMemcpyAsync(h->theta.as<float>() + h->goff[CLB_GROUP_DW_R], logit.data(), sizeof(float) * c.n_asu, cudaMemcpyHostToDevice, h->stream));
    }
  }
  if (init_scale >= 0.f) {   // manager.py:432-436: loc = prior.mean(), scale = prior.stddev() * init_scale
    std::vector<float> vl((size_t)R), vs((size_t)R);
    for (int64_t i = 0; i < R; ++i) {
      const double s = std::sqrt((double)es[i]);
      double mean, sd;
      if (centric[i]) { mean = s * std::sqrt(2.0 / M_PI); sd = s * std::sqrt(1.0 - 2.0 / M_PI); }
      else { mean = s * std::sqrt(M_PI) / 2.0; sd = s * std::sqrt(1.0 - M_PI / 4.0); }
      const float loc = (float)mean, scale = (float)(sd * init_scale);
      vl[i] = std::log(loc);
      vs[i] = std::log(scale - c.epsilon);
    }
    CLB_CUDA(h, cudaMemcpyAsync(h->theta.as<float>() + h->goff[CLB_GROUP_SF_LOC], vl.data(), sizeof(float) * R, cudaMemcpyHostToDevice, h->stream));
    CLB_CUDA(h, cudaMemcpyAsync(h->theta.as<float>() + h->goff[CLB_GROUP_SF_SCALE], vs.data(), sizeof(float) * R, cudaMemcpyHostToDevice, h->stream));
  }
  CLB_CUDA(h, cudaStreamSynchronize(h->stream));
  h->have_prior = true;
  return CLB_OK;
}

int64_t clb_group_size(const clb_handle* h, int32_t group) {
  if (!h || group < 0 || group >= CLB_N_GROUPS) return -1;
  return h->gsize[group];
}

static int copy_group(clb_handle* h, DevBuf& buf, int32_t group, float* host, int64_t n, bool to_host) {
  if (!h || group < 0 || group >= CLB_N_GROUPS || (!host && n > 0)) return fail(h, CLB_ERR_INVALID, "bad group/pointer");
  if (n != h->gsize[group]) return fail(h, CLB_ERR_INVALID, "group %d has %lld values, caller passed %lld", group, (long long)h->gsize[group], (long long)n);
  if (n == 0) return CLB_OK;
  CLB_CUDA(h, cudaSetDevice(h->cfg.device));
  float* dev = buf.as<float>() + h->goff[group];
  if (to_host) CLB_CUDA(h, cudaMemcpyAsync(host, dev, sizeof(float) * n, cudaMemcpyDeviceToHost, h->stream));
  else CLB_CUDA(h, cudaMemcpyAsync(dev, host, sizeof(float) * n, cudaMemcpyHostToDevice, h->stream));
  CLB_CUDA(h, cudaStreamSynchronize(h->stream));
  return CLB_OK;
}

int clb_get_params(clb_handle* h, int32_t group, float* out, int64_t n) { return copy_group(h, h->theta, group, out, n, true); }
int clb_set_params(clb_handle* h, int32_t group, const float* in, int64_t n) { return copy_group(h, h->theta, group, const_cast<float*>(in), n, false); }
int clb_get_grads(clb_handle* h, int32_t group, float* out, int64_t n) { return copy_group(h, h->grad, group, out, n, true); }

int clb_get_adam_state(clb_handle* h, int32_t group, float* m, float* v, int64_t n, int64_t* t) {
  if (!h) return CLB_ERR_INVALID;
  if (m) { int rc = copy_group(h, h->m, group, m, n, true); if (rc) return rc; }
  if (v) { int rc = copy_group(h, h->v, group, v, n, true); if (rc) return rc; }
  if (t) *t = h->adam_t;
  return CLB_OK;
}

int clb_set_trainable(clb_handle* h, int32_t group, int32_t trainable) {
  if (!h || group < 0 || group >= CLB_N_GROUPS) return fail(h, CLB_ERR_INVALID, "bad group");
  h->gtrain[group] = trainable ? 1 : 0;
  refresh_trainable(h);
  return CLB_OK;
}

int clb_enable_ipred(clb_handle* h, int32_t enable) {
  if (!h) return CLB_ERR_INVALID;
  h->want_ipred = enable != 0;
  return CLB_OK;
}

int clb_get_samples(clb_handle* h, float* z_f, int64_t n) {
  if (!h || !z_f || n != h->R * h->S) return fail(h, CLB_ERR_INVALID, "clb_get_samples: expected %lld values", h ? (long long)(h->R * h->S) : 0LL);
  CLB_CUDA(h, cudaSetDevice(h->cfg.device));
  CLB_CUDA(h, cudaMemcpyAsync(z_f, h->z.p, sizeof(float) * n, cudaMemcpyDeviceToHost, h->stream));
  CLB_CUDA(h, cudaStreamSynchronize(h->stream));
  return CLB_OK;
}

int clb_get_ipred(clb_handle* h, float* ipred, int64_t n) {
  if (!h || !ipred || !h->ipred.p || n != h->n_rows_total * h->S) return fail(h, CLB_ERR_INVALID, "clb_get_ipred: call clb_enable_ipred(1) and a step first; expected %lld values", h ? (long long)(h->n_rows_total * h->S) : 0LL);
  CLB_CUDA(h, cudaSetDevice(h->cfg.device));
  CLB_CUDA(h, cudaMemcpyAsync(ipred, h->ipred.p, sizeof(float) * n, cudaMemcpyDeviceToHost, h->stream));
  CLB_CUDA(h, cudaStreamSynchronize(h->stream));
  return CLB_OK;
}

int clb_reset_timers(clb_handle* h, int32_t enable) {
  if (!h) return CLB_ERR_INVALID;
  h->timing = enable != 0; h->ev_used = 0; h->obs_ms = 0.0; h->obs_launches = 0; h->total_launches = 0;
  return CLB_OK;
}

int clb_kernel_time_ms(clb_handle* h, double* ms, int64_t* launches, int64_t* total) {
  if (!h) return CLB_ERR_INVALID;
  CLB_CUDA(h, cudaSetDevice(h->cfg.device));
  CLB_CUDA(h, cudaStreamSynchronize(h->stream));
  for (size_t i = 0; i + 1 < h->ev_used; i += 2) {
    float t = 0.f;
    CLB_CUDA(h, cudaEventElapsedTime(&t, h->ev[i], h->ev[i + 1]));
    h->obs_ms += t;
  }
  h->ev_used = 0;
  if (ms) *ms = h->obs_ms;
  if (launches) *launches = h->obs_launches;
  if (total) *total = h->total_launches;
  return CLB_OK;
}

// ---------------------------------------------------------------------------------------
// the step
// ---------------------------------------------------------------------------------------
static int ensure_metrics(clb_handle* h, int n) {
  if (n <= h->metrics_cap) return CLB_OK;
  CLB_CUDA(h, h->metrics.alloc(sizeof(double) * 4 * n));
  h->metrics_cap = n;
  return CLB_OK;
}

static int step_begin_impl(clb_handle* h, const float* inj_u, const float* inj_eps) {
  const clb_config& c = h->cfg;
  if (!h->have_obs || !h->have_prior) return fail(h, CLB_ERR_STATE, "clb_step before clb_set_observations / clb_set_prior");
  CLB_CUDA(h, cudaSetDevice(c.device));
  cudaStream_t st = h->stream;
  if (h->pending_swap) {        // the prefetched rows become this step's input once their copy has landed
    CLB_CUDA(h, cudaStreamWaitEvent(st, h->ev_copy_done, 0));
    std::swap(h->rows.p, h->rows_alt.p); std::swap(h->rows.bytes, h->rows_alt.bytes);
    point_rows(h, h->rows.as<char>());
    h->pending_swap = false; h->alt_used = true; h->cur_rows ^= 1;
  }
  const int64_t R = h->R; const int S = h->S;
  float* theta = h->theta.as<float>();
  float* grad = h->grad.as<float>();
  const double kl_div = c.use_kl_weight ? (double)S * (double)c.n_refl_total : (double)S;
  const double kl_coef = c.use_kl_weight ? (double)c.kl_weight : 1.0;
  const double ll_div = c.use_kl_weight ? (double)S * (double)h->n_rows_total : (double)S;
  const float cq = (float)(kl_coef / kl_div), cl = (float)(1.0 / ll_div);

  const float* d_inj_u = nullptr; const float* d_inj_eps = nullptr;
  if (inj_u) {
    CLB_CUDA(h, h->inj_u.alloc(sizeof(float) * R * S));
    CLB_CUDA(h, cudaMemcpyAsync(h->inj_u.p, inj_u, sizeof(float) * R * S, cudaMemcpyHostToDevice, st));
    d_inj_u = h->inj_u.as<float>();
  }
  if (inj_eps) {
    CLB_CUDA(h, h->inj_eps.alloc(sizeof(float) * h->n_rows_total * S));
    CLB_CUDA(h, cudaMemcpyAsync(h->inj_eps.p, inj_eps, sizeof(float) * h->n_rows_total * S, cudaMemcpyHostToDevice, st));
    d_inj_eps = h->inj_eps.as<float>();
  }
  if (h->want_ipred) {
    CLB_CUDA(h, h->ipred.alloc(sizeof(float) * h->n_rows_total * S));
    CLB_CUDA(h, cudaMemsetAsync(h->ipred.p, 0, sizeof(float) * h->n_rows_total * S, st));
  }
  CLB_CUDA(h, cudaMemsetAsync(h->acc.p, 0, sizeof(double) * ACC_COUNT, st));
  CLB_CUDA(h, cudaMemsetAsync(h->var_sums.p, 0, sizeof(double) * 2 * kMaxVars, st));
  // gradients of the replicated groups are accumulated with atomics / overwritten by the reduction
  if (h->P > 2 * R) CLB_CUDA(h, cudaMemsetAsync(grad + 2 * R, 0, sizeof(float) * (h->P - 2 * R), st));
  const bool train_mlp = h->gtrain[CLB_GROUP_MLP] != 0 && !h->eval_mode;
  if (train_mlp) CLB_CUDA(h, cudaMemsetAsync(h->partials.p, 0, h->partial_bytes, st));

  const bool dw = c.prior == CLB_PRIOR_DOUBLE_WILSON;
  {
    ReflArgs a{};
    a.v_loc = theta + h->goff[CLB_GROUP_SF_LOC]; a.v_scale = theta + h->goff[CLB_GROUP_SF_SCALE];
    a.centric = h->centric.as<uint8_t>(); a.eps_sigma = h->eps_sigma.as<float>();
    a.dw_parent = dw ? h->dw_parent.as<int32_t>() : nullptr;
    a.refl_index = h->refl_index.as<uint32_t>(); a.inj_u = d_inj_u;
    a.z = h->z.as<float>(); a.gz = h->gz.as<float>(); a.acc = h->acc.as<double>();
    a.R = R; a.S = S; a.eps = c.epsilon; a.cq = cq; a.seed = c.seed; a.step = h->step_counter;
    if (!h->eval_mode) { CLB_CUDA(h, h->bwd_coef.alloc(sizeof(float4) * (size_t)R * S)); a.bwd_coef = h->bwd_coef.as<float4>(); }
    const int64_t nthr = ((R + 3) / 4) * S;          // four reflections per thread
    const unsigned nblk = (unsigned)((nthr + 255) / 256);
    if (h->det) { CLB_CUDA(h, h->kl_part.alloc(sizeof(double) * nblk)); a.kl_part = h->kl_part.as<double>(); h->kl_blocks = (int)nblk; }
    k_refl_sample<<<nblk, 256, 0, st>>>(a, (R % 4 == 0) ? 1 : 0);
    CLB_LAUNCHED(h);
    if (h->det) { k_sum_partials<<<1, 32, 0, st>>>(h->kl_part.as<double>(), (int)nblk, (int)nblk, 1, h->acc.as<double>() + ACC_LOGQ_MINUS_LOGP); CLB_LAUNCHED(h); }
  }
  if (dw) {
    DwArgs a{};
    a.z = h->z.as<float>(); a.gz = h->gz.as<float>();
    a.centric = h->centric.as<uint8_t>(); a.eps_sigma = h->eps_sigma.as<float>();
    a.dw_parent = h->dw_parent.as<int32_t>(); a.asu_id = h->asu_id.as<int32_t>();
    a.r_const = h->r_const.as<float>();
    a.r_logit = c.optimize_dw_r ? theta + h->goff[CLB_GROUP_DW_R] : nullptr;
    a.g_r_logit = c.optimize_dw_r ? grad + h->goff[CLB_GROUP_DW_R] : nullptr;
    a.acc = h->acc.as<double>(); a.R = R; a.S = S; a.cq = cq;
    const int64_t nthr = R * S;
    k_dw_prior<<<(unsigned)((nthr + 255) / 256), 256, 0, st>>>(a);
    CLB_LAUNCHED(h);
  }
  {
    ObsArgs a{};
    a.refl = h->d_refl; a.image = h->d_image; a.spot = h->d_spot; a.oidx = h->d_oidx;
    a.meta = h->d_meta; a.iobs = h->d_iobs; a.sig = h->d_sig;
    a.n_rows = h->n_rows; a.n_rows_total = h->n_rows_total; a.d = c.n_meta;
    a.theta_mlp = theta + h->goff[CLB_GROUP_MLP];
    a.theta_img = c.image_scales ? theta + h->goff[CLB_GROUP_IMAGE_SCALES] : nullptr;
    a.wpack = (h->use_tc || h->WP == 64) ? h->wpack.as<float>() : nullptr;
    if (h->WP == 64 && c.mlp_layers > 0) {
      k_pack_weights<<<(c.mlp_layers * 4096 + 255) / 256, 256, 0, st>>>(a.theta_mlp, h->lay, h->wpack.as<float>(), 64);
      CLB_LAUNCHED(h);
    }
    a.wimg = (h->use_tc2 || h->use_tc16) ? h->wimg.as<float>() : nullptr;
    if (h->use_tc16 && c.mlp_layers > 0) {
      k_pack_images16<<<(c.mlp_layers * 512 + 255) / 256, 256, 0, st>>>(a.theta_mlp, h->lay, h->wimg.as<float>(), h->bias_feat15 ? 1 : 0);
      CLB_LAUNCHED(h);
    }
    a.n_img_layers = c.image_layers; a.il_width = c.mlp_width; a.il_n_images = c.n_images;
    a.theta_il = c.image_layers > 0 ? theta + h->goff[CLB_GROUP_IMAGE_LAYERS] : nullptr;
    a.g_il = (c.image_layers > 0 && h->gtrain[CLB_GROUP_IMAGE_LAYERS] && !h->eval_mode) ? grad + h->goff[CLB_GROUP_IMAGE_LAYERS] : nullptr;
    if (h->use_tc) {
      if (c.mlp_layers > 0) {
        if (h->use_tc2) k_pack_images<<<(c.mlp_layers * 2048 + 255) / 256, 256, 0, st>>>(a.theta_mlp, h->lay, h->wimg.as<float>());
        else k_pack_weights<<<(c.mlp_layers * 1024 + 255) / 256, 256, 0, st>>>(a.theta_mlp, h->lay, h->wpack.as<float>(), 32);
        CLB_LAUNCHED(h);
      }
    }
    a.lay = h->lay;
    a.z = h->z.as<float>(); a.gz = h->gz.as<float>(); a.R = R; a.S = S;
    a.inj_eps = d_inj_eps;
    a.g_img = (c.image_scales && h->gtrain[CLB_GROUP_IMAGE_SCALES] && !h->eval_mode) ? grad + h->goff[CLB_GROUP_IMAGE_SCALES] : nullptr;
    a.partials = h->partials.as<double>(); a.scratch = h->scratch_ptr;
    a.partials32 = h->partials.as<float>();
    a.ipred_out = h->want_ipred ? h->ipred.as<float>() : nullptr;
    a.scale_mean_out = h->want_scale_moments ? h->scale_mom.as<float>() : nullptr;
    a.scale_std_out = h->want_scale_moments ? h->scale_mom.as<float>() + h->n_rows_total : nullptr;
    a.acc = h->acc.as<double>();
    a.lik.dof = c.dof; a.lik.half_dofp1 = 0.5f * (c.dof + 1.0f);
    a.lik.lnorm = (c.likelihood == CLB_LIK_STUDENTT)
                      ? (float)(lgamma_d(0.5 * (c.dof + 1.0)) - lgamma_d(0.5 * c.dof) - 0.5 * std::log((double)c.dof * M_PI)) : 0.f;
    a.cl = cl; a.bijector = c.scale_bijector; a.shift = c.scale_shift; a.eps = c.epsilon;
    a.theta_lik = c.refine_uncertainties ? theta + h->goff[CLB_GROUP_LIKELIHOOD] : nullptr;
    a.g_lik = (c.refine_uncertainties && h->gtrain[CLB_GROUP_LIKELIHOOD] && !h->eval_mode) ? grad + h->goff[CLB_GROUP_LIKELIHOOD] : nullptr;
    if (h->n_empty > 0) {
      const int blocks = (int)std::min<int64_t>((h->n_empty + 255) / 256, 4 * h->n_sms);
      const float* ei = h->empty_slots.as<float>();
      if (c.likelihood == CLB_LIK_NORMAL)
        k_ev11_empty<0><<<blocks, 256, 0, st>>>(ei, ei + h->n_empty, h->n_empty, a.theta_lik, a.g_lik, h->acc.as<double>(), a.lik, cl, (float)S);
      else
        k_ev11_empty<1><<<blocks, 256, 0, st>>>(ei, ei + h->n_empty, h->n_empty, a.theta_lik, a.g_lik, h->acc.as<double>(), a.lik, cl, (float)S);
      CLB_LAUNCHED(h);
    }
    a.seed = c.seed; a.step = h->step_counter; a.laue = c.laue; a.train_mlp = train_mlp ? 1 : 0;
    a.discard_scratch = h->discard_scratch ? 1 : 0;
    a.bias_feat15 = (h->use_tc16 && h->bias_feat15) ? 1 : 0;
    a.n_partials = h->n_partials;
    a.det = h->det ? 1 : 0;
    if (h->det) {
      a.dzf_rows = h->dzf_rows.as<float>(); a.ll_part = h->ll_part.as<double>();
      CLB_CUDA(h, cudaMemsetAsync(h->ll_part.p, 0, sizeof(double) * h->grid_obs, st));
    }
    if (h->timing) {
      while (h->ev.size() < h->ev_used + 2) { cudaEvent_t e; CLB_CUDA(h, cudaEventCreate(&e)); h->ev.push_back(e); }
      CLB_CUDA(h, cudaEventRecord(h->ev[h->ev_used], st));
    }
    CLB_CUDA(h, dispatch_obs(h, a)); h->obs_launches++; CLB_LAUNCHED(h);
    if (h->det) {
      k_sum_partials<<<1, 32, 0, st>>>(h->ll_part.as<double>(), h->grid_obs, h->grid_obs, 1, h->acc.as<double>() + ACC_LL); CLB_LAUNCHED(h);
      if (!h->eval_mode) {
        k_gz_reduce<<<(unsigned)((R * S + 255) / 256), 256, 0, st>>>(h->dzf_rows.as<float>(), h->refl_ptr.as<int32_t>(), h->refl_rows.as<int32_t>(),
                                                                    h->gz.as<float>(), R, S, h->n_rows);
        CLB_LAUNCHED(h);
      }
    }
    if (h->copy_stream) CLB_CUDA(h, cudaEventRecord(h->ev_rows_free[h->cur_rows], st));   // the row buffer of this step may be refilled after this point
    if (h->timing) { CLB_CUDA(h, cudaEventRecord(h->ev[h->ev_used + 1], st)); h->ev_used += 2; }
  }
  if (train_mlp) {
    const int np = h->lay.n_params;
    if (h->use_tc16) k_reduce_partials16<<<(np + 31) / 32, 256, 0, st>>>(h->partials.as<float>(), h->n_partials, h->lay, grad + h->goff[CLB_GROUP_MLP], h->det ? 1 : 0);
    else if (h->use_tc2) k_reduce_partials32<<<(np + 31) / 32, 256, 0, st>>>(h->partials.as<float>(), h->n_partials, h->lay, grad + h->goff[CLB_GROUP_MLP], h->det ? 1 : 0,
                                                                               (CLB_BIAS_COL && !h->use_pp && !h->use_tc3) ? 1 : 0);
    else k_reduce_partials<<<(np + 255) / 256, 256, 0, st>>>(h->partials.as<double>(), h->grid_obs * h->KS, h->lay, h->WP, grad + h->goff[CLB_GROUP_MLP]);
    CLB_LAUNCHED(h);
  }
  if (!h->eval_mode) {
    ReflBwdArgs a{};
    a.v_loc = theta + h->goff[CLB_GROUP_SF_LOC]; a.v_scale = theta + h->goff[CLB_GROUP_SF_SCALE];
    a.bwd_coef = h->bwd_coef.as<float4>(); a.gz = h->gz.as<float>();
    a.g_loc = grad + h->goff[CLB_GROUP_SF_LOC]; a.g_scale = grad + h->goff[CLB_GROUP_SF_SCALE];
    a.R = R; a.S = S;
    // fused tail of the per-reflection chain: sums of squares always; Adam on the surrogate slice whenever no norm-based
    // clipping is configured (then an element's update needs nothing but its own gradient)
    a.var_sums = h->var_sums.as<double>();
    h->adam_fused = h->gtrain[CLB_GROUP_SF_LOC] && h->gtrain[CLB_GROUP_SF_SCALE] && !(c.clipnorm > 0.f) && !(c.global_clipnorm > 0.f)
                    && !h->no_fused_adam;
    if (h->adam_fused) {
      a.theta_loc = theta + h->goff[CLB_GROUP_SF_LOC]; a.theta_scale = theta + h->goff[CLB_GROUP_SF_SCALE];
      a.m_loc = h->m.as<float>() + h->goff[CLB_GROUP_SF_LOC]; a.m_scale = h->m.as<float>() + h->goff[CLB_GROUP_SF_SCALE];
      a.v2_loc = h->v.as<float>() + h->goff[CLB_GROUP_SF_LOC]; a.v2_scale = h->v.as<float>() + h->goff[CLB_GROUP_SF_SCALE];
    }
    a.alpha = adam_alpha_host(c, h->adam_t + 1); a.beta1 = c.beta_1; a.beta2 = c.beta_2; a.adam_eps = c.adam_epsilon; a.clipvalue = c.clipvalue;
    a.stop_step = h->stop_step.as<int>(); a.step_index = (int)h->step_counter;
    const unsigned nblk = (unsigned)(((R + 3) / 4 + 255) / 256);
    if (h->det) { CLB_CUDA(h, h->ss_part.alloc(sizeof(double) * 4 * nblk)); a.ss_part = h->ss_part.as<double>(); }
    k_refl_backward<<<nblk, 256, 0, st>>>(a, (R % 4 == 0) ? 1 : 0);
    CLB_LAUNCHED(h);
    if (h->det) { k_sum_partials<<<1, 32, 0, st>>>(h->ss_part.as<double>(), (int)nblk, (int)nblk, 4, h->var_sums.as<double>()); CLB_LAUNCHED(h); }
  }
  h->in_step = true;
  return CLB_OK;
}

// after the all-reduce of the replicated gradients: per-variable norms + scalar packing
static int step_norms_impl(clb_handle* h) {
  if (!h->in_step) return fail(h, CLB_ERR_STATE, "clb_step_norms without clb_step_begin");
  cudaStream_t st = h->stream;
  refresh_trainable(h);
  int64_t maxsz = 1;
  for (int v = 2; v < h->vt.n_vars; ++v) maxsz = std::max(maxsz, h->vt.size[v]);     // variables 0 / 1 (the surrogate) are summed in k_refl_backward
  const int chunks = (int)std::min<int64_t>((maxsz + 256 * 8 - 1) / (256 * 8), 4 * h->n_sms);
  const dim3 grid(h->det ? 1 : std::max(chunks, 1), h->vt.n_vars);      // deterministic: one block per variable, no cross-block atomics
  if (h->comm == nullptr) {
    k_var_sumsq<<<grid, 256, 0, st>>>(h->grad.as<float>(), h->vt, h->var_sums.as<double>(), 0, 1);      // the surrogate's sums come from k_refl_backward
    CLB_LAUNCHED(h);
    k_pack_scalars<<<1, 128, 0, st>>>(h->acc.as<double>(), h->var_sums.as<double>(), h->vt, h->red.as<double>(), h->cfg.rank,
                                    (double)h->S * h->ll_const, 0);
    CLB_LAUNCHED(h);
    return CLB_OK;
  }
  // In-library exchange (clb_comm_init): everything that is rank-local -- sum(log q - log p), sum(ll), the sums of squares of
  // the surrogate gradients -- is known BEFORE the replicated gradients are reduced, so the float32 gradient buffer and the
  // float64 scalar buffer travel in ONE grouped NCCL all-reduce on the step's stream; the replicated variables' norms are
  // then taken from the reduced gradient, identically on every rank.
  // (the only rank-local variables are the surrogate's, whose sums k_refl_backward has already accumulated)
  k_pack_scalars<<<1, 128, 0, st>>>(h->acc.as<double>(), h->var_sums.as<double>(), h->vt, h->red.as<double>(), h->cfg.rank,
                                  (double)h->S * h->ll_const, 1);
  CLB_LAUNCHED(h);
  {
    const int64_t nf = h->P - 2 * h->R, nd = 2 + 2 * h->vt.n_vars;
    float* gf = h->grad.as<float>() + 2 * h->R;
    int rc = g_nccl.GroupStart();
    if (rc == 0 && nf > 0) rc = g_nccl.AllReduce(gf, gf, (size_t)nf, kNcclFloat32, kNcclSum, h->comm, st);
    if (rc == 0) rc = g_nccl.AllReduce(h->red.p, h->red.p, (size_t)nd, kNcclFloat64, kNcclSum, h->comm, st);
    const int rc2 = g_nccl.GroupEnd();
    if (rc == 0) rc = rc2;
    if (rc != 0) return fail(h, CLB_ERR_CUDA, "NCCL all-reduce failed: %s", g_nccl.GetErrorString(rc));
    h->exchanges++;
  }
  k_var_sumsq<<<grid, 256, 0, st>>>(h->grad.as<float>(), h->vt, h->red.as<double>() + 2, 2, 0);
  CLB_LAUNCHED(h);
  return CLB_OK;
}

static int step_end_impl(clb_handle* h, double* d_metrics) {
  const clb_config& c = h->cfg;
  if (!h->in_step) return fail(h, CLB_ERR_STATE, "clb_step_end without clb_step_begin");
  cudaStream_t st = h->stream;
  const int S = h->S;
  FinalizeArgs f{};
  f.acc = h->acc.as<double>(); f.red = h->red.as<double>(); f.vt = h->vt;
  f.metrics = d_metrics; f.var_scale = h->var_scale.as<float>(); f.adam_alpha = h->adam_alpha.as<float>();
  f.stop_step = h->stop_step.as<int>(); f.step = (int)h->step_counter;
  f.kl_div = c.use_kl_weight ? (double)S * (double)c.n_refl_total : (double)S;
  f.kl_coef = c.use_kl_weight ? (double)c.kl_weight : 1.0;
  f.ll_div = c.use_kl_weight ? (double)S * (double)h->n_rows_total : (double)S;
  f.clipnorm = c.clipnorm; f.global_clipnorm = c.global_clipnorm;
  f.lr = c.learning_rate; f.beta1 = c.beta_1; f.beta2 = c.beta_2; f.t = h->adam_t + 1;
  f.alpha = adam_alpha_host(c, h->adam_t + 1);
  k_finalize<<<1, 32, 0, st>>>(f);
  CLB_LAUNCHED(h);
  int64_t maxsz = 1;
  for (int v = h->adam_fused ? 2 : 0; v < h->vt.n_vars; ++v) if (h->vt.trainable[v]) maxsz = std::max(maxsz, h->vt.size[v]);
  const int chunks = (int)std::min<int64_t>((maxsz + 256 * 16 - 1) / (256 * 16), 8 * h->n_sms);
  k_adam<<<dim3(std::max(chunks, 1), h->vt.n_vars), 256, 0, st>>>(h->theta.as<float>(), h->m.as<float>(), h->v.as<float>(), h->grad.as<float>(),
                                                                   h->vt, h->var_scale.as<float>(), h->adam_alpha.as<float>(),
                                                                   c.clipvalue, c.beta_1, c.beta_2, c.adam_epsilon, h->stop_step.as<int>(), (int)h->step_counter,
                                                                   h->adam_fused ? 1 : 0);
  CLB_LAUNCHED(h);
  h->adam_t += 1;
  h->step_counter += 1;
  h->in_step = false;
  return CLB_OK;
}

// Forward only: the "NLL" / "F KLDiv" / "loss" of the current parameters on this handle's observations with fresh
// draws and no update (keras test_on_batch at variational.py:257-260).  grad_norm is reported as 0.
int clb_eval(clb_handle* h, const float* inj_u_f, const float* inj_eps_s, clb_metrics* out) {
  if (!h) return CLB_ERR_INVALID;
  if (h->cfg.world_size > 1 && h->comm == nullptr) return fail(h, CLB_ERR_STATE, "clb_eval with world_size > 1 needs clb_comm_init");
  int rc = ensure_metrics(h, 1); if (rc) return rc;
  h->eval_mode = true;
  rc = step_begin_impl(h, inj_u_f, inj_eps_s);
  h->eval_mode = false;
  if (rc) return rc;
  cudaStream_t st = h->stream;
  CLB_CUDA(h, cudaMemsetAsync(h->var_sums.p, 0, sizeof(double) * 2 * kMaxVars, st));
  k_pack_scalars<<<1, 128, 0, st>>>(h->acc.as<double>(), h->var_sums.as<double>(), h->vt, h->red.as<double>(), h->cfg.rank,
                                    (double)h->S * h->ll_const, 0);
  CLB_LAUNCHED(h);
  if (h->comm != nullptr) {      // the scalar ELBO terms of all ranks
    const int rc = g_nccl.AllReduce(h->red.p, h->red.p, (size_t)2, kNcclFloat64, kNcclSum, h->comm, st);
    if (rc != 0) return fail(h, CLB_ERR_CUDA, "NCCL all-reduce failed: %s", g_nccl.GetErrorString(rc));
  }
  const clb_config& c = h->cfg;
  FinalizeArgs f{};
  f.acc = h->acc.as<double>(); f.red = h->red.as<double>(); f.vt = h->vt; f.vt.n_vars = 0;
  f.metrics = h->metrics.as<double>(); f.var_scale = h->var_scale.as<float>(); f.adam_alpha = h->adam_alpha.as<float>() + 1;
  f.stop_step = h->stop_step.as<int>() + 1; f.step = 0;      // scratch slots: the training early-stop state is untouched
  f.kl_div = c.use_kl_weight ? (double)h->S * (double)c.n_refl_total : (double)h->S;
  f.kl_coef = c.use_kl_weight ? (double)c.kl_weight : 1.0;
  f.ll_div = c.use_kl_weight ? (double)h->S * (double)h->n_rows_total : (double)h->S;
  f.lr = c.learning_rate; f.beta1 = c.beta_1; f.beta2 = c.beta_2; f.t = 1; f.alpha = 0.f;
  k_finalize<<<1, 32, 0, st>>>(f);
  CLB_LAUNCHED(h);
  h->step_counter += 1;
  h->in_step = false;
  if (out) {
    double m[4];
    CLB_CUDA(h, cudaMemcpyAsync(m, h->metrics.p, sizeof m, cudaMemcpyDeviceToHost, st));
    CLB_CUDA(h, cudaStreamSynchronize(st));
    out->loss = m[0]; out->nll = m[1]; out->kl = m[2]; out->grad_norm = m[3];
  }
  return CLB_OK;
}

// Replaces: scale_dist.mean() / scale_dist.stddev() of VariationalMergingModel.scale_mean_stddev and
// prediction_mean_stddev (variational.py:47-121): per-observation moments of the scale distribution in the
// caller's row order (rows of other ranks / unused slots are 0).
int clb_get_scale_moments(clb_handle* h, float* mean, float* stddev, int64_t n) {
  if (!h || !mean || !stddev || n != h->n_rows_total) return fail(h, CLB_ERR_INVALID, "clb_get_scale_moments: expected %lld values", h ? (long long)h->n_rows_total : 0LL);
  CLB_CUDA(h, cudaSetDevice(h->cfg.device));
  CLB_CUDA(h, h->scale_mom.alloc(sizeof(float) * 2 * n));
  CLB_CUDA(h, cudaMemsetAsync(h->scale_mom.p, 0, sizeof(float) * 2 * n, h->stream));
  h->eval_mode = true; h->want_scale_moments = true;
  const bool ip = h->want_ipred; h->want_ipred = false;
  int rc = step_begin_impl(h, nullptr, nullptr);
  h->eval_mode = false; h->want_scale_moments = false; h->want_ipred = ip; h->in_step = false;
  if (rc) return rc;
  CLB_CUDA(h, cudaMemcpyAsync(mean, h->scale_mom.p, sizeof(float) * n, cudaMemcpyDeviceToHost, h->stream));
  CLB_CUDA(h, cudaMemcpyAsync(stddev, h->scale_mom.as<float>() + n, sizeof(float) * n, cudaMemcpyDeviceToHost, h->stream));
  CLB_CUDA(h, cudaStreamSynchronize(h->stream));
  return CLB_OK;
}

// Replaces: the numeric part of DataManager.get_results (io/manager.py:188-197, :209): merged F, SigF, I, SigI and
// the redundancy N of every surrogate entry of this handle.  Any output may be NULL.
int clb_get_results(clb_handle* h, float* F, float* SigF, float* I, float* SigI, float* N, int64_t n) {
  if (!h || n != h->R) return fail(h, CLB_ERR_INVALID, "clb_get_results: expected %lld values", h ? (long long)h->R : 0LL);
  if (!h->have_prior) return fail(h, CLB_ERR_STATE, "clb_get_results before clb_set_prior");
  CLB_CUDA(h, cudaSetDevice(h->cfg.device));
  cudaStream_t st = h->stream;
  CLB_CUDA(h, h->results.alloc(sizeof(float) * 5 * n));
  float* d = h->results.as<float>();
  float* theta = h->theta.as<float>();
  k_results<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(theta + h->goff[CLB_GROUP_SF_LOC], theta + h->goff[CLB_GROUP_SF_SCALE],
                                                          h->centric.as<uint8_t>(), n, h->cfg.epsilon, d, d + n, d + 2 * n, d + 3 * n);
  CLB_LAUNCHED(h);
  CLB_CUDA(h, cudaMemsetAsync(d + 4 * n, 0, sizeof(float) * n, st));
  if (h->have_obs) {
    k_count_obs<<<(unsigned)((h->n_rows + 255) / 256), 256, 0, st>>>(h->d_refl, h->n_rows, d + 4 * n);
    CLB_LAUNCHED(h);
  }
  float* outs[5] = {F, SigF, I, SigI, N};
  for (int q = 0; q < 5; ++q)
    if (outs[q]) CLB_CUDA(h, cudaMemcpyAsync(outs[q], d + (size_t)q * n, sizeof(float) * n, cudaMemcpyDeviceToHost, st));
  CLB_CUDA(h, cudaStreamSynchronize(st));
  return CLB_OK;
}

// ---- in-library exchange step ----
int clb_comm_unique_id(uint8_t* id128) {
  if (!id128) return fail(nullptr, CLB_ERR_INVALID, "clb_comm_unique_id: null argument");
  if (!g_nccl.load()) return fail(nullptr, CLB_ERR_STATE, "%s", g_nccl.err.c_str());
  NcclUniqueId id;
  const int rc = g_nccl.GetUniqueId(&id);
  if (rc != 0) return fail(nullptr, CLB_ERR_CUDA, "ncclGetUniqueId failed: %s", g_nccl.GetErrorString(rc));
  memcpy(id128, id.internal, sizeof id.internal);
  return CLB_OK;
}

int clb_comm_init(clb_handle* h, const uint8_t* id128) {
  if (!h || !id128) return fail(h, CLB_ERR_INVALID, "clb_comm_init: null argument");
  if (h->cfg.world_size <= 1) return fail(h, CLB_ERR_INVALID, "clb_comm_init: the handle was created with world_size %d", h->cfg.world_size);
  if (h->comm) return fail(h, CLB_ERR_STATE, "clb_comm_init: communicator already initialised");
  if (!g_nccl.load()) return fail(h, CLB_ERR_STATE, "%s", g_nccl.err.c_str());
  CLB_CUDA(h, cudaSetDevice(h->cfg.device));
  NcclUniqueId id;
  memcpy(id.internal, id128, sizeof id.internal);
  const int rc = g_nccl.CommInitRank(&h->comm, h->cfg.world_size, id, h->cfg.rank);
  if (rc != 0) { h->comm = nullptr; return fail(h, CLB_ERR_CUDA, "ncclCommInitRank failed: %s", g_nccl.GetErrorString(rc)); }
  return CLB_OK;
}

int clb_step_begin(clb_handle* h, const float* inj_u_f, const float* inj_eps_s) {
  if (!h) return CLB_ERR_INVALID;
  return step_begin_impl(h, inj_u_f, inj_eps_s);
}

int clb_step_norms(clb_handle* h) {
  if (!h) return CLB_ERR_INVALID;
  return step_norms_impl(h);
}

int clb_step_end(clb_handle* h, clb_metrics* out) {
  if (!h) return CLB_ERR_INVALID;
  int rc = ensure_metrics(h, 1); if (rc) return rc;
  rc = step_end_impl(h, h->metrics.as<double>()); if (rc) return rc;
  if (out) {
    double m[4];
    CLB_CUDA(h, cudaMemcpyAsync(m, h->metrics.p, sizeof m, cudaMemcpyDeviceToHost, h->stream));
    CLB_CUDA(h, cudaStreamSynchronize(h->stream));
    out->loss = m[0]; out->nll = m[1]; out->kl = m[2]; out->grad_norm = m[3];
  }
  return CLB_OK;
}

int clb_reduce_buffers(clb_handle* h, void** gf, int64_t* nf, void** sd, int64_t* nd) {
  if (!h) return CLB_ERR_INVALID;
  if (gf) *gf = h->grad.as<float>() + 2 * h->R;
  if (nf) *nf = h->P - 2 * h->R;
  if (sd) *sd = h->red.p;
  if (nd) *nd = 2 + 2 * h->vt.n_vars;
  return CLB_OK;
}

int clb_step(clb_handle* h, int32_t n_steps, const float* inj_u_f, const float* inj_eps_s,
             clb_metrics* out, int32_t* steps_done) {
  if (!h || n_steps <= 0) return fail(h, CLB_ERR_INVALID, "clb_step: n_steps must be positive");
  if (h->cfg.world_size > 1 && h->comm == nullptr)
    return fail(h, CLB_ERR_STATE, "clb_step with world_size > 1 needs clb_comm_init (or drive clb_step_begin/_norms/_end around your own all-reduces)");
  int rc = ensure_metrics(h, n_steps); if (rc) return rc;
  {
    const int big = 0x7fffffff;   // a new train_model call starts with a clean early-stop state
    CLB_CUDA(h, cudaSetDevice(h->cfg.device));
    CLB_CUDA(h, cudaMemcpyAsync(h->stop_step.p, &big, sizeof(int), cudaMemcpyHostToDevice, h->stream));
  }
  const int first_step = (int)h->step_counter;
  const int64_t t0 = h->adam_t;
  for (int i = 0; i < n_steps; ++i) {
    const float* u = inj_u_f ? inj_u_f + (size_t)i * h->R * h->S : nullptr;
    const float* e = inj_eps_s ? inj_eps_s + (size_t)i * h->n_rows_total * h->S : nullptr;
    rc = step_begin_impl(h, u, e); if (rc) return rc;
    rc = step_norms_impl(h); if (rc) return rc;
    rc = step_end_impl(h, h->metrics.as<double>() + 4 * i); if (rc) return rc;
  }
  std::vector<double> m((size_t)4 * n_steps);
  int stop = 0;
  CLB_CUDA(h, cudaMemcpyAsync(m.data(), h->metrics.p, sizeof(double) * 4 * n_steps, cudaMemcpyDeviceToHost, h->stream));
  CLB_CUDA(h, cudaMemcpyAsync(&stop, h->stop_step.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  CLB_CUDA(h, cudaStreamSynchronize(h->stream));
  int done = n_steps;
  if (stop != 0x7fffffff && stop >= first_step) {      // variational.py:271-274
    done = std::min(n_steps, stop - first_step + 1);
    h->adam_t = t0 + done;
  }
  if (out) for (int i = 0; i < done; ++i) { out[i].loss = m[4 * i]; out[i].nll = m[4 * i + 1]; out[i].kl = m[4 * i + 2]; out[i].grad_norm = m[4 * i + 3]; }
  if (steps_done) *steps_done = done;
  return CLB_OK;
}

}  // extern "C"

#ifdef CLB_PHASE_TIMING
// Debug build only (tools/phase_times.py): read and clear the per-phase cycle sums of k_obs.
extern "C" int clb_debug_phases(unsigned long long* out32) {
  unsigned long long zero[32] = {0};
  if (cudaMemcpyFromSymbol(out32, clb::g_phase, sizeof zero) != cudaSuccess) return -1;
  if (cudaMemcpyToSymbol(clb::g_phase, zero, sizeof zero) != cudaSuccess) return -1;
  return 0;
}
#endif
