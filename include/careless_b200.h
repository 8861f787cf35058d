/*
 * careless_b200.h -- C-ABI of libcareless_b200.so: the B200-native ELBO-gradient + Adam step
 * of careless's VariationalMergingModel.
 *
 * The reference (rs-station/careless v0.5.4) is pure Python over TensorFlow and defines NO
 * FFI of its own; the boundary it exposes for this path is the Python object protocol of
 *   careless/models/merging/variational.py:15      VariationalMergingModel(...)
 *   careless/models/merging/variational.py:226     train_model(data, steps, ...)
 *   careless/models/merging/variational.py:185     train_step_with_gradient_norm(...)
 *   careless/io/manager.py:380                     DataManager.build_model(...)
 * Each entry point below names the reference interface it replaces.  All functions return
 * 0 on success and a negative clb_status on failure; clb_last_error() gives the message.
 * Host arrays are borrowed for the duration of the call and copied; device memory is
 * owned by the handle.  A handle is bound to one CUDA device and one stream and is not
 * thread-safe.  There is no CPU fallback: clb_create() fails without an sm_100 device.
 */
#ifndef CARELESS_B200_H
#define CARELESS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CLB_ABI_VERSION 4

typedef struct clb_handle clb_handle;

typedef enum {
  CLB_OK = 0,
  CLB_ERR_INVALID = -1,     /* bad argument / unsupported configuration */
  CLB_ERR_CUDA = -2,        /* CUDA runtime error (message has the cudaError string) */
  CLB_ERR_NO_DEVICE = -3,   /* no sm_100 device: the library never falls back to the CPU */
  CLB_ERR_STATE = -4        /* call order violated (e.g. step before set_observations) */
} clb_status;

/* likelihood kinds: careless/models/likelihoods/mono.py:16-37, laue.py:68-100 */
enum { CLB_LIK_NORMAL = 0, CLB_LIK_STUDENTT = 1 };
/* prior kinds: careless/models/priors/wilson.py:29-80 (Wilson), :82-175 (DoubleWilson) */
enum { CLB_PRIOR_WILSON = 0, CLB_PRIOR_DOUBLE_WILSON = 1 };
/* scale bijector: careless/io/manager.py:450-463 (--scale-bijector exp|softplus) */
enum { CLB_BIJ_EXP = 0, CLB_BIJ_SOFTPLUS = 1 };
/* row order chosen by the host prep inside clb_set_observations */
enum { CLB_ORDER_AUTO = 0, CLB_ORDER_REFL = 1, CLB_ORDER_SPOT = 2, CLB_ORDER_IMAGE = 3, CLB_ORDER_NONE = 4 };

/* parameter groups for clb_get_params / clb_set_params / clb_get_grads / clb_set_trainable.
 * They are the reference's trainable variables:
 *   SF_LOC, SF_SCALE  surrogate_posteriors.py:104-131 (raw, i.e. inverse-bijected values)
 *   MLP               scaling/nn.py:55-79, keras order [kernel_0, bias_0, ..., kernel_out, bias_out]
 *   IMAGE_SCALES      scaling/image.py:21   (n_images-1 values; image 0 is pinned to 1)
 *   DW_R              priors/wilson.py:105-110 (logits of r, one per ASU; --optimize-double-wilson-r)
 *   IMAGE_LAYERS      scaling/image.py:73-88 (--image-layers): per layer kernel (n_images, W, W) as (out, in), bias (n_images, W)
 *   LIKELIHOOD        likelihoods/mono.py:42-44 (--refine-uncertainties): raw Sdfac, Sdadd, SdB (softplus-transformed, init 1)
 */
enum { CLB_GROUP_SF_LOC = 0, CLB_GROUP_SF_SCALE = 1, CLB_GROUP_MLP = 2, CLB_GROUP_IMAGE_SCALES = 3,
       CLB_GROUP_DW_R = 4, CLB_GROUP_IMAGE_LAYERS = 5, CLB_GROUP_LIKELIHOOD = 6, CLB_N_GROUPS = 7 };

/* Everything DataManager.build_model (careless/io/manager.py:380-507) decides, flattened. */
typedef struct {
  int32_t abi_version;        /* = CLB_ABI_VERSION */
  int32_t device;             /* CUDA device ordinal */
  void*   stream;             /* cudaStream_t to run on; NULL = library-owned stream */

  int64_t n_refl;             /* R: surrogate entries held by this handle (all ASU reflections) */
  int64_t n_refl_total;       /* R over all ranks (= n_refl on one GPU); used by --kl-weight mean */
  int32_t n_meta;             /* d: metadata columns */
  int32_t mlp_width;          /* W   (--mlp-width, args/scaling.py:27-31) */
  int32_t mlp_layers;         /* L   (--mlp-layers, args/scaling.py:21-25) */
  int32_t n_images;           /* max(image_id)+1 when image scales or image layers are on, else 0 */
  int32_t image_scales;       /* HybridImageScaler(MLPScaler, ImageScaler): manager.py:484-487 */
  int32_t mc_samples;         /* S   (--mc-samples, args/common.py:11-15) */
  int32_t likelihood;         /* CLB_LIK_* */
  float   dof;                /* --studentt-likelihood-dof */
  int32_t laue;               /* careless poly: harmonic segment-sum before the likelihood */
  int32_t prior;              /* CLB_PRIOR_* */
  int32_t n_asu;              /* DoubleWilson: number of ASUs (length of r) */
  int32_t optimize_dw_r;      /* --optimize-double-wilson-r */
  int32_t scale_bijector;     /* CLB_BIJ_* */
  float   scale_shift;        /* additive tfb.Shift(scale_multiplier), scaling/nn.py:84-87; 0 = none */
  float   epsilon;            /* --epsilon (args/common.py:38-42) */
  int32_t use_kl_weight;      /* 0: 'sum' reduction (default); 1: kl_weight * mean, variational.py:172-177 */
  float   kl_weight;

  /* tf_keras Adam built at manager.py:494-501; defaults args/optimizer.py:5-45 */
  float   learning_rate, beta_1, beta_2, adam_epsilon;
  float   clipnorm, clipvalue, global_clipnorm;   /* <= 0 means "not set" */

  uint64_t seed;              /* Philox key for in-kernel draws (--seed, args/tf_options.py:50-54) */
  int32_t rank, world_size;   /* reflection-partitioned data parallelism; 0,1 on one GPU */
  int32_t image_layers;       /* NeuralImageScaler per-image dense layers (--image-layers, args/scaling.py:33-37); needs n_images */
  int32_t deterministic;      /* 1: bitwise run-to-run reproducible steps -- no floating-point atomics whose order could vary: dL/dz_f is
                                 reduced per reflection in a fixed row order, weight-gradient partials are exclusive per CTA with
                                 one writer per address, scalar sums are added in block order.  Supported for MLPScaler models with
                                 the Wilson prior (no image scales / image layers / Ev11 / DoubleWilson); costs time, see DESIGN.md */
  int32_t refine_uncertainties; /* Ev11 error model (--refine-uncertainties, likelihoods/mono.py:39-73, laue.py:49-65): group
                                 CLB_GROUP_LIKELIHOOD holds the raw (pre-softplus) Sdfac, Sdadd, SdB, each initialised to 1 */
} clb_config;

/* Per-step metrics: the keys of the history dict returned by train_model (variational.py:214-224, 262-268). */
typedef struct {
  double loss;        /* "loss"      = kl_term - log-likelihood term */
  double nll;         /* "NLL"       */
  double kl;          /* "F KLDiv"   */
  double grad_norm;   /* "Grad Norm" (global norm BEFORE the non-finite filter, :205) */
} clb_metrics;

int clb_abi_version(void);
const char* clb_last_error(const clb_handle* h);   /* h may be NULL: error of the last failed clb_create */

/* Replaces: VariationalMergingModel.__init__ + DataManager.build_model + model.compile
 * (variational.py:15-45, manager.py:380-507).  Parameters start at the reference's
 * initial values (identity MLP, image scales 1); the surrogate is initialised by
 * clb_set_prior (prior mean / stddev, manager.py:432-436). */
int clb_create(const clb_config* cfg, clb_handle** out);
void clb_destroy(clb_handle* h);

/* Replaces: tf.convert_to_tensor of the input tuple at careless/careless.py:62 and the
 * accessors of careless/models/base.py:22-121.  Arrays are in the REFERENCE layout
 * (formatter.py:382-400 / :631-653): int64 ids of shape (N,[1]), metadata (N,d) row-major
 * float32, intensities/uncertainties float32 (Laue: the n_spots per-spot values first, then
 * 1.0 padding).  harmonic_id == NULL for mono.  obs_index (optional, NULL = 0..N-1) is the
 * GLOBAL row index of each observation (used as RNG counter and to look up injected draws);
 * refl_id indexes this handle's n_refl surrogate entries.
 * The call sorts/pads the rows into the device layout.  By default that happens ON THE GPU
 * (csrc/clb_prep.cuh: range checks, stable radix sort by refl_id / harmonic_id / image_id, padding, gather) --
 * the counterpart of the grouping the reference's formatter does with pandas on the host
 * (io/formatter.py:145, :617); CLB_DEVICE_PREP=0 or the deterministic mode use the host version
 * (clb_prepare_rows), which produces bit-identical rows. */
int clb_set_observations(clb_handle* h, int64_t n_rows, int64_t n_rows_total,
                         const int64_t* refl_id, const int64_t* image_id,
                         const float* metadata, const float* intensities, const float* uncertainties,
                         const int64_t* harmonic_id, const int64_t* obs_index, int32_t order);
/* The host prep of clb_set_observations alone (no CUDA, usable without a GPU): writes the sorted /
 * padded SoA rows into caller arrays of `capacity` rows (meta_out: n_meta x capacity is NOT the
 * layout -- it is n_meta x *n_padded, row stride *n_padded).  Call once with capacity 0 to learn
 * *n_padded.  Exists so the layout (refl_id order, harmonic grouping) can be checked bit-exactly. */
int clb_prepare_rows(int64_t n_rows, int64_t n_refl, int32_t n_meta, int32_t n_images, int32_t laue,
                     int32_t likelihood, float dof,
                     const int64_t* refl_id, const int64_t* image_id, const float* metadata,
                     const float* intensities, const float* uncertainties, const int64_t* harmonic_id,
                     const int64_t* obs_index, int32_t order, int32_t image_tile /* > 0: image layers, rows per tile */,
                     int64_t capacity, int64_t* n_padded, int32_t* refl_out, int32_t* image_out, int32_t* spot_out,
                     uint32_t* oidx_out, float* meta_out, float* iobs_out, float* sig_out, double* ll_const);
/* Copy of the device-resident prepared rows back to the caller (same arrays and conventions as clb_prepare_rows'
 * outputs): lets tests compare the device-side preparation with the host one bit by bit.  prep_ms (may be NULL)
 * receives the wall time of the row preparation inside the last clb_set_observations. */
int clb_download_rows(clb_handle* h, int64_t capacity, int64_t* n_padded, int32_t* refl_out, int32_t* image_out,
                      int32_t* spot_out, uint32_t* oidx_out, float* meta_out, float* iobs_out, float* sig_out,
                      double* ll_const, double* prep_ms);
/* Re-upload of the already prepared (pinned) device-layout rows: the host->device copy of
 * one step's inputs, used by the end-to-end measurement. */
int clb_upload_observations(clb_handle* h);
/* Input pipeline: start copying the prepared rows into the handle's SECOND device buffer on a copy stream and return;
 * the next clb_step / clb_step_begin / clb_eval waits for that copy on the device and switches buffers.  Called once
 * per step after the step's kernels have been launched (clb_step_begin ... clb_step_end without metrics), the copy of
 * step t+1's inputs overlaps the compute of step t (the tf.data prefetch a keras fit() loop would do; the reference's
 * train_model keeps its single batch resident, careless.py:61-70). */
int clb_prefetch_observations(clb_handle* h);

/* Replaces: WilsonPrior / DoubleWilsonPrior construction (priors/wilson.py:29-49, 82-138) and
 * the surrogate initialisation of manager.py:432-436.  centric (R) uint8, multiplicity (R),
 * sigma (R) Wilson Sigma.  DoubleWilson only (else NULL): dw_parent (R) int32 = surrogate index
 * of the parent reflection, -1 = parent absent, -2 = root entry; asu_id (R) int32; r (n_asu).
 * refl_index (optional) = global reflection index per entry (RNG counter).  init_scale is
 * --structure-factor-init-scale; pass a negative value to leave the surrogate untouched. */
int clb_set_prior(clb_handle* h, const uint8_t* centric, const float* multiplicity, const float* sigma,
                  const int32_t* dw_parent, const int32_t* asu_id, const float* r,
                  const int64_t* refl_index, float init_scale);

/* Replaces: keras get_weights/set_weights, save_weights/load_weights (careless.py:48-56,79-80)
 * and `.trainable = False` (careless.py:50-56,104). */
int64_t clb_group_size(const clb_handle* h, int32_t group);
int clb_get_params(clb_handle* h, int32_t group, float* out, int64_t n);
int clb_set_params(clb_handle* h, int32_t group, const float* in, int64_t n);
int clb_get_grads(clb_handle* h, int32_t group, float* out, int64_t n);      /* of the last step (pre-filter) */
int clb_get_adam_state(clb_handle* h, int32_t group, float* m, float* v, int64_t n, int64_t* t);
int clb_set_trainable(clb_handle* h, int32_t group, int32_t trainable);

/* Replaces: the hot loop train_model -> train_step_with_gradient_norm
 * (variational.py:255-256, :185-224): n_steps full-batch ELBO gradient + Adam steps.
 * inj_u_f  (n_steps, S, R)        uniforms in (0,1) for the truncated-normal surrogate, or NULL
 * inj_eps_s(n_steps, S, N_total)  standard normals for the scale sample, indexed by obs_index, or NULL
 * (NULL => in-kernel Philox4x32-10 keyed by seed/step/index).  metrics_out (n_steps) may be NULL.
 * Stops early after a step whose gradient norm is non-finite (:271-274); returns the number
 * of steps taken in *steps_done (may be NULL). */
int clb_step(clb_handle* h, int32_t n_steps, const float* inj_u_f, const float* inj_eps_s,
             clb_metrics* metrics_out, int32_t* steps_done);

/* Replaces: keras test_on_batch on held-out data (variational.py:257-260): forward pass only -- fresh draws,
 * no gradients, no update -- of the CURRENT parameters on this handle's observations.  Typical use: a second
 * handle holds the validation rows and receives the training handle's parameters through clb_set_params. */
int clb_eval(clb_handle* h, const float* inj_u_f, const float* inj_eps_s, clb_metrics* metrics_out);

/* Multi-GPU form of one step, split around the two all-reduces (sum) over NCCL:
 *   clb_step_begin  -> kernels up to the local gradients
 *   (caller all-reduces the float32 buffer: gradients of the replicated groups MLP / image scales / r)
 *   clb_step_norms  -> per-variable gradient norms; packs {sum(log q - log p), sum(ll), norms} as float64
 *   (caller all-reduces the float64 scalar buffer)
 *   clb_step_end    -> global grad norm, non-finite filter, clipping, Adam; metrics
 * Both buffers (device pointers, owned by the handle) come from clb_reduce_buffers.
 * With world_size == 1, clb_step() is exactly begin + norms + end without the all-reduces. */
int clb_step_begin(clb_handle* h, const float* inj_u_f, const float* inj_eps_s);
int clb_step_norms(clb_handle* h);
int clb_step_end(clb_handle* h, clb_metrics* metrics_out);
int clb_reduce_buffers(clb_handle* h, void** grads_f32, int64_t* n_f32, void** scalars_f64, int64_t* n_f64);

/* In-library exchange step (B200-native addition; the reference is single-device, careless/parser.py:25-40).  Rank 0
 * obtains an id (128 bytes, an ncclUniqueId) and hands it to the other ranks by any side channel (the Python layer
 * broadcasts it over torch.distributed / a TCP store); every rank then calls clb_comm_init on its handle (created with its
 * rank / world_size).  From then on clb_step(n) runs n steps on world_size GPUs with no host code between the steps:
 * clb_step_norms issues ONE grouped NCCL all-reduce {replicated gradients f32 | scalars + local norms f64} on the
 * handle's stream, and the caller-driven form is just begin -> norms -> end.  NCCL is bound with dlopen at the first
 * call; a process that never calls these needs no NCCL. */
int clb_comm_unique_id(uint8_t* id128);
int clb_comm_init(clb_handle* h, const uint8_t* id128);

/* Debug / parity hooks (variational.py:154,167): sampled structure factors z_f (S,R) and
 * predicted intensities ipred (S,N) in the caller's original row order, from the last step. */
int clb_get_samples(clb_handle* h, float* z_f, int64_t n);
int clb_enable_ipred(clb_handle* h, int32_t enable);
int clb_get_ipred(clb_handle* h, float* ipred, int64_t n);
int clb_synchronize(clb_handle* h);

/* "Next" rows of the scope table: the numeric part of DataManager.get_results
 * (io/manager.py:188-197, :209) and the scale moments behind get_predictions (variational.py:47-121). */
int clb_get_results(clb_handle* h, float* F, float* SigF, float* I, float* SigI, float* N, int64_t n_refl);
int clb_get_scale_moments(clb_handle* h, float* mean, float* stddev, int64_t n_rows_total);

/* Device timing of the dominant kernel (CUDA events on the handle's stream), for bench.py's roofline. */
int clb_kernel_time_ms(clb_handle* h, double* obs_kernel_ms_sum, int64_t* obs_kernel_launches, int64_t* total_launches);
int clb_reset_timers(clb_handle* h, int32_t enable_event_timing);

#ifdef __cplusplus
}
#endif
#endif /* CARELESS_B200_H */
