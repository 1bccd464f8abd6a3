/*
 * oat_b200.h — C-ABI of the B200-native OATomobile RIP/DIM hot path.
 *
 * The reference (OATML/oatomobile) has no FFI: its boundary for this path is
 * the Python class API of `oatomobile/baselines/torch/__init__.py:17-21`.
 * Every entry point below names the reference function it replaces
 * (paths relative to the reference root).  `oatomobile_b200/` binds these
 * with ctypes and re-exposes the reference's Python classes on top; the
 * binding a reference maintainer would add is shown in INTEGRATION.md.
 *
 * Conventions
 *  - plain C types only; every tensor argument is a *device* pointer to
 *    contiguous float32 unless its name starts with `h_` (host pointer);
 *  - `stream` is a `cudaStream_t` passed as `void*` (0 = legacy default stream);
 *  - every function returns 0 on success, non-zero on failure;
 *    `oat_last_error()` then returns a thread-local message;
 *  - kernels are enqueued on `stream` and NOT synchronised;
 *  - there is no CPU fallback: without a CUDA device every call fails.
 */
#ifndef OAT_B200_H_
#define OAT_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define OAT_API __attribute__((visibility("default")))
#else
#define OAT_API
#endif

#define OAT_ABI_VERSION 1
#define OAT_HIDDEN 64        /* GRU hidden size (sequence.py:36, dim/model.py:67) */
#define OAT_ENC_FEATURES 128 /* encoder output width (dim/model.py:53)            */

/* A named host tensor: one entry of a reference `state_dict()`.            */
typedef struct OatTensor {
  const char* name;  /* e.g. "_encoder._model.features.0.0.weight"          */
  const void* h_data; /* host pointer, float32 (int64 entries are ignored)   */
  int32_t ndim;
  int64_t shape[4];
} OatTensor;

typedef struct OatModel OatModel;       /* one ImitativeModel / BehaviouralModel */
typedef struct OatEnsemble OatEnsemble; /* E models on one GPU + workspace       */

enum { OAT_KIND_DIM = 0, OAT_KIND_CIL = 1, OAT_KIND_FLOW = 2 /* AutoregressiveFlow alone */,
       OAT_KIND_ENCODER = 3 /* MobileNetV2 alone: keys `_model.features...`, `_model.classifier.1` */ };
enum { OAT_ALGO_WCM = 0, OAT_ALGO_BCM = 1, OAT_ALGO_MA = 2 };

OAT_API const char* oat_last_error(void);
OAT_API int oat_abi_version(void);

/* Replaces `ImitativeModel.__init__ + load_state_dict` (dim/model.py:39-68) and
 * `BehaviouralModel.__init__` (cil/model.py:34-66) on the device side: takes the
 * reference state_dict (328 / 326 entries), folds the 52 eval-mode BatchNorms
 * into their convolutions, re-lays the weights out for the kernels and uploads
 * them to `device`.  `in_channels` is read from the stem's shape.  With
 * OAT_KIND_FLOW only the decoder is packed, from the keys of a stand-alone
 * `AutoregressiveFlow` (`_decoder.weight_ih`, `_locscale._model.0.weight`, ...;
 * oatomobile/torch/networks/sequence.py:53-65).                                  */
OAT_API int oat_model_create(const OatTensor* tensors, int32_t num_tensors, int32_t kind,
                     int32_t device, OatModel** out);
OAT_API int oat_model_destroy(OatModel* model);
OAT_API int oat_model_in_channels(const OatModel* model);

/* Groups E models that live on this GPU (RIPAgent.__init__, rip/agent.py:49-50).
 * Owns the activation workspace, grown on demand by `oat_ensemble_reserve`
 * (called implicitly by oat_encode).  All models share kind, device and in_channels.
 * An ensemble of OAT_KIND_FLOW models holds decoders only (the replicas a rank keeps of
 * every model's AutoregressiveFlow when the flow stage is sharded by scenes): usable with
 * oat_rip_sample_score, rejected by oat_encode / oat_encode_features.  With
 * OAT_KIND_ENCODER models only oat_encode_features applies.                          */
OAT_API int oat_ensemble_create(OatModel* const* models, int32_t num_models, OatEnsemble** out);
OAT_API int oat_ensemble_destroy(OatEnsemble* ens);
OAT_API int oat_ensemble_reserve(OatEnsemble* ens, int32_t batch);

/* Selects the kernel family of the 34 pointwise convolutions + classifier:
 * 1 = tcgen05/TMA tensor-core GEMM with 3xTF32 error compensation (default),
 * 0 = FP32 SIMT GEMM.  Both meet the 1e-4 parity bar; see DESIGN.md. */
OAT_API int oat_ensemble_set_pw_impl(OatEnsemble* ens, int32_t impl);

/* Selects which of the encoder's first blocks run as fused kernels (fused.cu): bit 0 =
 * features.0 + features.1 (stem, depthwise, project) in one kernel; bits 1..3 = expand 1x1 +
 * depthwise 3x3 of features.2 / .3 / .4 in one kernel, the 6x expanded tensor staying in
 * shared memory; bit 4 = depthwise 3x3 + project of features.1 in one kernel (ignored when
 * bit 0 is set); bit 5 = expand 1x1 + depthwise 3x3 of features.5-17 inside the tcgen05 GEMM
 * (the depthwise window slides over the slab its epilogue stages in shared memory; needs the
 * tcgen05 pointwise family; correct but measured SLOWER than the separate launches on B200 —
 * the depthwise arithmetic lands on the GEMM's eight epilogue warps — so it is opt-in).
 * Results agree with the unfused path to rounding.
 * Default: 30 (OAT_FUSE_DEFAULT), or the environment variable OAT_FUSE.               */
OAT_API int oat_ensemble_set_fusion(OatEnsemble* ens, int32_t mask);
OAT_API int oat_ensemble_get_fusion(const OatEnsemble* ens);
/* Kernel family of the fused expand+depthwise blocks when the pointwise family is tcgen05:
 * 0 = FP32 FMA kernel, 1 = auto (default: the pipelined tcgen05 kernel for features.2, the
 * FP32 kernel for features.3/4 — whichever measured faster on B200), 2 = tcgen05 for all. */
OAT_API int oat_ensemble_set_fusion_tc(OatEnsemble* ens, int32_t mode);

/* Selects the flow kernel family process-wide: 1 = tcgen05 3xTF32 recurrent GEMMs
 * with the state resident in shared/tensor memory (default), 0 = FP32 SIMT.        */
OAT_API int oat_set_flow_impl(int32_t impl);

/* transforms.downsample_visual_features + transpose_visual_features
 * (oatomobile/torch/transforms.py:34-49, called from dim/model.py:245-251):
 * lidar [B,C,H,W] -> visual [B,C,100,100], bilinear align_corners=True then H<->W. */
OAT_API int oat_transform_visual(const float* lidar, int32_t B, int32_t C, int32_t H, int32_t W,
                         float* visual, void* stream);

/* Same, reading the simulator / on-disk layout lidar [B,H,W,C] directly, i.e. fused with
 * the HWC->CHW transposes of rip/agent.py:69 and datasets/carla.py:138-140.           */
OAT_API int oat_transform_visual_hwc(const float* lidar, int32_t B, int32_t H, int32_t W, int32_t C,
                             float* visual, void* stream);

/* `MobileNetV2.forward` (oatomobile/torch/networks/perception.py:53-55) for every model of
 * the ensemble, eval mode: visual [B,C,100,100] -> features [E,B,128] (stem, 17 inverted-
 * residual blocks, 1x1 320->1280, global average pool, Linear 1280->128; no merger).  Works
 * for ensembles of any kind, including OAT_KIND_ENCODER (a MobileNetV2 module on its own). */
OAT_API int oat_encode_features(OatEnsemble* ens, const float* visual, int32_t B, float* features,
                        void* stream);

/* `MLP.forward` (oatomobile/torch/networks/mlp.py:70-72) for a ReLU stack: layer l is
 * Linear(sizes[l] -> sizes[l+1]) with DEVICE pointers weights[l] = W [sizes[l+1]][sizes[l]]
 * (PyTorch layout) and biases[l] (or null); ReLU after every layer but the last, and after
 * the last iff `activate_final` (mlp.py:65-66).  x [B,sizes[0]] -> out [B,sizes[num_layers]].
 * One launch; 1 <= num_layers <= 8, widths <= 4096.  The pointer arrays live on the host. */
OAT_API int oat_mlp_forward(const float* const* weights, const float* const* biases,
                    const int32_t* sizes, int32_t num_layers, int32_t activate_final,
                    const float* x, int32_t B, float* out, void* stream);

/* `ImitativeModel._params` (dim/model.py:173-219) for every model of the ensemble:
 * visual [B,C,100,100] (shared), scalars [B,S] = cat(velocity(3), is_at_traffic_light(1),
 * traffic_light_state(1) [, mode(1) for CIL]) -> z [E,B,64].                      */
OAT_API int oat_encode(OatEnsemble* ens, const float* visual, const float* scalars,
               int32_t B, float* z, void* stream);

/* `AutoregressiveFlow._forward` (oatomobile/torch/networks/sequence.py:95-151):
 * x [N,T,2], z [N/rows_per_z,64] -> y [N,T,2], logabsdet [N] (may be NULL).
 * Row n uses z[n / rows_per_z] (rows_per_z = 1 reproduces the reference call). */
OAT_API int oat_flow_forward(const OatModel* model, const float* x, const float* z,
                     int64_t N, int32_t T, int32_t rows_per_z,
                     float* y, float* logabsdet, void* stream);

/* `AutoregressiveFlow._inverse` (sequence.py:153-216): y [N,T,2] ->
 * x [N,T,2] (may be NULL), log_prob [N], logabsdet [N].                          */
OAT_API int oat_flow_inverse(const OatModel* model, const float* y, const float* z,
                     int64_t N, int32_t T, int32_t rows_per_z,
                     float* x, float* log_prob, float* logabsdet, void* stream);

/* K-sample sample-and-score (BASELINE.json metric; SURVEY.md 3.5, assembled from
 * rip/agent.py:106-119,137): proposals y = f_p(x; z_p) from local model
 * `proposal_idx` (pass -1 when `y` is an INPUT produced elsewhere, e.g. by the
 * rank that owns model 0), then q[m,b,k] = log_prob_m - logabsdet_m
 * (+ per-sample goal log-likelihood, dim/model.py:143-171, when goal != NULL).
 * z [E,B,64], x [B,K,T,2], goal [B,G,2] or NULL, y [B,K,T,2], q [E,B,K].        */
OAT_API int oat_rip_sample_score(OatEnsemble* ens, int32_t proposal_idx, const float* z,
                         const float* x, const float* goal, int32_t G, float epsilon,
                         int32_t B, int32_t K, int32_t T,
                         float* y, float* q, void* stream);

/* Ensemble aggregation + plan selection (rip/agent.py:121-127,137):
 * s[b,k] = min_m(-q) ("WCM") | max_m(-q) ("BCM") | mean_m(-q) ("MA") as written in
 * the reference; kstar[b] = argmin_k s (lowest k on ties); plan[b] = y[b,kstar[b]].
 * q [E,B,K]; s [B,K] (may be NULL); kstar int32 [B]; sbest [B]; plan [B,T,2].    */
OAT_API int oat_rip_aggregate(const float* q, int32_t E, int32_t B, int32_t K, int32_t algo,
                      const float* y, int32_t T, float* s, int32_t* kstar,
                      float* sbest, float* plan, void* stream);

/* Gradient-based MAP planner, ONE kernel launch: the Adam-on-latent loops of
 * `ImitativeModel.forward` (dim/model.py:97-141; algo = -1, one model) and
 * `RIPAgent.__call__` (rip/agent.py:84-137; algo = OAT_ALGO_*, E models): per step
 * y = f_0(x; z_0), per-model posteriors mean_b(log_prob - logabsdet) + goal likelihood,
 * loss aggregation as written, analytic back-propagation to x, torch.optim.Adam update
 * (lr, betas 0.9/0.999, eps 1e-8), best-x bookkeeping with the post-step x (as written),
 * finally plan = f_0(x_best).  x [B,T,2] in: initial latent, out: final latent;
 * x_best/plan [B,T,2] out; z [E,B,64]; goal [B,G,2] or NULL; loss_out [num_steps] or
 * NULL; workspace: oat_plan_workspace_floats(B, E, T) floats of device scratch.    */
OAT_API int oat_plan(OatModel* const* models, int32_t num_models, int32_t algo, const float* z,
             const float* goal, int32_t G, float epsilon, int32_t B, int32_t T,
             int32_t num_steps, float lr, float* x, float* x_best, float* plan,
             float* workspace, int64_t workspace_floats, float* loss_out, void* stream);
OAT_API int64_t oat_plan_workspace_floats(int32_t B, int32_t num_models, int32_t T);

/* `ImitativeModel._goal_likelihood` (dim/model.py:143-171): y_last [B,2] = y[:, -1],
 * goal [B,G,2] -> rows [B] (per-row log-likelihood, may be NULL) and mean [1].       */
OAT_API int oat_goal_likelihood(const float* y_last, const float* goal, int32_t B, int32_t G,
                        float epsilon, float* rows, float* mean, void* stream);

/* `carla_lidar_measurement_to_ndarray` (oatomobile/utils/carla.py:165-233): point cloud
 * points [N,3] -> BEV histogram bev [200,200,2] float32 in {0,.2,..,1} (x, y, below|above
 * z = -2.5), bit-exact with the NumPy reference.  counts: 80000 uint32 of scratch.      */
OAT_API int oat_lidar_bev(const float* points, int64_t num_points, int32_t pixels_per_meter,
                  int32_t hist_max_per_pixel, int32_t meters_max, uint32_t* counts, float* bev,
                  void* stream);

/* `BehaviouralModel.forward` roll-out (cil/model.py:106-127) after the encoder:
 * z [B,64] -> y [B,T,2].                                                          */
OAT_API int oat_cil_rollout(const OatModel* model, const float* z, int32_t B, int32_t T,
                    float* y, void* stream);

/* TEST HOOK (tests/test_gpu_tc_gemm.py): one grouped pointwise GEMM on the tensor
 * cores, C[e,m,n] = act(sum_k A[e,m,k] W[e,n,k] + bias[e,n]) (+ R).  Allocates and
 * frees its TF32-split weight copies; synchronises the stream.                   */
OAT_API int oat_debug_tc_gemm(const float* A, const float* W, const float* bias, const float* R,
                      float* C, int32_t M, int32_t K, int32_t N, int32_t E, int32_t relu6,
                      void* stream);

/* TEST HOOK (tests/test_gpu_fused.py): the encoder up to and including `blocks`
 * inverted-residual blocks (0 = the stem alone; with fusion bit 0 set the first available
 * prefix is blocks = 1): visual [B,C,100,100] -> out [E][B][h][h][c] (NHWC).         */
OAT_API int oat_debug_encoder_prefix(OatEnsemble* ens, const float* visual, int32_t B,
                                     int32_t blocks, float* out, void* stream);

/* ---- training step ----------------------------------------
 * A named DEVICE tensor of a model that is being trained: `param` points at the
 * storage of the reference `state_dict` entry (reference layout, float32), `grad`
 * at the storage of its gradient (NULL for BatchNorm running statistics).       */
typedef struct OatTrainTensor {
  const char* name;
  void* param;
  void* grad;
  int32_t ndim;
  int64_t shape[4];
} OatTrainTensor;

typedef struct OatTrainer OatTrainer;

/* Replaces the model + optimiser set-up of the train scripts
 * (oatomobile/baselines/torch/dim/train.py:114-119, cil/train.py:112-118).
 * Binds the trainer to the caller-owned parameter / gradient storage of one
 * `ImitativeModel` (kind OAT_KIND_DIM) or `BehaviouralModel` (OAT_KIND_CIL).  If all
 * gradients live in one flat buffer pass it as (grad_flat, grad_flat_floats) so it is
 * cleared with one memset per step; otherwise pass NULL/0.                       */
OAT_API int oat_trainer_create(const OatTrainTensor* tensors, int32_t num_tensors, int32_t kind,
                       int32_t device, float* grad_flat, int64_t grad_flat_floats,
                       OatTrainer** out);
OAT_API int oat_trainer_destroy(OatTrainer* trainer);

/* Forward in `model.train()` mode + loss + backward of the reference `train_step`
 * up to (not including) `optimizer.step()`:
 *   DIM (dim/train.py:190-203): z = _params(batch); loss = -mean(log_prob - logabsdet)
 *       of `target` (the perturbed `player_future[..., :2]`) under `_decoder._inverse`;
 *   CIL (cil/train.py:176-184): loss = mean_b sum_{t,d} |forward(batch) - target|.
 * visual [B,C,100,100] (post-`transform`), scalars [B,5|6] = velocity(3) |
 * is_at_traffic_light | traffic_light_state (| mode), target [B,T,2], dropout_mask
 * [B,1280] = Bernoulli(0.8)/0.8 for the classifier's Dropout(0.2) or NULL (p = 0).
 * BatchNorm layers use batch statistics and update running_mean / running_var
 * (momentum 0.1) in place.  Every gradient is overwritten.  Outputs: loss [1];
 * optional z [B,64] and (CIL) pred [B,T,2].                                       */
OAT_API int oat_train_forward_backward(OatTrainer* trainer, const float* visual, const float* scalars,
                               const float* target, const float* dropout_mask, int32_t B,
                               int32_t T, float* loss, float* z, float* pred, void* stream);

/* Debug/test access to the post-activation outputs of the last
 * oat_train_forward_backward call: index 0..51 = the 52 conv+BN units in execution order,
 * NHWC rows [B*h*h][channels]; 52..54 = the three merger layers [B][64].  Always reports
 * the shape; when `out` is not NULL the activation is copied (device to device, on
 * `stream`) into it.  The gradient tests use this to evaluate the float64 oracle on the
 * same ReLU/ReLU6 branch the kernels took.                                          */
OAT_API int oat_trainer_activation(const OatTrainer* trainer, int32_t index, float* out,
                           int64_t* rows, int32_t* channels, void* stream);

/* `torch.optim.Adam.step()` (+ L2 `weight_decay`) over one flat parameter range, with
 * gradients first multiplied by `grad_scale` (1/world_size after an all-reduce SUM) and
 * optionally clipped to global norm `clip_norm` (> 0; `clip_grad_norm_`,
 * dim/train.py:207-208).  `step` is the 1-based step count; `norm_ws` is one double of
 * device scratch (needed only when clipping).                                     */
OAT_API int oat_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                  int32_t step, float lr, float beta1, float beta2, float eps,
                  float weight_decay, float grad_scale, float clip_norm, double* norm_ws,
                  void* stream);

/* Number of kernel launches issued by this library since load (bench.py's
 * `gpu_launches`; no reference counterpart). */
OAT_API int64_t oat_launch_count(void);

/* Per-kernel-family device timing for bench.py's `roofline` (no reference counterpart: the
 * reference has no profiler hook).  Between begin and end every kernel this
 * library launches is followed by a cudaEventRecord on its stream; `oat_profile_end`
 * synchronises, attributes the time between consecutive events to the launch in between and
 * writes {"<family>": {"ms": total, "launches": n}, ...} as NUL-terminated JSON text.
 * Launch the kernels one by one while a profile is open (not inside a stream capture).   */
OAT_API int oat_profile_begin(void* stream);
OAT_API int oat_profile_end(char* json, int64_t capacity);

#ifdef __cplusplus
}
#endif
#endif /* OAT_B200_H_ */
