/*
 * mpl_b200.h — C ABI of the B200-native MPL lifter forward.
 *
 * One shared library (libmpl_b200.so, hand-written CUDA for sm_100a) replaces the PyTorch-eager forward of the
 * reference's lifting network.  Every entry point cites the reference interface it stands in for
 * (paths relative to the reference checkout, aghasemzadeh/OpenMPL):
 *
 *   MPL/lib/models/multiview_mpl.py:95-317   MultiView_MPL.__init__  (constructor kwargs  -> MplDesc / mpl_create)
 *   MPL/lib/models/multiview_mpl.py:450-525  MultiView_MPL.forward   (poses, rays, centers -> mpl_forward)
 *   MPL/lib/utils/utils.py:148-153           load_state_dict of checkpoints (-> mpl_param_* / mpl_pack_weights)
 *   MPL/lib/core/evaluate.py:91-125          calc_mpjpe / calc_distance_per_dim (-> mpl_mpjpe_accumulate)
 *   MPL/lib/utils/pose_utils.py:61-143       PoseUtils.procrustes (-> mpl_pmpjpe_accumulate)
 *   MPL/lib/dataset/joints_dataset_mpl.py:615-648,701-715,762-772,817-820,872-904
 *                                            per-sample input construction (-> mpl_build_inputs)
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / C++ types cross this boundary;
 *   - the library never allocates user-visible device memory: inputs, outputs, the packed weight blob and the
 *     workspace are caller-owned device buffers, borrowed for the duration of the call;
 *   - all work is enqueued on the caller's stream on the caller's current device; no hidden synchronisation;
 *   - every function returns MPL_OK (0) or a negative MplStatus; mpl_last_error() gives the thread-local message;
 *   - entry points are re-entrant; a handle may be used from one thread at a time (one handle per device replica,
 *     like one nn.Module replica per GPU in the reference's DataParallel, MPL/run/valid_mpl.py:177-178).
 */
#ifndef MPL_B200_H_
#define MPL_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MPL_ABI_VERSION 2

typedef enum MplStatus {
  MPL_OK = 0,
  MPL_ERR_INVALID_ARGUMENT = -1, /* bad pointer / size / struct_size                                             */
  MPL_ERR_CONFIG_RUNTIME = -2,   /* flag combination whose first forward raises RuntimeError in the reference    */
  MPL_ERR_CONFIG_INDEX = -3,     /* flag combination whose first forward raises IndexError in the reference      */
  MPL_ERR_CUDA = -4,             /* a CUDA runtime / driver call failed                                           */
  MPL_ERR_WORKSPACE = -5,        /* workspace or packed buffer too small                                          */
  MPL_ERR_UNSUPPORTED = -6       /* requested precision mode cannot serve this shape (no silent fallback)         */
} MplStatus;

/* Arithmetic of the Linear layers.  LayerNorm, softmax, residual stream and all accumulators are fp32 in every mode. */
typedef enum MplPrecision {
  MPL_PREC_FP32 = 0, /* fp32 operands, fp32 FMA (CUDA cores)                                        */
  MPL_PREC_TF32 = 1, /* the fp32-grade tensor-core mode (the "fp32/TF32 path" of the parity contract): FPT projections
                        on tcgen05 with every operand split into two bf16 planes (hi + lo, 16 significand bits) and
                        hi.hi + hi.lo + lo.hi accumulated in fp32 -- single-pass kind::tf32 (11 bits) cannot hold the
                        1e-3 bound on every shipped configuration; LayerNorm / attention / head fp32               */
  MPL_PREC_BF16 = 2  /* FPT projections: tcgen05 kind::f16 bf16 operands; SPT: bf16 mma; fp32 accum */
} MplPrecision;

/* POD mirror of the keyword arguments of MultiView_MPL.__init__ (multiview_mpl.py:95-117).
 * Dropout / drop-path rates are omitted: the path is inference only (identity in eval, multiview_mpl.py:79). */
typedef struct MplDesc {
  int32_t struct_size; /* = sizeof(MplDesc); ABI guard */
  int32_t num_joints;
  int32_t in_chans;
  int32_t embed_dim_ratio;
  int32_t depth;
  int32_t num_heads;
  int32_t num_views;
  int32_t hidden_dim;
  float mlp_ratio;
  float qk_scale; /* 0 = None -> head_dim^-0.5 (multiview_mpl.py:46) */
  int32_t qkv_bias;
  int32_t add_confidence_input;
  int32_t mult_confidence_emb;
  int32_t concat_confidence_emb;
  int32_t confidence_input_as_third;
  int32_t pose_3d_emb_learnable;
  int32_t linear_weighted_mean;
  int32_t add_3D_pos_encoding_in_Spatial;
  int32_t input_rays_as_token;
  int32_t add_3D_pos_encoding_to_rays;
  int32_t confidence_as_attention_uncertainty_weight;
  int32_t multiple_spatial_blocks;
  int32_t no_transformer_spt;
  int32_t no_transformer_fpt;
  int32_t confidence_in_FPT;
  int32_t deep_head;
  int32_t head_kadkhod;
  int32_t FPT_blocks_view_keypoint_tokens;
  int32_t precision; /* MplPrecision */
  /* implementation switches (not reference keywords); 0-initialised fields select the defaults */
  int32_t ln_fusion;      /* bf16 mode: fold the FPT LayerNorms into the projection GEMMs.  0 = separate LayerNorm kernels,
                             non-zero (make_desc default: 1) = fused: the GEMM that updates the residual stream also emits
                             the next GEMM's operand and per-row (sum, sum^2); the next GEMM multiplies the raw rows by
                             W diag(gamma) and applies (mean, rstd) in its epilogue */
  int32_t gemm_cta_group; /* 0 / 2: CTA pair per 256x256 tile (cta_group::2, cluster 2x1x1); 1: one CTA per 128x256 tile */
  int32_t spt_hidden, fpt_hidden; /* int(width * mlp_ratio) of the SPT / FPT Mlp as the reference's float64 arithmetic gives it
                                     (0 = derive from the float mlp_ratio above) */
  int32_t qkv_attn_fusion; /* bf16 LayerNorm-fused mode, view tokens, 136-wide heads (or an even number of 68-wide ones), 2 to 8 views: QKV projection and
                              cross-view attention run as ONE kernel (no q|k|v tensor).  0 = two kernels, non-zero (make_desc
                              default: 1) = fused where the shape allows */
  int32_t chunk_streams;  /* 0 / 1: one pose chunk at a time on the caller's stream (default).  2: batches of >= 16384 poses
                             run as two interleaved pose chunks on two internal streams (forked from / joined to the
                             caller's stream) so that kernels of one chunk may overlap those of the other; the
                             workspace doubles.  Same results; measured neutral on B200 (profiles/r2_experiments.md) */
} MplDesc;

typedef struct MplModel MplModel;          /* opaque host-side handle */
typedef struct CUstream_st* mpl_stream_t;  /* == cudaStream_t */

/* Thread-local message of the last failing call on this thread ("" if none). */
const char* mpl_last_error(void);
int mpl_abi_version(void);

/* MultiView_MPL.__init__ (multiview_mpl.py:95-317).  Host only.  Invalid flag combinations fail here with the
 * status class of the exception the reference raises at its first forward. */
int mpl_create(const MplDesc* desc, MplModel** out);
void mpl_destroy(MplModel* m);

/* state_dict contract (utils.py:148-153): entry i has the reference's key (without the `features.` prefix of
 * MultiView_MPL_G) and element count; order = registration order of the reference module. */
int mpl_num_params(const MplModel* m);
int mpl_param_info(const MplModel* m, int index, const char** name, int64_t* numel, int32_t* is_int64);

/* Derived dims (multiview_mpl.py:140-142,272-274): which = 0 tok_w, 1 fpt_dim, 2 fpt_tokens, 3 E, 4 spt_hidden,
 * 5 fpt_hidden, 6 outputs (1, or 3 for head_kadkhod), 7 output columns the LAST fc2 of the FPT stack computes (E where the
 * residual stream is kept channel-permuted and the head reads its pose half only, else fpt_dim). */
int64_t mpl_dim(const MplModel* m, int which);

/* Weight repack: fp32 state_dict tensors (device pointers, in mpl_param_info order) -> one caller-owned blob
 * holding what the kernels read (fp32 copies, BatchNorm folded into the preceding Linear, bf16 / tf32 operand
 * copies of the FPT projection matrices).  Call again after the parameters change. */
size_t mpl_packed_bytes(const MplModel* m);
int mpl_pack_weights(MplModel* m, const void* const* params, int num_params, void* packed, size_t packed_bytes,
                     mpl_stream_t stream);

/* Workspace for batches of up to `max_batch` poses (the forward processes larger batches in chunks of
 * mpl_chunk_poses(), so the workspace stops growing there). */
size_t mpl_workspace_bytes(const MplModel* m, int64_t max_batch);
int64_t mpl_chunk_poses(const MplModel* m);
int mpl_set_chunk_poses(MplModel* m, int64_t chunk);

/* MultiView_MPL.forward (multiview_mpl.py:450-525), eval mode.
 *   poses[v]   -> view v 2D joints (x, y, conf)  [B, J, 3] fp32, consecutive poses `pose_stride` floats apart
 *   rays[v]    -> view v ray points              [B, J, 3] fp32, same stride
 *   centers[v] -> view v camera centre           [B, 1, 3] fp32, consecutive poses `center_stride` floats apart
 *                 (the reference's list-of-views convention, core/function_mpl.py:344-350, is V separate
 *                  tensors: stride J*3 and 3; one packed [B,V,J,3] tensor is the same call with stride V*J*3)
 *   out        -> [B, J, 3] fp32;  aux1/aux2 -> the two intermediate predictions of head_kadkhod (else NULL)
 * The pointer arrays are host arrays of V device pointers. */
int mpl_forward(MplModel* m, const void* packed, const float* const* poses, const float* const* rays,
                const float* const* centers, int64_t pose_stride, int64_t center_stride, float* out, float* aux1,
                float* aux2, int64_t batch, void* workspace, size_t workspace_bytes, mpl_stream_t stream);

/* Small-batch path (the reference's runner calls the model at TEST.BATCH_SIZE = 256, configs/h36m/mpl_amass/hm_0_*.yaml:144,
 * MPL/run/valid_mpl.py:205-210, where ~130 launches per forward are latency-bound): with max_batch > 0, an mpl_forward of at most
 * max_batch poses is captured ONCE per (batch, packed, input / output / workspace pointers) as a CUDA graph and later calls
 * with the same arguments replay it with a single graph launch on the caller's stream.  The caller keeps those buffers
 * alive and at the same addresses (the nn.Module stages inputs in persistent buffers for this).  Up to 16 graphs are kept
 * per handle (least recently used evicted); 0 switches the path off and frees them.  A stream that is itself being
 * captured, and profiling mode, take the plain path. */
int mpl_set_graph_batch(MplModel* m, int64_t max_batch);
int mpl_graph_stats(const MplModel* m, int64_t* captures, int64_t* replays);

/* Number of kernel launches the last mpl_forward on this handle enqueued (bench.py's gpu_launches). */
int64_t mpl_last_launch_count(const MplModel* m);

/* Optional per-launch timing with CUDA events on the caller's stream (the reference's only profiling hook is an
 * unsynchronised time.time() around the model call, core/function_mpl.py:346-351).  While enabled, every kernel
 * launch of mpl_forward is bracketed by two events; mpl_profile_collect waits for the last forward's events and
 * returns milliseconds and launch counts per kernel category (mpl_profile_category_name).  enabled = 1 profiles the
 * production schedule (two chunks in flight: launches of the two streams overlap, so the per-category times sum to more
 * than the step); enabled = 2 additionally runs one chunk at a time, so the times add up to the (slower) serial step. */
int mpl_set_profile(MplModel* m, int enabled);
int mpl_profile_categories(void);
const char* mpl_profile_category_name(int category);
int mpl_profile_collect(MplModel* m, double* ms_per_category, int64_t* launches_per_category, int n);

/* calc_mpjpe / calc_distance_per_dim (evaluate.py:91-125) + the unit / root-centring / confidence-mask rules of
 * core/function_mpl.py:674-687, as running fp64 sums so ranks can be combined with one all-reduce.
 *   acc layout (doubles): [0,J) sum_b ||pred-gt|| per joint (absolute); [J,2J) same, root-relative;
 *   [2J,5J) sum_b |pred-gt| per joint-dim over unmasked entries (absolute); [5J,8J) same, root-relative;
 *   [8J,11J) unmasked count per joint-dim; [11J] pose count.
 *   conf3d may be NULL (no masking); unit_scale = 100 if OUTPUT_IN_METER else 1.
 *   room_affine: NULL, or a HOST array of 6 floats (scale x, y, z, offset x, y, z): the un-scaling validate() applies to the
 *   predictions and targets of room-normalised datasets before it stores them (function_mpl.py:476-488), v * scale + offset in
 *   fp32: (s, s, s, centre) for 'room_scaled_equal', (sx, sy, 1, 0, 0, 0) otherwise. */
#define MPL_METRIC_ACC_LEN(J) (11 * (J) + 1)
int mpl_mpjpe_accumulate(const float* pred, const float* gt, const float* conf3d, int64_t batch, int num_joints,
                         float unit_scale, const float* room_affine, double* acc, mpl_stream_t stream);

/* Procrustes-aligned error (P-MPJPE) as running fp64 sums: per pose, PoseUtils.procrustes(A = gt, B = pred)
 * (MPL/lib/utils/pose_utils.py:61-143) after the unit rule above, then the per-joint distances of calc_mpjpe between the
 * aligned prediction Z and gt.  scaling: 1 = similarity (the reference default), 0 = rigid.  reflection: -1 = 'best'
 * (the reference default: whatever the SVD gives), 0 = forbid, 1 = force (pose_utils.py:112-120).
 *   acc layout (doubles): [0,J) sum_b ||Z_j - gt_j||; [J] sum_b d (normalised residual); [J+1] sum_b scale;
 *   [J+2] pose count. */
#define MPL_PMETRIC_ACC_LEN(J) ((J) + 3)
int mpl_pmpjpe_accumulate(const float* pred, const float* gt, int64_t batch, int num_joints, float unit_scale,
                          int scaling, int reflection, double* acc, mpl_stream_t stream);

/* Per-sample input construction of the dataset (joints_dataset_mpl.py:615-648,701-715,762-772,817-820,872-904),
 * batched: raw detector output (u, v, conf) in pixels + per-view calibration -> the model's poses / rays / centers.
 *   pix   [B, V, J, 3] fp32 (u, v, conf);  calib [V, 18] fp64 = R (9, row-major world->cam), t (3), fx, fy, cx, cy, w, h
 *   poses / rays [B, V, J, 3], centers [B, V, 1, 3] fp32 (packed layout accepted by mpl_forward). */
int mpl_build_inputs(const float* pix, const double* calib, int64_t batch, int num_views, int num_joints, float* poses,
                     float* rays, float* centers, mpl_stream_t stream);

/* MHP-style synthetic data on the device (what MHP/utils.py:261-318 + MPL/lib/utils/calib.py:42-77 do to AMASS poses:
 * place a 3D pose in the room, project it through every calibration).  Pose i of `seed` depends only on (seed, start + i)
 * (numpy Philox4x64-10 stream, identical to openmpl_b200/synth.py), so any sharding of the global index range gives the
 * same data.  room [4] fp64 = (min_x, max_x, min_y, max_y); calib as in mpl_build_inputs.
 *   pix    [B, V, J, 3] fp32 raw pixels (u, v, conf) -- feed to mpl_build_inputs;  target [B, J, 3] fp32, metres. */
int mpl_synth_project(uint64_t seed, int64_t start, int64_t batch, int num_views, int num_joints, const double* calib,
                      const double* room, int conf_ones, float* pix, float* target, mpl_stream_t stream);

/* ---- unit-test hooks for the building blocks (used by tests/ only) ------------------------------------------- */
/* Y[M,N] = epilogue(A[M,K] . W[N,K]^T): the tcgen05 projection kernel in isolation.
 *   dtype: MPL_PREC_BF16 (A, W bf16) or MPL_PREC_TF32 (split mode: A = [2][M][K], W = [2][N][K] bf16 planes hi, lo)
 *   epilogue: 0 bias -> operand format (bf16, or two planes [2][M][N]) / fp32 per `out_fp32`; 1 bias+GELU;
 *             2 bias + residual(fp32, in place in Y)
 *   cta_group: 1 or 2 (see MplDesc.gemm_cta_group). */
int mpl_test_gemm(const void* A, const void* W, const float* bias, void* Y, int64_t M, int N, int K, int dtype,
                  int epilogue, int out_fp32, int cta_group, mpl_stream_t stream);

/* The LayerNorm-fused epilogues of the bf16 mode in isolation.
 *   epilogue 4 / 5 (LayerNorm-apply, QKV / fc1): Y[M,N] (bf16, or fp16 when out_fp16) = act(rstd * (A W'^T - mu * colsum) + bias),
 *       A = raw residual rows rounded to bf16, W' = bf16(W diag(gamma)), colsum[n] = sum_k W'[n,k], bias = b + W beta,
 *       (mu, rstd) from stats_in = [slots_in][M rounded up to 256] float2 partial (sum x, sum x^2) per row, eps = LayerNorm eps;
 *   epilogue 6 (residual-emit, proj / fc2): the residual stream x [M,N] as two bf16 planes (Y = hi, x_lo = lo), updated in place:
 *       x += A W^T + bias; stats_out = [mpl_test_gemm_ln_slots(N)][M rounded up to 256] float2 partial (sum, sum^2) of the new x;
 *       ab_fp16: A and W hold fp16 instead of bf16. */
int mpl_test_gemm_ln(const void* A, const void* W, const float* bias, void* Y, int64_t M, int N, int K, int epilogue,
                     const float* colsum, const void* stats_in, int slots_in, void* stats_out, void* x_lo, float eps,
                     int ab_fp16, int out_fp16, int cta_group, mpl_stream_t stream);
int mpl_test_gemm_ln_slots(int N);
/* Epilogue 6 on the first N columns of residual planes whose rows are `ldy` >= N elements apart (the last fc2 of the stack
 * updates the pose half of the channel-permuted residual stream only); the other columns stay. */
int mpl_test_gemm_emit_pitch(const void* A, const void* W, const float* bias, void* Y, int64_t M, int N, int K, void* stats_out,
                             void* x_lo, int ab_fp16, int ldy, int cta_group, mpl_stream_t stream);
/* The fused QKV projection + cross-view attention kernel (multiview_mpl.py:48-64 behind the folded norm1) in isolation:
 *   xb [M, D] bf16 raw residual rows, W [3D, D] / bias [3D] (or NULL) / gamma, beta [D] fp32 on the device, stats as above,
 *   att [M, D] bf16 out; M = poses * V rows, D = H * 136 (or H * 68, H even), 2 <= V <= 8.
 *   scratch: device buffer of at least align256(H*416*D*2) + 2 * align256(H*416*4) bytes for the packed operands. */
int mpl_test_qkv_attn(const void* xb, const float* W, const float* bias, const float* gamma, const float* beta,
                      const void* stats, int slots, float eps, float scale, void* att, int64_t M, int D, int H, int V,
                      void* scratch, size_t scratch_bytes, mpl_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* MPL_B200_H_ */
