/* shiftgcn_b200.h -- C ABI of the B200-native Shift-GCN hot path (libshiftgcn_b200.so).
 *
 * This is the drop-in boundary for the reference's only native interface, the pybind11 module `shift_cuda`
 * (model/Temporal_shift/cuda/shift_cuda.cpp:44-47: forward / backward), widened to the fused spatial
 * (model/shift_gcn.py:77-142) and temporal (model/shift_gcn.py:48-74) units that the reference runs as chains of
 * torch library kernels.  Conventions:
 *   - extern "C", plain pointers and sizes, no C++ / torch types; every pointer is a DEVICE pointer unless noted;
 *   - every entry point returns 0 on success or a negative code; sgcn_last_error() gives the message (per host thread);
 *   - kernels are enqueued on the `stream` argument (a cudaStream_t), never synchronise, allocate nothing, and are
 *     CUDA-graph capturable.  The caller owns all buffers (the reference allocates outputs with at::zeros inside the
 *     extension, shift_cuda_kernel.cu:408,440,480-481; here nothing needs pre-zeroing except the fp64 reduction
 *     scratch, which every finalize call hands back zeroed);
 *   - activations of the fused entry points are channels-last rows [(n, t, v), C] in fp32, the reference's own
 *     internal layout (model/shift_gcn.py:123); the *_nchw entry points keep the reference extension's contiguous
 *     (N, C, H, W) layout and its float / double dispatch (shift_cuda_kernel.cu:413).
 * Built for sm_100a only.  The reference-side binding a maintainer would add is shown in INTEGRATION.md.
 */
#ifndef SHIFTGCN_B200_H_
#define SHIFTGCN_B200_H_

#ifdef __cplusplus
extern "C" {
#endif

/* ---------------------------------------------------------------- library ---------------------------------- */
int sgcn_abi_version(void);
const char* sgcn_last_error(void);
/* 0 when the current CUDA device is compute capability 10.x, error otherwise (there is no fallback path) */
int sgcn_device_check(void);
/* Traversal order of the full-tensor kernels (process-wide on/off setting, alternation state per device).  snake != 0: every kernel walks its
 * tiles in the opposite order of the kernel launched before it, so that it starts on the part of its inputs the
 * previous kernel touched last (still resident in the 126 MB L2); 0 (default): always ascending.  The call also
 * restarts the alternation: after 1 the next kernel descends, after 2 it ascends (callers restart it at the top of
 * every step so that all steps, and a captured CUDA graph, see the same orders).  Results do not depend on the
 * order.  Returns the previous on/off setting. */
int sgcn_set_traversal(int snake);
/* Cap on the grid of the persistent tile kernels (sgcn_rowgemm, sgcn_wgrad); 0 = one CTA per SM (default).  Used by the
 * parity tests: with a handful of CTAs a small tensor already gives every CTA many tiles, so the multi-tile steady
 * state of the pipelines (second TMEM accumulator, stage-ring wrap, mbarrier phase parities, cross-tile dW accumulation,
 * descending traversal) is compared against the oracle at sizes the oracle finishes in seconds.  Returns the previous cap. */
int sgcn_set_max_ctas(int n);
/* tcgen05 descriptor self test (tests only): mode 0: D[128,N] = A[128,K] * B[N,K]^T; mode 1: D[128,N] = A[128,M]^T * B[128,N] */
int sgcn_selftest_umma(const float* a, const float* b, float* d, int mode, int K, int N, int M, void* stream);
/* descriptor probe (tests only): D[128,N] = A*B, each operand K-major or MN-major with explicit swizzle / layout / LBO / SBO */
int sgcn_selftest_probe(const float* a, const float* b, float* d, int K, int N, int a_mn, int b_mn, int swz, int layout,
                        int lbo, int sbo, int kstep, void* stream);

/* ---------------------------------------------------------------- stand-alone temporal shift (NCHW) ------- */
/* Replaces shift_cuda.forward (shift_cuda.cpp:19-23 -> shift_cuda_kernel.cu:405-431, kernel K1 :12-76).
 * out[n,c,h,w] = bilinear sample of in[n,c] at (h*stride + ypos[c], w + xpos[c]), zero padded; out height = h/stride.
 * ypos already includes the +0.5 the Python op adds for stride != 1 (cuda/shift.py:14-19). */
int sgcn_shift_fwd_nchw_f32(const float* in, float* out, const float* xpos, const float* ypos, long long n, int c,
                            int h, int w, int stride, void* stream);
int sgcn_shift_fwd_nchw_f64(const double* in, double* out, const double* xpos, const double* ypos, long long n, int c,
                            int h, int w, int stride, void* stream);
/* Replaces shift_cuda.backward (shift_cuda.cpp:25-42 -> shift_cuda_kernel.cu:433-523): K2/K3 (grad_in), K4 + the
 * mean/sum reductions + K5 (grad_xpos = 0, grad_ypos = sign * 0.01 or 1e-4).  raw_pos (optional, [2][c]) receives the
 * reduced sums before K5; scratch is a zeroed double[2*c] that is handed back zeroed. */
int sgcn_shift_bwd_nchw_f32(const float* grad_out, const float* in, const float* xpos, const float* ypos,
                            float* grad_in, float* grad_xpos, float* grad_ypos, float* raw_pos, double* scratch,
                            long long n, int c, int h, int w, int stride, void* stream);
int sgcn_shift_bwd_nchw_f64(const double* grad_out, const double* in, const double* xpos, const double* ypos,
                            double* grad_in, double* grad_xpos, double* grad_ypos, double* raw_pos, double* scratch,
                            long long n, int c, int h, int w, int stride, void* stream);

/* ---------------------------------------------------------------- input streams of the ensemble ---------- */
/* Bone / motion / bone-motion streams derived from the joint batch on the device.  Replaces the numpy loops of
 * inference_pipeline.py:284-309 (derive_modalities), data_gen/gen_bone_data.py:44-58 and
 * data_gen/gen_motion_data.py:18-34 (bit-identical fp32 results):
 *   joint [N, C, T, V, M];  parent: int[V] (0-based parent joint, parent[v] == v for the root) or NULL = no bone step;
 *   motion != 0: out(t) = s(t+1) - s(t), last frame 0.
 * rows == 0: out keeps the [N, C, T, V, M] layout.  rows != 0: out is the channels-last row tensor [(N*M), T, V, C]
 * the units consume, and scale / shift (both [M*V*C], feature (m, v, c), or both NULL) apply the model's input
 * BatchNorm in inference form (model/shift_gcn.py:193-198): out = value * scale + shift. */
int sgcn_input_stream(const float* joint, float* out, const int* parent, const float* scale, const float* shift,
                      long long N, int C, int T, int V, int M, int motion, int rows, void* stream);

/* Sliding-window inference over one sequence (inference_pipeline.py:252-281 create_sliding_windows, :342-366
 * run_ensemble_inference, :377-386 aggregate_per_frame).  seq [C, Ttot, V, M] stays on the device; window w is the
 * frames start[w] .. start[w]+win-1, zero padded past Ttot; streams are derived from the PADDED window like
 * sgcn_input_stream (same parent / motion / rows / scale / shift arguments).  out: [W, C, win, V, M] or rows
 * [(W*M), win, V, C]. */
int sgcn_window_stream(const float* seq, float* out, const int* start, const int* parent, const float* scale,
                       const float* shift, long long W, int C, int Ttot, int win, int V, int M, int motion, int rows,
                       void* stream);
/* score[w] = softmax(logits[w, :])[cls] in fp64 (logits [W, num_class] fp32: the alpha-weighted ensemble sum);
 * per_frame[f] = mean of score[w] over the windows with start[w] <= f < start[w] + real[w], 0 where none covers f
 * (the reference divides by max(count, 1)).  score [W] and per_frame [total_frames] are fp64 device buffers.
 * logits == NULL: score already holds the window scores, only the aggregation runs. */
int sgcn_window_scores(const float* logits, const int* start, const int* real, double* score, double* per_frame, int W,
                       int num_class, int cls, int total_frames, void* stream);

/* Model head (model/shift_gcn.py:212-216: mean over (T, V) and the M persons, then fc).  pool_sums [N*M, C] fp64 holds
 * the per-(sample, channel) SUMS over the rows of the last unit's output -- sgcn_tshift_fwd (mode 1) accumulates them when
 * SgcnTShift::stats is set, and may then skip storing its output (out == NULL, inference).  sgcn_head_fwd:
 * pooled[n, c] = inv_count * sum_m pool_sums[(n, m), c]  (inv_count = 1 / (T*V*M)),  logits = pooled W^T + b  (W [K, C]);
 * pool_sums is handed back zeroed.  sgcn_head_bwd: dW = dlogits^T pooled, db = column sums of dlogits,
 * gpool[(n, m), c] = scale * sum_k dlogits[n, k] W[k, c]  -- the gradient of every row of the last unit's output
 * (scale = inv_count), spread over the rows by sgcn_bcast_rows. */
int sgcn_head_fwd(double* pool_sums, const float* W, const float* b, float* pooled, float* logits, int N, int M, int C,
                  int K, double inv_count, void* stream);
int sgcn_head_bwd(const float* dlogits, const float* pooled, const float* W, float* dW, float* db, float* gpool, int N,
                  int M, int C, int K, float scale, void* stream);

/* Input BatchNorm of the model in training mode (model/shift_gcn.py:196-198: BatchNorm1d(M*V*C) over (N, T)).
 * sgcn_data_bn_stats: stats[f][2] (fp64, zeroed by the caller / handed back zeroed by sgcn_bn_fwd_finalize) +=
 * {sum x, sum x^2} of feature f = (m*V + v)*C + c, x in the input layout [N, C, T, V, M] (V*M <= 128).  The scale / shift
 * tables of sgcn_bn_fwd_finalize are then applied by sgcn_input_stream together with the change to the row layout.
 * sgcn_data_bn_bwd: sums[f][2] += {sum g, sum g*xhat} with g the gradient wrt the BatchNorm output in the row layout
 * [(N*M), T, V, C] (V*C <= 128): beta.grad and gamma.grad after sgcn_reduce_export. */
int sgcn_data_bn_stats(const float* x, double* stats, long long N, int C, int T, int V, int M, void* stream);
int sgcn_data_bn_bwd(const float* g, const float* x, const float* mean, const float* invstd, double* sums, long long N,
                     int C, int T, int V, int M, void* stream);

/* Feeder augmentation on the device: feeders/tools.py:58-101 random_move for a batch [N, C, T, V, M] (C >= 2), in place.
 * node: int[K+1] frame indices 0 = node[0] < ... < node[K] = T (the reference's `node`, move_time = K); vals: fp64
 * [N, 4, K+1] = the angle (degrees), scale, x and y translation drawn at every node for every sample.  Between two nodes
 * the parameters follow np.linspace exactly; the x / y channels of frame t become R(a_t) * s_t * (x, y) + (tx_t, ty_t). */
int sgcn_random_move(float* data, const double* vals, const int* node, long long N, int C, int T, int V, int M, int K,
                     void* stream);

/* ---------------------------------------------------------------- first spatial unit (3 input channels) -- */
/* l1.gcn1 = Shift_gcn(3, 64) (model/shift_gcn.py:178, 121-142) including its `down` branch (1x1 conv + BatchNorm2d,
 * :82-86).  z and the conv output are recomputed from x wherever needed; only h, g cross HBM at full size.
 * x: rows [groups*V, 3]; h, g: rows [groups*V, 64].  Statistics buffers are zeroed fp64 scratch in the layouts the
 * finalize kernels below consume (pairs {sum, sum of squares} / {sum g, sum g*xhat}). */
typedef struct SgcnStem {
  const float* x;        /* [groups*V, 3]                                                                    */
  const float* maskmul;  /* tanh(Feature_Mask)+1 [V, 3]                                                       */
  const float* W;        /* Linear_weight [3, 64]                                                             */
  const float* bias;     /* Linear_bias [64] or NULL                                                          */
  const float* Wd;       /* down conv weight [64, 3]                                                          */
  const float* bd;       /* down conv bias [64] or NULL                                                       */
  /* forward, mode 0 (batch statistics) */
  double* stats_vd;      /* [V*64][2]  sums of z                                                              */
  double* stats_r;       /* [64][2]    sums of the conv output                                                */
  /* forward, mode 1 (apply) */
  const float* sc1;      /* BatchNorm1d scale / shift [V*64]                                                  */
  const float* sh1;
  const float* sc2;      /* BatchNorm2d (down) scale / shift [64]                                             */
  const float* sh2;
  float* h;              /* out: relu(BN1d(z) + BN2d(conv))                                                   */
  double* stats_h;       /* optional [64][2]: sums of h for the temporal unit's first BatchNorm               */
  /* backward (h above is an input here) */
  const float* g;        /* gradient wrt h (ReLU mask not applied)                                            */
  const float* mean1;    /* mode 0: batch mean / invstd of z [V*64] and of the conv output [64]               */
  const float* invstd1;
  const float* mean2;
  const float* invstd2;
  double* vd_sums;       /* mode 0 out: [V*64][2] {sum gm, sum gm*zhat}                                       */
  double* r_sums;        /* mode 0 out: [64][2]   {sum gm, sum gm*rhat}                                       */
  const float* al;       /* mode 1: dz = al*gm + be*z + ga [V*64]  (sgcn_bn1d_bwd_finalize)                   */
  const float* be;
  const float* ga;
  const float* a2;       /* mode 1: dr = a2*gm + b2*r + c2 [64]    (sgcn_bn1d_bwd_finalize with V = 1)        */
  const float* b2;
  const float* c2;
  double* dw_raw;        /* mode 1 out: [64][8] {dW[0..2][d], dWd[d][0..2], dbd[d], 0}                         */
  double* dmask_raw;     /* mode 1 out: [V*3] raw Feature_Mask gradient (sgcn_mask_grad_finalize)             */
  float* dx;             /* mode 1 out: [groups*V, 3]                                                         */
  long long groups;
  int V, D;
} SgcnStem;
int sgcn_stem_fwd(const SgcnStem* p, int mode, void* stream);
int sgcn_stem_bwd(const SgcnStem* p, int mode, void* stream);

/* ---------------------------------------------------------------- fused tensor-core contractions ---------- */
typedef struct SgcnRowGemm {
  const float* in0;    /* SPATIAL: x   | LERP: h      | PLAIN: rows       | DY: grad wrt gcn output (after ReLU mask) */
  const float* in1;    /* DY: z (pre-BN spatial output)                                                   */
  float* out;          /* [rows, N]                                                                       */
  const float* wimg;   /* canonical weight image from sgcn_prep_weight_image                              */
  const float* pro_a;  /* SPATIAL: tanh(mask)+1 [V,K] | LERP: BN scale [K] | DY: alpha [V,K]               */
  const float* pro_b;  /* LERP: BN shift [K]          | DY: beta  [V,K]                                    */
  const float* pro_c;  /* LERP: effective ypos [K]    | DY: gamma [V,K]                                    */
  const float* bias;   /* [N] or NULL                                                                     */
  const float* epi_a;  /* ROT_FUSED: BN scale [V,N]   | SPATIAL_BWD: tanh(mask)+1 [V,N]                    */
  const float* epi_b;  /* ROT_FUSED: BN shift [V,N]                                                        */
  const float* res;    /* ROT_FUSED: residual rows [rows,N] | SPATIAL_BWD: gradient added as-is (or NULL)  */
  const float* res2;   /* SPATIAL_BWD: block-residual gradient g_y (or NULL)                               */
  const float* res2m;  /* SPATIAL_BWD: block output y; g_y counts where y > 0 (NULL: g_y is already masked)    */
  const float* xin;    /* SPATIAL_BWD: unit input x (for the mask gradient)                                */
  double* stats;       /* ROT_RAW: per-(v,n) {sum, sum of squares}, accumulated                            */
  double* red0;        /* SPATIAL_BWD: raw mask gradient [V,N], accumulated                                */
  long long groups;    /* number of (n,t) groups = rows / V                                                */
  int V;               /* joints per group                                                                 */
  int G;               /* groups per tile, G*V <= 128                                                      */
  int T;               /* frames per sample (LERP bounds)                                                  */
  int K;               /* contraction width = input channels (64/128/256)                                  */
  int N;               /* output channels (64/128/256)                                                     */
  int relu;            /* ROT_FUSED / LINEAR: apply ReLU | SPATIAL_BWD: out *= [xin > 0] (pre-masked gradient
                          for the unit whose ReLU produced xin)                                           */
  /* PLAIN x LINEAR only (conv + BatchNorm side branches, model/shift_gcn.py:82-86, 31-45); all 0 = plain dense GEMM   */
  int k0;              /* > 0: the first k0 input channels come from in0 [rows, k0], the other K-k0 from in1      */
  int in0_gs, in1_gs;  /* frame stride of the input rows: group g reads group g*gs (strided 1x1 convolution)       */
  int out_gs;          /* frame stride of the output rows (transposed strided convolution)                         */
  int accum;           /* out += result instead of out = result                                                    */
  /* contraction precision: SGCN_PREC_TF32 (operands rounded to TF32, the default) or SGCN_PREC_FP32 (3xTF32: both
   * operands split into a TF32 head and a TF32 tail, D = Ah*Bh + Al*Bh + Ah*Bl with fp32 accumulation -- fp32-accurate,
   * ~3x the tensor-core work; wimg must then come from sgcn_prep_weight_image_split)                              */
  int prec;
} SgcnRowGemm;

enum { SGCN_PREC_TF32 = 0, SGCN_PREC_FP32 = 1 };

enum { SGCN_PRO_SPATIAL = 0, SGCN_PRO_LERP = 1, SGCN_PRO_PLAIN = 2, SGCN_PRO_DY = 3 };
enum { SGCN_EPI_ROT_RAW = 0, SGCN_EPI_ROT_FUSED = 1, SGCN_EPI_LINEAR = 2, SGCN_EPI_SPATIAL_BWD = 3 };

int sgcn_rowgemm(const SgcnRowGemm* params, int prologue, int epilogue, void* stream);

typedef struct SgcnWgrad {
  const float* a_src;   /* SPATIAL: unit input x [rows, CA]            | TEMPORAL: dpre [rows, CA] (grad wrt conv output) */
  const float* a_tab0;  /* SPATIAL: joint-rotated mask multiplier [V, CA] (maskmul_rot of sgcn_mask_prepare_rot)            */
  const float* b_src;   /* SPATIAL: grad wrt gcn output gh [rows, CB]  | TEMPORAL: tcn input h [rows, CB]              */
  const float* b_src2;  /* SPATIAL: pre-BN output z [rows, CB]                                                         */
  const float* b_tab0;  /* SPATIAL: alpha [V, CB]                      | TEMPORAL: BN scale [CB]                       */
  const float* b_tab1;  /* SPATIAL: beta  [V, CB]                      | TEMPORAL: BN shift [CB]                       */
  const float* b_tab2;  /* SPATIAL: gamma [V, CB]                      | TEMPORAL: effective ypos [CB]                 */
  float* dw;            /* [CA, CB] fp32, accumulated with atomics (caller zeroes it)                                  */
  long long groups;
  int V, G, T;
  int CA, CB;
  int a_gs;             /* PLAIN: row group g of A is group g*a_gs of a_src (strided 1x1 convolution), 0 or 1 = dense  */
  int b_gs;             /* PLAIN: the same for B / b_src                                                           */
  int prec;             /* SGCN_PREC_TF32 / SGCN_PREC_FP32 (3xTF32 split of BOTH activation operands)               */
} SgcnWgrad;

/* PLAIN: dW[a, b] = sum_rows a_src[row, a] * b_src[row, b] with both operands copied as they are (Gram matrices and
 * input-gradient correlations of the conv + BatchNorm side branches, model/shift_gcn.py:82-86, 31-45) */
enum { SGCN_WG_SPATIAL = 0, SGCN_WG_TEMPORAL = 1, SGCN_WG_PLAIN = 2 };

int sgcn_wgrad(const SgcnWgrad* params, int mode, void* stream);

/* ---------------------------------------------------------------- bandwidth-bound SIMT kernels ------------ */
/* forward temporal shift with stride: s = Shift(q); mode 0 -> stats[c] += {sum, sumsq}; mode 1 -> out = [relu](s*scale+shift+res) */
typedef struct SgcnTShift {
  const float* q;         /* [n, T_in, V, C]                       */
  const float* res;       /* [n, T_out, V, C] or NULL (mode 1)     */
  float* out;             /* [n, T_out, V, C] (mode 1)             */
  const float* ypos_eff;  /* [C] ypos (+0.5 when stride != 1)      */
  const float* scale;     /* [C] folded BN scale (mode 1)          */
  const float* shift;     /* [C] folded BN shift (mode 1)          */
  double* stats;          /* [C][2] (mode 0)                       */
  long long n_samples;
  int T_in, T_out, V, C, stride, relu;
} SgcnTShift;

/* backward of  out = [relu](BN(Shift(q)) + res):
 * mode 0: sums[c][5] += { g, g*shat, g*dq, dq, shat*dq };  mode 1: dpre = [q>0] * Shift^T(k1*(g - m1 - shat*m2)), dbias[c] += dpre */
typedef struct SgcnTShiftBwd {
  const float* q;
  const float* gy;        /* grad wrt out [n, T_out, V, C]                      */
  const float* y;         /* out (ReLU mask source), needed when relu != 0      */
  const float* ypos_eff;
  const float* mean;      /* [C] BN mean used in the forward                    */
  const float* invstd;    /* [C]                                                */
  const float* k1;        /* [C] gamma*invstd        (mode 1)                   */
  const float* m1;        /* [C] sum(g)/count  or 0  (mode 1)                   */
  const float* m2;        /* [C] sum(g*shat)/count or 0 (mode 1)                */
  double* sums;           /* [C][5] (mode 0)                                    */
  float* dpre;            /* [n, T_in, V, C] (mode 1)                           */
  double* dbias;          /* [C] (mode 1)                                       */
  long long n_samples;
  int T_in, T_out, V, C, stride, relu;
} SgcnTShiftBwd;

/* backward of  p = Shift_1(BN(h)):  du = Shift^T(dp)
 * mode 0: sums[c][3] += { du, du*hhat, dp*dU }   (skipped when gate != NULL and *gate == 0, see sgcn_tshift_in_bwd_sums);
 * mode 1: gh = [h>0] * k1*(du - m1 - hhat*m2);  vd_sums[v,c][2] += { gh, gh*zhat } when z != NULL;
 *         pos_sums[c] += dp*dU (the position-gradient sum of mode 0; needs scale / shift) */
typedef struct SgcnTShiftInBwd {
  const float* dp;        /* grad wrt p [n, T, V, C]                            */
  const float* h;         /* tcn input (gcn output)                             */
  const float* z;         /* pre-BN spatial output or NULL                      */
  const float* ypos_eff;
  const float* mean;      /* [C] BN(h) statistics                               */
  const float* invstd;
  const float* scale;     /* [C] folded BN(h) scale / shift (mode 0)            */
  const float* shift;
  const float* k1;
  const float* m1;
  const float* m2;
  const float* zmean;     /* [V, C] BN1d statistics of z (mode 1, z != NULL)    */
  const float* zinvstd;
  double* sums;           /* [C][3] (mode 0)                                    */
  double* vd_sums;        /* [V, C][2] (mode 1)                                 */
  float* gh;              /* [n, T, V, C] (mode 1)                              */
  double* pos_sums;       /* [C] (mode 1)                                       */
  const int* gate;        /* device flag or NULL (mode 0)                       */
  long long n_samples;
  int T, V, C, relu_h;
} SgcnTShiftInBwd;

/* The two BatchNorm backward sums of sgcn_tshift_in_bwd mode 0 WITHOUT a pass over dp and h.  With p = Shift_1(U),
 * U = scale*h + shift, dp = dpre * W_t and dW_t = dpre^T p (model/shift_gcn.py:66-70):
 *     sum du        = sum_d W_t[d,c] * dbt[d]  -  sum over the few frames r next to the sequence ends of dp(r)*(1 - e_c(r))
 *     sum du * U    = sum_d W_t[d,c] * dW_t[d,c]
 *     sum du * hhat = invstd * ((sum du*U - shift*sum du) / scale - mean * sum du)
 * (dbt = column sums of dpre = the conv-bias gradient, e_c(r) = total weight of the shift taps of frame r that fall
 * inside the sequence).  Only the boundary frames of dp are read.  sums[c][0..1] receive the two sums.  gate[0] is
 * set to 1 -- and nothing is written to sums -- when some |gamma[c]| = |scale/invstd| < 1e-3 makes the division
 * ill-conditioned: sgcn_tshift_in_bwd mode 0 launched with the same gate then computes the sums the long way. */
typedef struct SgcnTShiftInSums {
  const float* dp;        /* [n, T, V, C]                                       */
  const float* ypos_eff;  /* [C]                                                */
  const float* Wt;        /* [C out][C in] temporal_linear.weight               */
  const float* dWt;       /* its gradient                                       */
  const float* dbt;       /* [C] conv-bias gradient                             */
  const float* mean;      /* [C] BN(h) tables                                   */
  const float* invstd;
  const float* scale;
  const float* shift;
  double* sums;           /* [C][3]                                             */
  int* gate;              /* device flag                                        */
  long long n_samples;
  int T, V, C;
} SgcnTShiftInSums;
int sgcn_tshift_in_bwd_sums(const SgcnTShiftInSums* p, void* stream);
/* K5 (shift_cuda_kernel.cu:371-395) on pos_sums[c] / n_batch: grad_xpos = 0, grad_ypos = sign * 0.01 (1e-4 at 0);
 * raw_out (optional) receives the means; pos_sums is handed back zeroed. */
int sgcn_shift_pos_finalize(double* pos_sums, float* grad_xpos, float* grad_ypos, float* raw_out, int C, double n_batch,
                            void* stream);

int sgcn_bn_res_relu_fwd(const float* z, const float* res, float* h, const float* scale, const float* shift,
                         double* stats_out, long long rows, int V, int D, int relu, void* stream);
int sgcn_tshift_fwd(const SgcnTShift* p, int mode, void* stream);
int sgcn_tshift_bwd(const SgcnTShiftBwd* p, int mode, void* stream);
int sgcn_tshift_in_bwd(const SgcnTShiftInBwd* p, int mode, void* stream);

/* ---------------------------------------------------------------- conv + BatchNorm side branches ---------- */
/* Parameter-sized fp64 arithmetic of BatchNorm2d(Conv2d 1x1(x)) (`down`, model/shift_gcn.py:82-86; strided `tcn`
 * residual, :31-45,157-158).  The full-size work is four tensor-core contractions (sgcn_wgrad PLAIN, sgcn_rowgemm
 * PLAIN x LINEAR); these entry points turn their small results into the folded weights and the gradients. */
typedef struct SgcnSideFold {
  double* sx_sums;            /* training: [C][2] channel sums of x from sgcn_channel_stats; handed back zeroed          */
  const float* XX;            /* training: [C, C] Gram matrix x^T x                                                   */
  const float* Wd;            /* [D, C] conv weight                                                                   */
  const float* bd;            /* [D] conv bias or NULL                                                                */
  const float* gamma;         /* [D] BatchNorm weight / bias                                                          */
  const float* beta;
  float* running_mean;        /* [D] updated when training (may be NULL then), read otherwise                         */
  float* running_var;
  long long* num_batches_tracked; /* or NULL                                                                          */
  float* Wf;                  /* out [D, C]: gamma*invstd*Wd                                                          */
  float* bf;                  /* out [D]:    beta + gamma*invstd*(bd - mean_r)                                        */
  double* mean_r;             /* out [D]: batch (or running) mean / invstd of the conv output                         */
  double* invstd;
  double* sx;                 /* out [C]: channel sums of x (training)                                                */
  int* counter;               /* zeroed int scratch (training); handed back zeroed                                    */
  double rows, eps, momentum; /* momentum < 0: cumulative average 1 / num_batches_tracked (caller already incremented it) */
  int C, D, training;
} SgcnSideFold;
int sgcn_side_fold(const SgcnSideFold* p, void* stream);

typedef struct SgcnSideBwd {
  const float* P;             /* [C, D] correlation x^T G                                                             */
  const float* sg;            /* [D] column sums of G                                                                 */
  const float* XX;            /* training: [C, C]                                                                     */
  const double* sx;           /* training: [C]                                                                        */
  const float* Wd;            /* [D, C]                                                                               */
  const float* bd;            /* [D] or NULL                                                                          */
  const float* gamma;         /* [D]                                                                                  */
  const double* invstd;       /* [D] from sgcn_side_fold                                                              */
  const double* mean_r;
  float* dgamma;              /* out [D]                                                                              */
  float* dbeta;               /* out [D]                                                                              */
  float* dWd;                 /* out [D, C]                                                                           */
  float* dbd;                 /* out [D]                                                                              */
  float* Wcat;                /* out [D + C, C]: dx = [G | x] Wcat + kvec                                             */
  float* kvec;                /* out [C]                                                                              */
  double* coef;               /* scratch [2 * D]                                                                      */
  double rows;
  int C, D, training;
} SgcnSideBwd;
int sgcn_side_bwd(const SgcnSideBwd* p, void* stream);

/* out[n, r, c] = g[n, c] * scale for every row r < rows_per_n: the gradient of the global mean over (T, V)
 * (model/shift_gcn.py:212-214) written directly in the row layout.  mask_y (optional, [n, rows_per_n, C]): the pooled rows
 * themselves, the output of a unit that ends in a ReLU (:162); out is then multiplied by [mask_y > 0], i.e. it is the
 * gradient in front of that ReLU, and the unit's backward kernels skip their own reads of the mask source. */
int sgcn_bcast_rows(const float* g, float* out, const float* mask_y, long long n, long long rows_per_n, int C, float scale,
                    void* stream);

/* stats[c][2] += {sum, sumsq} over rows */
int sgcn_channel_stats(const float* x, double* stats, long long rows, int C, void* stream);
/* the same over `groups` row groups of V rows, group g being group g*gs of x (the frames a strided 1x1 conv reads) */
int sgcn_channel_stats_groups(const float* x, double* stats, long long groups, int V, int C, int gs, void* stream);
/* gh = g*[h>0]; vd_sums[v,c][2] += {gh, gh*zhat} */
int sgcn_relu_bn1d_bwd_stats(const float* g, const float* h, const float* z, const float* zmean, const float* zinvstd,
                             float* gh, double* vd_sums, long long groups, int V, int C, void* stream);
/* out = g*[y>0] */
int sgcn_relu_mask_grad(const float* g, const float* y, float* out, long long numel, void* stream);

/* ---- small per-feature kernels (prep.cu) ---- */

/* batch-norm forward finalize: {sum, sumsq} -> mean/invstd/scale/shift, running-stat update (momentum, unbiased var),
 * num_batches_tracked += 1, stats cleared.  training == 0: scale/shift from the running statistics, nothing updated.
 * momentum < 0 selects nn.BatchNorm(momentum=None): the caller has ALREADY incremented num_batches_tracked and the
 * update factor is 1 / num_batches_tracked (cumulative moving average). */
int sgcn_bn_fwd_finalize(double* stats, const float* gamma, const float* beta, float* running_mean, float* running_var,
                         long long* num_batches_tracked, float* mean, float* invstd, float* scale, float* shift,
                         int features, double count, double momentum, double eps, int training, void* stream);

/* backward finalize of the output shift + BN: from sums[c][5] -> dgamma, dbeta, k1, m1, m2 and the K5-constrained
 * position gradient (raw = k1*(S2 - m1*S3 - m2*S4) / n_batch).  raw_out (optional) receives the raw means. */
int sgcn_tshift_bwd_finalize(double* sums, const float* gamma, const float* invstd, float* dgamma, float* dbeta,
                             float* k1, float* m1, float* m2, float* grad_xpos, float* grad_ypos, float* raw_out,
                             int C, double count, double n_batch, int training, void* stream);

/* backward finalize of BN + input shift: sums[c][3] -> dgamma, dbeta, k1, m1, m2, K5-constrained position gradient */
int sgcn_tshift_in_bwd_finalize(double* sums, const float* gamma, const float* invstd, float* dgamma, float* dbeta,
                                float* k1, float* m1, float* m2, float* grad_xpos, float* grad_ypos, float* raw_out,
                                int C, double count, double n_batch, int training, void* stream);

/* backward finalize of the BN1d over (v,d): vd_sums[f][2] -> dgamma, dbeta, alpha/beta/gamma tables of
 * dz = alpha*gh + beta*z + gamma, and the Linear_bias gradient dbias[d] = sum_v k*(S_g - count*m1). */
int sgcn_bn1d_bwd_finalize(double* vd_sums, const float* gamma, const float* mean, const float* invstd, float* dgamma,
                           float* dbeta, float* alpha, float* beta, float* gam, float* dbias, int V, int D,
                           double count, int training, void* stream);

/* maskmul = tanh(mask) + 1 */
int sgcn_mask_prepare(const float* mask, float* maskmul, int n, void* stream);
/* same, plus the table seen from the SOURCE joint of the shift_in gather (model/shift_gcn.py:127-129):
 * maskmul_rot[v, c] = maskmul[(v - c) mod V, c] -- what kernels that load x[., v, c] rows and scatter them need */
int sgcn_mask_prepare_rot(const float* mask, float* maskmul, float* maskmul_rot, int V, int C, void* stream);
/* dmask = raw * (1 - tanh(mask)^2); raw cleared */
int sgcn_mask_grad_finalize(double* raw, const float* mask, float* dmask, int n, void* stream);

/* canonical (SWIZZLE_128B, TF32-rounded) image of B[n][k] = src[n*ld_n + k*ld_k], chunked by 64 k */
int sgcn_prep_weight_image(const float* src, long long ld_n, long long ld_k, int N, int K, float* image, void* stream);
/* fp32-accurate mode: two canonical images back to back, image[0 .. N*K) = TF32 head of the weights, image[N*K .. 2*N*K) =
 * TF32 tail (weight - head); consumed by sgcn_rowgemm with prec = SGCN_PREC_FP32 */
int sgcn_prep_weight_image_split(const float* src, long long ld_n, long long ld_k, int N, int K, float* image,
                                 void* stream);

/* double -> float copy of a reduction buffer (+ clear) */
int sgcn_reduce_export(double* src, float* dst, int n, double scale, void* stream);

/* ---------------------------------------------------------------- optimizer step (optim.cu) ----------------- */
/* The step after backward (main.py:301-322 parameter groups, :412-414 optimizer.step, App. E-7 of SURVEY.md), as ONE
 * kernel over flat fp32 buffers of n_param elements:
 *     g   = grad[i] * hyper[2]                                   (1/world after a summed all-reduce, else 1)
 *     g   = K5(grad[n_param + ypos_src[i]])  where ypos_src[i] >= 0  (shift_cuda_kernel.cu:371-395 on the REDUCED raw
 *           sums that ride behind the gradients: sign * 0.01, or 1e-4 when the sum is exactly 0);  ypos_src may be NULL
 *     d   = g + weight_decay[i] * param[i];   buf = hyper[1] * buf + d   (zero-initialised momentum buffer)
 *     param[i] -= hyper[0] * (nesterov ? d + hyper[1] * buf : buf);      grad[i] = g
 * hyper = {lr, momentum, gradient scale} lives in DEVICE memory: a captured CUDA graph follows the learning-rate
 * schedule (main.py:342-351) without re-capture. */
int sgcn_sgd_epilogue(float* param, float* grad, float* momentum_buf, const float* weight_decay, const int* ypos_src,
                      const float* hyper, long long n_param, int nesterov, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SHIFTGCN_B200_H_ */
