// rowgemm.h -- parameter block of sgcn_rowgemm (plain C; mirrored by ctypes in shiftgcn_b200/_lib.py
// and documented in include/shiftgcn_b200.h).  All pointers are device pointers.
#pragma once
#ifdef __cplusplus
extern "C" {
#endif

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
  const float* res2m;  /* SPATIAL_BWD: block output y; g_y counts where y > 0                              */
  const float* xin;    /* SPATIAL_BWD: unit input x (for the mask gradient)                                */
  double* stats;       /* ROT_RAW: per-(v,n) {sum, sum of squares}, accumulated                            */
  double* red0;        /* SPATIAL_BWD: raw mask gradient [V,N], accumulated                                */
  long long groups;    /* number of (n,t) groups = rows / V                                                */
  int V;               /* joints per group                                                                 */
  int G;               /* groups per tile, G*V <= 128                                                      */
  int T;               /* frames per sample (LERP bounds)                                                  */
  int K;               /* contraction width = input channels (64/128/256)                                  */
  int N;               /* output channels (64/128/256)                                                     */
  int relu;            /* ROT_FUSED / LINEAR: apply ReLU                                                   */
} SgcnRowGemm;

enum { SGCN_PRO_SPATIAL = 0, SGCN_PRO_LERP = 1, SGCN_PRO_PLAIN = 2, SGCN_PRO_DY = 3 };
enum { SGCN_EPI_ROT_RAW = 0, SGCN_EPI_ROT_FUSED = 1, SGCN_EPI_LINEAR = 2, SGCN_EPI_SPATIAL_BWD = 3 };

int sgcn_rowgemm(const SgcnRowGemm* params, int prologue, int epilogue, void* stream);

#ifdef __cplusplus
}
#endif
