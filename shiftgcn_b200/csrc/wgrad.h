// wgrad.h -- parameter block of sgcn_wgrad (plain C; mirrored by ctypes in shiftgcn_b200/_lib.py).
#pragma once
#ifdef __cplusplus
extern "C" {
#endif

typedef struct SgcnWgrad {
  const float* a_src;   /* SPATIAL: unit input x [rows, CA]            | TEMPORAL: dpre [rows, CA] (grad wrt conv output) */
  const float* a_tab0;  /* SPATIAL: tanh(mask)+1 [V, CA]                                                              */
  const float* b_src;   /* SPATIAL: grad wrt gcn output gh [rows, CB]  | TEMPORAL: tcn input h [rows, CB]              */
  const float* b_src2;  /* SPATIAL: pre-BN output z [rows, CB]                                                         */
  const float* b_tab0;  /* SPATIAL: alpha [V, CB]                      | TEMPORAL: BN scale [CB]                       */
  const float* b_tab1;  /* SPATIAL: beta  [V, CB]                      | TEMPORAL: BN shift [CB]                       */
  const float* b_tab2;  /* SPATIAL: gamma [V, CB]                      | TEMPORAL: effective ypos [CB]                 */
  float* dw;            /* [CA, CB] fp32, accumulated with atomics (caller zeroes it)                                  */
  long long groups;
  int V, G, T;
  int CA, CB;
} SgcnWgrad;

enum { SGCN_WG_SPATIAL = 0, SGCN_WG_TEMPORAL = 1 };

int sgcn_wgrad(const SgcnWgrad* params, int mode, void* stream);

#ifdef __cplusplus
}
#endif
