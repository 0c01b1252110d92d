/*
 * devias_b200 -- C-ABI of the B200-native DEVIAS hot path (sm_100a).
 *
 * The reference (KHU-VLL/DEVIAS) is pure PyTorch and has no FFI of its own; its "operator interface"
 * for this path is the set of torch library calls made by model/modeling_slot.py and agg_block/*.py
 * (SURVEY.md section 2.2).  Each entry point below replaces one of those call sites and cites it.
 * All pointers are DEVICE pointers owned by the caller (PyTorch); no entry point allocates, synchronises
 * or touches the host copy of the data.  `stream` is a cudaStream_t passed as void*.
 * Every function returns 0 on success, DEVIAS_ERR_* otherwise; devias_last_error() gives the detail.
 * There is no CPU fallback anywhere behind this ABI.
 */
#ifndef DEVIAS_B200_H_
#define DEVIAS_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DEVIAS_OK 0
#define DEVIAS_ERR_ARG 1
#define DEVIAS_ERR_CUDA 2
#define DEVIAS_ERR_UNSUPPORTED 3

/* library / device introspection */
int devias_abi_version(void);
const char* devias_last_error(void);
/* number of kernels this library has launched since load (monotonic; used for bench.py's gpu_launches) */
int64_t devias_launch_count(void);

/* ---- GEMM on tcgen05/TMEM fed by TMA ------------------------------------------------------------
 * D[m,n] = sum_k A[m,k] * B[n,k]  (bf16 operands, fp32 accumulation in tensor memory) + fused epilogue.
 * Replaces F.linear / nn.Linear / Conv3d-as-GEMM and their autograd dgrad/wgrad:
 *   model/modeling_slot.py:101 (qkv), :113 (proj), :61,:65 (fc1, fc2), :176 (patch embed).
 * Operand layouts: *_mn_major = 0: the operand is stored [rows = m or n][cols = k], k contiguous (ld = row stride);
 *                  *_mn_major = 1: stored [rows = k][cols = m or n] (a transposed view; used by wgrad/dgrad).
 * Epilogues (DEVIAS_EPI_*):
 *   STORE_BF16 : out_bf16[m,n]  = acc + bias[n]
 *   STORE_F32  : out_f32[m,n]   = acc + bias[n]
 *   GELU_BF16  : out_bf16 = acc + bias[n] (pre-activation), out2_bf16 = GELU_erf(acc + bias[n])      (modeling_slot.py:61-62)
 *   DGELU_BF16 : out_bf16 = acc * GELU'(aux_bf16[m,n])                                               (backward of :62)
 *   RESID_F32  : out_f32  = aux_f32[m % aux_row_mod (or m), n] + row_scale[m / rows_per_scale] * (acc + bias[n])
 *                (residual add :150-151 with per-sample drop-path scale; patch-embed + sin-cos table :354-355)
 *   ATOMIC_F32 : out_f32 += acc  via red.global.add (split-K weight gradients)
 * bias / row_scale may be NULL.  n % 32 == 0, k % 8 == 0, ld* % 8 == 0 required.
 */
#define DEVIAS_EPI_STORE_BF16 0
#define DEVIAS_EPI_STORE_F32 1
#define DEVIAS_EPI_GELU_BF16 2
#define DEVIAS_EPI_DGELU_BF16 3
#define DEVIAS_EPI_RESID_F32 4
#define DEVIAS_EPI_ATOMIC_F32 5
int devias_gemm_bf16(const void* a, int64_t lda, int a_mn_major, const void* b, int64_t ldb, int b_mn_major, int m, int n,
                     int k, int epilogue, void* out, int64_t ldo, void* out2, int64_t ldo2, const float* bias,
                     const void* aux, int64_t ldaux, int aux_row_mod, const float* row_scale, int rows_per_scale,
                     int split_k, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DEVIAS_B200_H_ */
