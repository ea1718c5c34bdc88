/**
 * \file model_kernels.h
 * \brief Launch API of the fused Llama-block elementwise kernels (see model_kernels.cu).
 */
#ifndef PS_KERNELS_MODEL_KERNELS_H_
#define PS_KERNELS_MODEL_KERNELS_H_
#include <cstddef>
#include "kernels/ps_kernels.h"

#ifdef __cplusplus
extern "C" {
#endif
/*! \brief qkv[T,(H+2KV)*D] -> rotated q[T,H,D], k[T,KV,D] and dense v[T,KV,D]; cos/sin fp32 [seq, D/2] */
int ps_launch_rope_split(const void* qkv, void* q, void* k, void* v, const float* cos_t,
                         const float* sin_t,
                         size_t tokens, int seq, int n_heads, int n_kv, int hd, ps_stream_t stream);
/*! \brief backward: (dq, dk, dv) dense -> dqkv rows, inverse rotation on the q and k parts */
int ps_launch_rope_merge_bwd(const void* dq, const void* dk, const void* dv, void* dqkv,
                             const float* cos_t, const float* sin_t, size_t tokens, int seq,
                             int n_heads, int n_kv, int hd, ps_stream_t stream);
/*! \brief out[T,F] = silu(gu[:, :F]) * gu[:, F:] */
int ps_launch_swiglu_fwd(const void* gu, void* out, size_t tokens, int f, ps_stream_t stream);
int ps_launch_swiglu_bwd(const void* gu, const void* dout, void* dgu, size_t tokens, int f,
                         ps_stream_t stream);
#ifdef __cplusplus
}
#endif
#endif  // PS_KERNELS_MODEL_KERNELS_H_
