// Parameters of the streaming attention step kernels (attn_stream.cu), filled by st_attn_step_fwd/bwd.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

namespace st {

struct StreamParams {
  int P, A, E;
  const void *att1, *Fe;          // (rows*P, A), (rows*P, E) of T
  const float *att2, *wf;         // (rows, A), (A)
  // forward
  const float *bfp, *b_embed;
  float *alphas_w, *S, *ctx_out;  // alphas[b*alpha_stride + p] (written), S (rows,P) +=, ctx (rows, ld_ctx)
  __nv_bfloat16* ctx_bf16;        // optional bf16 copy of ctx (rows, ld_ctx_bf16)
  int alpha_stride, ld_ctx, ld_ctx_bf16;
  // backward
  const float *alphas_r, *dal, *dctx;
  int dal_stride, ld_dctx;
  float *de_out, *datt2;
  __nv_bfloat16* datt2_bf16;      // optional bf16 copy (rows, A)
  float* gt_out;                  // optional (rows, A): datt2 before the factor w_f
};

// Returns 1 if the streaming path handles this shape (and launched it), 0 if the caller must use
// the generic kernels of attn.cu, < 0 (st_status) on error.
int attn_stream_try(bool bwd, int rows, int in_bf16, int act, const StreamParams& p, cudaStream_t s);

}  // namespace st
