/*
 * showtell_b200.h -- C ABI of libshowtell_b200.so: the B200 (sm_100a) kernels behind the
 * show-tell caption-decoder classes.
 *
 * The reference (guptakhil/show-tell) has NO plugin / FFI interface: its decoder is a set of plain
 * Python torch.nn.Module classes (rnn.py:10 RNN, LSTM/rnn_lstm.py:8 RNN, Attention/rnn_attn.py:33
 * RNN_Attn, Attention/rnn_attn_LSTM.py:33 RNN_Attn) whose arithmetic is done by PyTorch library
 * calls.  Each entry point below therefore cites the reference call site whose library call it
 * replaces; the Python classes in showtell_b200/ keep the reference's names/signatures and bind
 * these symbols with ctypes.
 *
 * Conventions
 *   - Plain pointers and sizes only.  Every pointer is a DEVICE pointer unless the name ends in
 *     _host.  The library never allocates tensors; the caller owns inputs, outputs and workspaces.
 *   - All matrices are dense row-major; "ld" = elements between consecutive rows.
 *   - Every call is asynchronous on `stream` (a cudaStream_t passed as void*), re-entrant, and
 *     returns ST_OK or a negative st_status.  No exceptions, no aborts.  st_last_error() returns
 *     a thread-local message for the last failure on the calling thread.
 *   - "packed time-major" = rows of step 0 (batch_sizes[0] of them), then step 1, ... exactly the
 *     PackedSequence.data order of torch.nn.utils.rnn.pack_padded_sequence (rnn.py:31).
 *   - kind: ST_GRU gate order [r|z|n] (nn.GRU), ST_LSTM gate order [i|f|g|o] (nn.LSTM).
 *   - There is no CPU fallback anywhere in this library.
 */
#ifndef SHOWTELL_B200_H_
#define SHOWTELL_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* st_stream_t; /* cudaStream_t */

typedef enum {
  ST_OK = 0,
  ST_ERR_BAD_SHAPE = -1,    /* a dimension is <= 0, too large, or violates an alignment rule   */
  ST_ERR_UNSORTED = -2,     /* batch_sizes not non-increasing (lengths not sorted descending)  */
  ST_ERR_UNSUPPORTED = -3,  /* kind / dtype / option not supported by this build               */
  ST_ERR_CUDA = -4,         /* a CUDA runtime / driver call failed; see st_last_error()        */
  ST_ERR_NULL = -5,         /* a required pointer is NULL                                      */
  ST_ERR_WORKSPACE = -6     /* workspace too small                                             */
} st_status;

enum { ST_GRU = 0, ST_LSTM = 1 };
enum { ST_MAX_STEPS = 128 }; /* longest padded caption the sequence kernels accept */

int st_version(void);
const char* st_last_error(void);
/* ---- collate on the device (utils.py:61-77 `create_batch`): samples sorted by caption length, longest first, ties in
 * their original order (Python's stable sort).  lengths (B) int64 on the device -> perm (B): sorted position r holds
 * original sample perm[r]; sorted_len (B); batch_sizes (T) int32 = #{i : len_i > t} (what rnn.py:31's
 * pack_padded_sequence derives).  T may be 0 (no batch_sizes). */
int st_collate_sort(const int64_t* lengths, int B, int T, int64_t* perm, int64_t* sorted_len, int32_t* batch_sizes,
                    st_stream_t stream);
/* dst row r = src row perm[r] for rows of row_bytes bytes (features of any dtype / shape; rows <= 65535). */
int st_gather_rows_bytes(void* dst, const void* src, const int64_t* perm, int rows, int64_t row_bytes, st_stream_t stream);
/* The sorted caption matrix, re-padded with zeros from each caption's length on (utils.py:72-75):
 * dst (B, T_out) <- src (B, T_in) rows perm[r]. */
int st_collate_captions(int64_t* dst, const int64_t* src, const int64_t* perm, const int64_t* sorted_len, int B, int T_in,
                        int T_out, st_stream_t stream);
/* Number of kernels this library has launched in this process (bench.py's gpu_launches). */
int64_t st_launch_count(void);
/* Development / test aid: programmatic dependent launch of the per-step kernels on (default) / off. */
int st_debug_set_pdl(int on);
/* A/B timing aid: k-blocks per TMA operation of the BPTT kernel's operand ring (0 = the library's choice, 1 / 2 / 4). */
int st_debug_set_bwd_kp(int kp);
/* ... and its K split: low byte 0 = choose, 1 = every CTA streams the whole K extent, 4 = K split over clusters of 4
 * unit tiles; bits 8.. = batch-tile height of the K-split kernel (0 = choose, 64, 128). */
int st_debug_set_bwd_ks(int ks);
/* 1 = launch the multi-step tensor-core BPTT kernel cooperatively (needed only if two persistent recurrent kernels may
 * run on one device at the same time); default 0 = ordinary launch (no wait for a drained GPU). */
int st_debug_set_coop(int on);
/* CTAs (= SMs) the BPTT kernel st_rnn_seq_tc_bwd occupies for a batch of B rows: the host side gives the rest of the
 * GPU to the GEMMs it runs beside it (st_gemm_set_sm_limit). */
int st_rnn_seq_tc_bwd_ctas(int kind, int H, int B);
/* Cap on the persistent grid of the st_gemm_bf16* / st_vocab_ce_* launches that follow (0 = all SMs): a GEMM launched
 * beside a cooperative recurrent kernel must leave that kernel's SMs free, or the two serialise. */
int st_gemm_set_sm_limit(int n);
/* 1: the fp32 outputs of the st_gemm_bf16 calls that follow have already been cleared by the caller (a stream-K launch
 * reduces partial tiles into C and otherwise clears it itself, a memset in front of the kernel); 0 = default.  Lets a
 * step clear such a buffer early, off its critical path. */
int st_gemm_set_c_zeroed(int on);
/* Device facts the host side sizes grids with. */
int st_device_info(int* sm_count, int* cc_major, int* cc_minor, int64_t* smem_optin_bytes);

/* ------------------------------------------------------------------------------------------
 * fp32 GEMM (CUDA cores, FFMA): C[M,N] = alpha * op(A) * op(B) + beta * C + bias[N]
 *   transA = 0: A is (M,K) lda;  1: A is (K,M) lda.   transB = 0: B is (K,N) ldb;  1: B is (N,K) ldb.
 * Replaces every nn.Linear / GRU / LSTM input-side matmul of the fp32 reference:
 * rnn.py:32-33 (W_ih hoist, vocabulary projection), rnn_attn.py:23,24,62,70 (encoder_att,
 * decoder_att, init_h, embed) and their autograd transposes (dX = dY W, dW = dY^T X).
 * bias may be NULL.  beta = 0 never reads C.
 * ------------------------------------------------------------------------------------------ */
int st_sgemm(int transA, int transB, int M, int N, int K, float alpha, const float* A, int lda,
             const float* B, int ldb, float beta, float* C, int ldc, const float* bias,
             st_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * bf16 GEMM on the tensor cores (tcgen05.mma, TMEM accumulators, TMA operand loads):
 *   C[M,N] = alpha * A[M,K] . B[N,K]^T + bias[N] + beta * C   A, B bf16 row-major, K contiguous
 *   (beta != 0 accumulates into an fp32 C)
 * C is fp32 (c_is_bf16 = 0) or bf16.  The bf16-mode replacement of the same nn.Linear call sites
 * as st_sgemm.  A, B 16-byte aligned, lda/ldb multiples of 8; M, N, K arbitrary (TMA zero-fills
 * the tails).
 * ------------------------------------------------------------------------------------------ */
int st_gemm_bf16(int M, int N, int K, const void* A, int lda, const void* B, int ldb, void* C, int ldc,
                 int c_is_bf16, const float* bias, float alpha, float beta, st_stream_t stream);
/* Same, with operand-major flags: a_mn != 0 means A is passed as its transpose, the (K, M) row-major matrix
 * (lda >= M); b_mn likewise B as (K, N) (ldb >= N).  The kernel consumes those layouts in place (MN-major UMMA
 * operands), so the transposed products of the backward pass -- dW = dY^T X (a_mn = b_mn = 1: autograd of
 * nn.Linear, rnn.py:33) and dX = dY W (b_mn = 1) -- need no transposed copies. */
int st_gemm_bf16_ex(int M, int N, int K, const void* A, int lda, int a_mn, const void* B, int ldb, int b_mn, void* C,
                    int ldc, int c_is_bf16, const float* bias, float alpha, float beta, st_stream_t stream);

/* fp32 (rows, cols) -> bf16 copy `dst` (rows, cols) and/or bf16 transpose `dstT` (cols, rows);
 * either may be NULL.  Produces the K-major operands st_gemm_bf16 needs for dX = dY W and
 * dW = dY^T X. */
int st_cast_bf16(const float* src, int rows, int cols, int lds, void* dst, int ldd, void* dstT, int lddT,
                 st_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Vocabulary projection fused with cross-entropy (rnn.py:33 + main.py:94,149), bf16 mode.
 * Forward: logits = Hs Wv^T + bv are formed tile by tile in TMEM and reduced on the fly to a per-row
 * (max, sum-exp) partial per 128-column tile; the (N,V) logits never reach HBM.
 *   Hs (M, H) bf16, Wv (V, H) bf16, bv (V) fp32, target (M) int64
 *   part_max / part_sum (M, st_vocab_ce_parts(V)) scratch; tlogit (M) scratch
 *   lse (M) out; loss_sum (1) out = sum_m (lse_m - logit_m[target_m])
 * Backward: recomputes each logits tile and writes dlogits = (softmax - onehot) * scale as bf16,
 * row-major P (M, ldp) and transposed PT (V, ldpt) (PT may be NULL): the operands of
 * dHs = P Wv and dWv = P^T Hs.
 * ------------------------------------------------------------------------------------------ */
int st_vocab_ce_parts(int V);
/* fp32-accurate GEMM on the tensor cores (3xTF32): C[M,N] = alpha * A . B^T + bias (+ beta * C), where each fp32
 * operand is given as hi + lo from st_split_tf32 (hi = upper 11 mantissa bits, lo = the exact remainder) and the
 * product is hi.hi + hi.lo + lo.hi with fp32 accumulation.  Replaces the fp32 nn.Linear products of the decoding
 * loops (rnn.py:50,88; beam_search generate_function), whose token ids are defined against fp32 arithmetic.
 * Operands 16-byte aligned, lda / ldb multiples of 4. */
int st_split_tf32(const float* src, int rows, int cols, int lds, float* hi, float* lo, int ldd, st_stream_t stream);
int st_gemm_tf32x3(int M, int N, int K, const float* A_hi, const float* A_lo, int lda, const float* B_hi,
                   const float* B_lo, int ldb, float* C, int ldc, const float* bias, float alpha, float beta,
                   st_stream_t stream);
/* The same fp32-accurate product fused with a per-row top-K of A.B^T + bias (rnn.py:51 `max(1)[1]`, rnn.py:63,90-91
 * `topk(k)`; beam_search.py:84 `argsort`): the (M, N) logits are never written.  Each epilogue warp keeps the best `topk`
 * (<= 8) of its columns per row (cand_val / cand_idx: scratch of M * st_topk_parts(N) entries each), a merge kernel picks
 * the `topk` best per row: value descending, the LOWER column index first among equal values.
 * Outputs (any may be NULL): val / idx (M, out_stride), tok[m * tok_stride] = best index as int64.
 * row_max / row_sum (both or neither; then part_stats = scratch of M * st_topk_parts(N) / 4 floats): the soft-max
 * normaliser of each row, row_sum = sum_j exp(x_j - row_max)  (beam_search.py:85-88 costs -log p). */
int st_topk_parts(int N);
int st_gemm_tf32x3_topk(int M, int N, int K, const float* A_hi, const float* A_lo, int lda, const float* B_hi,
                        const float* B_lo, int ldb, const float* bias, int topk, float* cand_val, int32_t* cand_idx,
                        float* val, int32_t* idx, int out_stride, int64_t* tok, int tok_stride, float* part_stats,
                        float* row_max, float* row_sum, st_stream_t stream);
/* Screening pass of the decoding loops: the bf16 product A . B^T + bias with only the ST_SCREEN_SLOTS largest values (and
 * their columns) of every part (128 columns; 64 when N <= 128) kept per row -- cand_val / cand_idx (M, *npart_out,
 * ST_SCREEN_SLOTS), *npart_out * ST_SCREEN_SLOTS <= st_topk_parts(N).  A kept value carries its column's position in the
 * low 7 mantissa bits (relative perturbation < 2^-16; equal approximations rank lower column first).  The caller bounds
 * the bf16 rounding error plus that perturbation and re-scores the survivors exactly (st_vocab_topk_screen). */
enum { ST_SCREEN_SLOTS = 3 };
int st_gemm_bf16_screen(int M, int N, int K, const void* A_bf16, int lda, const void* B_bf16, int ldb, const float* bias,
                        float* cand_val, int32_t* cand_idx, int* npart_out, st_stream_t stream);
/* out[0] = max over rows of the row's 2-norm (W (rows, cols) fp32 contiguous): the weight factor of the screening bound. */
int st_row_norm_max(const float* W, int rows, int cols, float* out, st_stream_t stream);
/* Exact per-row top-K (K <= 8) of h . Wv^T + bv (rnn.py:51 `max(1)[1]`, rnn.py:63,90-91 `topk`) WITHOUT the (M, V) logits
 * and at a fraction of the fp32-accurate product's cost: bf16 screening GEMM (st_gemm_bf16_screen), a per-row bound of its
 * rounding error (c |h| wmax, wmax = st_row_norm_max(Wv)), exact fp32 re-scoring of every column that can still be in
 * the top K (decode.cu: screen_select_kernel states the argument).  h / Wv fp32 and their bf16 roundings (M, H) / (V, H),
 * H % 8 == 0; cand_val / cand_idx: scratch of M * st_topk_parts(V) entries.  Outputs as st_gemm_tf32x3_topk: value
 * descending, the lower column first among equal values. */
int st_vocab_topk_screen(int M, int V, int H, const float* h, const void* h_bf16, const float* Wv, const void* Wv_bf16,
                         const float* bv, const float* wmax, int K, float* cand_val, int32_t* cand_idx, float* val,
                         int32_t* idx, int out_stride, int64_t* tok, int tok_stride, st_stream_t stream);
/* Development / test aid: pin the kernel behind st_gemm_bf16 and st_vocab_ce_*: 2 = CTA-pair kernel
 * (cta_group::2, 256x256 tiles, stream-K), 128 / 256 = single-CTA kernel with that tile width, 0 = choose. */
int st_debug_gemm_variant(int variant);
int st_vocab_ce_fwd(int M, int V, int H, const void* Hs, int ldh, const void* Wv, int ldw, const float* bv,
                    const int64_t* target, float* part_max, float* part_sum, float* tlogit, float* lse,
                    float* loss_sum, st_stream_t stream);
int st_vocab_ce_bwd(int M, int V, int H, const void* Hs, int ldh, const void* Wv, int ldw, const float* bv,
                    const int64_t* target, const float* lse, float scale, void* P, int ldp, void* PT, int ldpt,
                    st_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Input packing (rnn.py:29-31 Embedding + cat + pack_padded_sequence; rnn_attn.py:101 + :70).
 * X (N, ldx) row n=(t,b):  with_feature=1: t==0 ? feature[b] : emb[caption[b, t-1]]   (rnn.py:30)
 *                          with_feature=0: emb[caption[b, t]]                          (rnn_attn.py:70)
 * caption is int64 (B, T_cap) row-major.  Only the first E columns of each X row are written.
 * ------------------------------------------------------------------------------------------ */
/*
 * Token ids outside [0, V) never index out of bounds: the access is clamped and the id is reported through
 * st_token_error (the reference's nn.Embedding / CrossEntropyLoss device-assert on such an id). */
int st_pack_inputs(float* X, int ldx, const float* emb, int E, int V, const float* feature,
                   const int64_t* caption, int T_cap, int with_feature, int nsteps,
                   const int* batch_sizes_host, st_stream_t stream);
/* Same rows written as bf16 (the operand of the hoisted tensor-core input projection): no fp32 copy + cast. */
int st_pack_inputs_bf16(void* X, int ldx, const float* emb, int E, int V, const float* feature,
                        const int64_t* caption, int T_cap, int with_feature, int nsteps,
                        const int* batch_sizes_host, st_stream_t stream);
/* Number of out-of-range token ids the packing kernels have met since the last clear (host-mapped status word,
 * no synchronisation: a step still in flight is seen by a later call); *first_bad_id = the first such id. */
int st_token_error(int64_t* first_bad_id, int clear);
/* Backward of the above: dEmb[token] += dX row (atomic), dfeature[b] = dX row (0,b).
 * dEmb must be zero-initialised (or hold the running gradient); dfeature may be NULL. */
int st_pack_inputs_bwd(const float* dX, int ldx, float* dEmb, int E, int V, float* dfeature,
                       const int64_t* caption, int T_cap, int with_feature, int nsteps,
                       const int* batch_sizes_host, st_stream_t stream);
/* Packed targets for the loss: out[n=(t,b)] = caption[b,t]  (main.py:145). */
int st_pack_targets(int64_t* out, const int64_t* caption, int T_cap, int V, int nsteps,
                    const int* batch_sizes_host, st_stream_t stream);

/* dst[i, :width] = table[idx[i * idx_stride], :]  (nn.Embedding lookup of the decode loops, rnn.py:53). */
int st_gather_rows(float* dst, int ld_dst, const float* table, int width, const int64_t* idx, int idx_stride,
                   int n, st_stream_t stream);

/* out[c] (+)= sum_r M[r, c]  -- bias gradients.  M fp32 or bf16; accumulate=0 overwrites. */
int st_colsum(float* out, const void* M, int m_is_bf16, int rows, int cols, int ld, int accumulate,
              st_stream_t stream);

/* dst[i][:] = src[i][:] * (*g) for n <= ST_SCALE_MAX fp32 tensors in one launch; g is a device scalar.
 * This is autograd's chain rule for forward_loss: loss.backward() (main.py:150) hands the step's
 * already-computed gradients a grad_output that is only known on the device. */
#define ST_SCALE_MAX 32
int st_scale_multi(int n, const float* const* src, float* const* dst, const int64_t* count, const float* g,
                   st_stream_t stream);

/* One recurrent step of the attention decoders with the context half of the input projection folded in
 * (rnn_attn.py:70 / rnn_attn_LSTM.py:72: unit([emb_t | embed(ctx_t)], h_{t-1}), step t of the packed sequence):
 *   gates = Gx[rows of t] + [h_{t-1} | X[rows of t]] . [Whh | Wx]^T + bhh   on tcgen05, one kernel per step.
 * Gx (N, G*H) fp32 = hoisted W_ih[:, :E] emb + b_ih; X (N, EX) bf16 = embed(ctx) rows (packed, ld = ldx);
 * Whh (G*H, H) / Wx (G*H, EX) bf16 K-major; h_{t-1} = h0 (t = 0: fp32 h0 + bf16 h0_bf16, (B0, H)) or rows of
 * step t-1 of Hs / Hs_bf16.  Writes rows of step t of Hs, Hs_bf16, Cs (LSTM), gates / ghn (may be NULL).
 * Replaces {st_gemm_bf16(beta = 1) into Gx; st_rnn_seq_tc_fwd over [t, t+1)}. */
int st_rnn_step_x_tc_supported(int kind, int H, int EX);
int st_rnn_step_x_tc_fwd(int kind, int H, int EX, int nsteps, const int* batch_sizes_host, int t, const float* Gx,
                         const void* X_bf16, int ldx, const void* Whh_bf16, const void* Wx_bf16, int ldwx,
                         const float* bhh, const float* h0, const void* h0_bf16, const float* c0, float* Hs,
                         void* Hs_bf16, float* Cs, float* gates, float* ghn, st_stream_t stream);

/* Backward of that step (step t of the reverse pass, layer 0 of the attention decoders): gate gradients as
 * st_rnn_seq_tc_bwd over [t, t+1) AND, from the same stream of the gate-gradient tile, d embed(ctx_t) =
 * dG_t . W_ih[:, E:] into rows of step t of dX (N, ldx) fp32 -- replaces {st_rnn_seq_tc_bwd; st_gemm_bf16}.
 * WhhT (H, G*H), WxT (EX = H, G*H; ld = ldwxt) bf16.  dstate (2, B0, H) carries dh / dc between the calls, which
 * must walk t = nsteps-1 ... 0 (the first one zeroes the barrier counters).  H % 64 == 0, context width == H.
 * Optional (both or neither; single-layer decoders): WdecT (H, A) bf16 = attn.decoder_att.weight^T and datt2 (N, A)
 * bf16, the attention-query gradients: the attention of step t+1 read h_t as its query (rnn_attn.py:69), so the
 * kernel first adds datt2[rows of t+1] . W_dec to the carried dh_t -- replaces the st_gemm_bf16(beta = 1) into
 * dstate after every attention backward except the last (t = 0), which the caller still issues. */
int st_rnn_step_x_tc_bwd(int kind, int H, int nsteps, const int* batch_sizes_host, int t, const void* WhhT_bf16,
                         const void* WxT_bf16, int ldwxt, const float* h0, const float* c0, const float* Hs,
                         const float* Cs, const float* gates, const float* ghn, const float* dHs, void* dG, void* dGT,
                         void* dGh, void* dGhT, int ldt, float* dstate, float* dX, int ldx, int* barrier,
                         const void* WdecT_bf16, int ldwdt, const void* datt2_bf16, int ldq, int A, st_stream_t stream);

/* Encoder head of the base models, cnn.py:37-38,49: nn.BatchNorm1d(E, momentum) over the rows of Y (B, E) =
 * Linear(2048, E)(pooled features) (the Linear product itself is st_sgemm / st_gemm_bf16).
 * Forward, training (use_running_stats = 0): batch statistics (biased variance), running_mean / running_var
 * (may be NULL) updated with `momentum` (unbiased variance), save_mean / save_invstd (E) kept for the backward.
 * Forward, eval: normalises with the running statistics.  Backward (training statistics): dgamma, dbeta (E) and dY. */
int st_bn1d_fwd(const float* Y, int ldy, int B, int E, const float* gamma, const float* beta, float eps, float momentum,
                int use_running_stats, float* running_mean, float* running_var, float* save_mean, float* save_invstd,
                float* out, int ldo, st_stream_t stream);
int st_bn1d_bwd(const float* Y, int ldy, const float* dOut, int ldd, int B, int E, const float* gamma,
                const float* save_mean, const float* save_invstd, float* dgamma, float* dbeta, float* dY, int ldg,
                st_stream_t stream);

/* Optimizer step over n <= ST_OPT_MAX fp32 tensors in one launch (main.py:97-100,152, main_attn.py:91-94,134:
 * torch.optim.SGD(lr, momentum) / torch.optim.Adam(lr)); arithmetic as torch's reference implementations.
 * param / grad / state are host arrays of device pointers, count[i] elements each.  momentum_buf may be NULL when
 * momentum == 0; first_step != 0 initialises the momentum buffers with the gradient (torch's first step).
 * `step` is Adam's 1-based step count.  grad_scale: optional device scalar multiplied into every gradient (the
 * grad_output of forward_loss's autograd node), NULL = 1.
 * shadow_bf16 (NULL, or n entries each NULL or a contiguous bf16 buffer of count[i] elements): rewritten with the
 * updated parameter in the same pass -- the bf16 operands of the next step's tensor-core kernels. */
#define ST_OPT_MAX 32
int st_sgd_step(int n, float* const* param, const float* const* grad, float* const* momentum_buf, void* const* shadow_bf16,
                const int64_t* count, float lr, float momentum, int first_step, const float* grad_scale, st_stream_t stream);
int st_adam_step(int n, float* const* param, const float* const* grad, float* const* exp_avg, float* const* exp_avg_sq,
                 void* const* shadow_bf16, const int64_t* count, float lr, double beta1, double beta2, float eps, int64_t step,
                 const float* grad_scale, st_stream_t stream);

/* out[r] = sum_c M[r, c], M bf16 (rows, ld): db_v from the transposed dlogits. */
int st_rowsum_bf16(float* out, const void* M, int rows, int cols, int ld, st_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Data-parallel gradient exchange.  The reference is single-process; under batch
 * sharding the step after main.py:150 `loss.backward()` / main_attn.py:133 is a sum of every
 * parameter gradient over the ranks before main.py:152 `optimizer.step()`.
 *
 * In-place sum all-reduce of `count` fp32 values held in a SYMMETRIC buffer: the same allocation on
 * every GPU of the node, rank r's copy mapped into this process at peers_host[r] (NVLink peer
 * mapping; peers_host[rank] is the local copy), optionally with one multicast mapping of all
 * copies (`multicast`, NULL if the fabric has none: multimem.ld_reduce / multimem.st run the sum
 * inside the NVSwitch).  flags_host[r]: rank r's flag area, st_allreduce_flag_words(world)
 * uint32, zeroed ONCE before the first call (the barriers restore them to zero).  One kernel of
 * `nblocks` CTAs per rank; every rank must issue the same sequence of calls with the same count and
 * nblocks.  Stream-ordered, CUDA-graph capturable; count % 4 == 0 and 16-byte aligned mappings. */
enum { ST_AR_MAX_WORLD = 8, ST_AR_MAX_BLOCKS = 64 };
int st_allreduce_flag_words(int world);
/* Wall-clock bound of a cross-GPU barrier wait inside st_allreduce_sum_f32 (default 5 minutes; ms <= 0 restores it).  A
 * wait that times out sets an error word instead of trapping: st_allreduce_error() returns 0, or 1 + the peer that did
 * not arrive (the results of that exchange are then undefined).  Reading it costs no synchronisation (mapped host word). */
int st_allreduce_set_timeout_ms(int64_t ms);
int st_allreduce_error(void);
/* Development / A-B aid: world == 2 uses peer loads / stores instead of the multicast reduce (default 1; every rank
 * must use the same setting). */
int st_debug_allreduce_pair_p2p(int on);
int st_allreduce_sum_f32(void* const* peers_host, void* multicast, void* const* flags_host, int rank, int world,
                         int64_t count, int nblocks, st_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Recurrent sequence, forward (the inside of nn.GRU / nn.LSTM over a PackedSequence,
 * rnn.py:32, rnn_lstm.py:30; one call per layer).  Persistent cooperative kernel: each CTA keeps
 * its slice of W_hh in shared memory for all steps, fuses h.W_hh^T with the gate math and the
 * state update, and meets the other CTAs of its batch tile at a device-wide barrier per step.
 *   Gx   (N, g*H)  input-side pre-activations W_ih x + b_ih, packed time-major
 *   Whh  (g*H, H), bhh (g*H)
 *   h0, c0 (B0, H) or NULL (= zeros); B0 = batch_sizes[0]
 *   Hs   (N, H)    out: h_t rows, packed time-major
 *   Cs   (N, H)    out (LSTM only): c_t rows
 *   gates(N, g*H)  out, saved for backward: GRU [r|z|n] post-activation, LSTM [i|f|g|o]
 *   ghn  (N, H)    out (GRU only): W_hn h + b_hn, saved for backward
 * gates/ghn may be NULL for inference.  barrier: >= 64 ints of scratch.
 * Only steps [t_begin, t_end) are run (0, nsteps for a whole sequence; single steps are used by
 * the decoders and by the attention models, whose step input depends on the previous state).
 * ------------------------------------------------------------------------------------------ */
int st_rnn_seq_fwd(int kind, int H, int nsteps, const int* batch_sizes_host, int t_begin, int t_end,
                   const float* Gx, const float* Whh, const float* bhh, const float* h0,
                   const float* c0, float* Hs, float* Cs, float* gates, float* ghn, int* barrier,
                   st_stream_t stream);

/* Backward through time of one layer (autograd of rnn.py:32), steps t_hi-1 down to t_lo.
 *   dHs    (N, H)    in: gradient w.r.t. every h_t row from above (vocab projection / next layer)
 *   dG     (N, g*H)  out: gradient w.r.t. Gx rows (= dgi)
 *   dGh    (N, g*H)  out: gradient w.r.t. (W_hh h + b_hh) rows; equals dG for LSTM (may alias),
 *                    differs in the n gate for GRU (da_n * r)
 *   dstate (2, B0, H) in/out: running gradient w.r.t. the carried state, [0] = dh, [1] = dc.
 *                    Rows are read only where a later step was processed in this or an earlier
 *                    call with the same buffer; after t_lo == 0 it holds dh0 / dc0.  A caller that
 *                    splits the range (attention models) may add further terms into the dh rows
 *                    between calls.
 *   barrier: >= 64 ints of scratch. */
int st_rnn_seq_bwd(int kind, int H, int nsteps, const int* batch_sizes_host, int t_hi, int t_lo,
                   const float* Whh, const float* h0, const float* c0, const float* Hs,
                   const float* Cs, const float* gates, const float* ghn, const float* dHs,
                   float* dG, float* dGh, float* dstate, int* barrier, st_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * The same recurrence on the tensor cores (bf16 mode): tcgen05.mma with the W_hh slice resident in
 * shared memory for all steps, h_{t-1} tiles TMA-loaded each step, fp32 accumulators in TMEM, the
 * gate math and the fp32 c / h state carried in registers across steps (persistent cooperative
 * launch over steps [t_begin, t_end) / t_hi-1..t_lo; a partial range resumes from the Hs / Cs rows
 * of step t_begin-1, resp. from `dstate`, which is also where a partial backward leaves the carried
 * gradient).  Returns ST_ERR_UNSUPPORTED when the shape does not qualify (st_rnn_seq_tc_supported) or
 * the grid is not co-resident -- the caller then uses st_rnn_seq_fwd.  dbih / dbhh may be NULL
 * (skipped; they are only produced by a call that reaches t_lo = 0).
 *   Whh_bf16 (g*H, H) bf16;  h0 (B0,H) fp32 together with its bf16 copy h0_bf16, or both NULL
 *   Hs (N,H) fp32 and Hs_bf16 (N,H) bf16 out; Cs, gates, ghn as in st_rnn_seq_fwd.
 * Backward: WhhT_bf16 (H, g*H) bf16 (the transpose).  Outputs the gate gradients directly as the
 * bf16 GEMM operands dG (N, g*H) and dGT (g*H, ldt >= N, multiple of 8) (GRU: also dGh / dGhT),
 * the bias gradients dbih / dbhh (g*H) and dstate (2, B0, H) = (dh0, dc0).
 * ------------------------------------------------------------------------------------------ */
/* Cluster-resident variant of st_rnn_seq_tc_fwd (rnn_cluster.cu): a thread-block cluster of H/32 CTAs
 * keeps the recurrence of a 16/32/48-row batch slice entirely in shared memory -- W_hh slices resident,
 * h_t exchanged through distributed shared memory, no global memory on the step-to-step chain.
 * Same arguments and outputs as st_rnn_seq_tc_fwd (no barrier workspace).  H in {64,128,256,512};
 * returns ST_ERR_UNSUPPORTED if the GPU cannot co-schedule clusters of that size. */
int st_rnn_cluster_supported(int kind, int H);
int st_rnn_cluster_fwd(int kind, int H, int nsteps, const int* batch_sizes_host, int t_begin, int t_end,
                       const float* Gx, const void* Whh_bf16, const float* bhh, const float* h0, const void* h0_bf16,
                       const float* c0, float* Hs, void* Hs_bf16, float* Cs, float* gates, float* ghn,
                       st_stream_t stream);
int st_rnn_seq_tc_supported(int kind, int H);
int st_rnn_seq_tc_fwd(int kind, int H, int nsteps, const int* batch_sizes_host, int t_begin, int t_end,
                      const float* Gx,
                      const void* Whh_bf16, const float* bhh, const float* h0, const void* h0_bf16,
                      const float* c0, float* Hs, void* Hs_bf16, float* Cs, float* gates, float* ghn,
                      int* barrier, st_stream_t stream);
int st_rnn_seq_tc_bwd(int kind, int H, int nsteps, const int* batch_sizes_host, int t_hi, int t_lo,
                      const void* WhhT_bf16,
                      const float* h0, const float* c0, const float* Hs, const float* Cs, const float* gates,
                      const float* ghn, const float* dHs, void* dG, void* dGT, void* dGh, void* dGhT, int ldt,
                      float* dbih, float* dbhh, float* dstate, int* barrier, st_stream_t stream);

/* Hprev (N,H): row (t,b) = h_{t-1}[b] (h0 or zeros at t=0) -- the B operand of dW_hh = dGh^T Hprev. */
int st_shift_states(float* Hprev, const float* Hs, const float* h0, int H, int nsteps,
                    const int* batch_sizes_host, st_stream_t stream);
/* Same on bf16 rows (Hs_bf16 of the tensor-core recurrent kernels -> the operand of dW_hh = dGh^T H_prev). */
int st_shift_states_bf16(void* Hprev, const void* Hs, const void* h0, int H, int nsteps,
                         const int* batch_sizes_host, st_stream_t stream);

/* Development / test aid: 1 = skip the bulk-copy relayout kernel (keeps the register-path kernels reachable for the
 * parity tests), 0 = default. */
int st_debug_relayout_legacy(int on);

/* ------------------------------------------------------------------------------------------
 * Soft attention (Attention/rnn_attn.py:8-31 Attention_Net, :60-76 rnn_iterator).  The reference
 * recomputes encoder_att(f) inside every step; here the state-independent parts are hoisted
 * (att1 = F W_enc^T + b_enc and Fe = F W_embed^T, both one GEMM per iteration) and each step runs
 * one fused kernel.  Storage type of F / att1 / Fe: fp32 (in_bf16 = 0) or bf16.
 * act: 0 = LeakyReLU(0.2) (the reference, rnn_attn.py:18), 1 = tanh.
 *
 * st_attn_relayout: f (B,C,P) fp32 channels-first (cnn_attn.py:49) -> F (B*P, C), optional
 *   FT (C, ldft >= B*P) (the K-major operand of dW_enc = datt1^T F), mean_f (B,C) (rnn_attn.py:62).
 * st_attn_step_fwd, one CTA per live batch row b < rows:
 *   e_p = w_f . act(att1[b,p,:] + att2[b,:]) + b_f;  alpha = softmax_P(e)            (rnn_attn.py:25-26)
 *   ctx_out[b,:E] = sum_p alpha_p Fe[b,p,:] + b_embed   (= embed(sum_p alpha_p f_p), rnn_attn.py:29,70)
 *   alphas[b*alpha_stride + p] = alpha_p (the caller points it at alphas[:, t, :]);  S[b,p] += alpha_p
 *   ctx_out_bf16 (optional): bf16 copy of ctx_out, the A operand of the tensor-core W_ih[:,E:] product
 * st_attn_step_bwd: dalpha_p = <dctx[b,:], Fe[b,p,:]> + dalpha[b*dalpha_stride + p] (may be NULL);
 *   de = alpha * (dalpha - sum alpha dalpha) -> de_out (rows,P);  datt2 (rows,A) = sum_p de_p w_f act'(.)
 *   (+ optional bf16 copy datt2_bf16 (rows,A); + optional gt (rows,A) = sum_p de_p act'(.), i.e. datt2 before the
 *   factor w_f: with it st_attn_hoist_bwd forms the step part of dw_f as sum att2 * gt instead of once more per tuple).
 *   Rows whose A*sizeof and E*sizeof are 512/1024/2048 bytes run the streaming kernels (bulk-async-copy ring,
 *   attn_stream.cu), other shapes a generic kernel.
 * st_attn_hoist_bwd (after the loop; de (N,P), att2 (N,A) packed time-major):
 *   datt1[b,p,a] = w_f[a] sum_t de[t,b,p] act'(att1[b,p,a] + att2[t,b,a]) (+ transposed copy),
 *   dwf[a] = sum_{t,b,p} de act(.)  (gt, optional: the (N,A) tensor st_attn_step_bwd wrote, see there)
 * st_attn_ctx_all: ctx[(t,b), c] = sum_p alphas[b,t,p] F[b,p,c] for all live (t,b) (for dW_embed); ctx / ctxT
 *   are fp32 (out_bf16 = 0) or bf16.
 * st_attn_penalty: pen_sum = sum (1 - S)^2, Gpen = -2 coef (1 - S)          (main_attn.py:131)
 * st_add_rows: dst[r,:] += src[r,:] (attention query gradient into the carried dh rows).
 * ------------------------------------------------------------------------------------------ */
int st_attn_relayout(const float* f, int B, int C, int P, void* F, void* FT, int ldft, int out_bf16,
                     float* mean_f, st_stream_t stream);
/* Same with the channels-first grid handed over in bf16 (a ResNet trunk run under autocast
 * produces it; halves the largest host-to-device / HBM read of the step). */
int st_attn_relayout_bf16in(const void* f_bf16, int B, int C, int P, void* F, void* FT, int ldft, int out_bf16,
                            float* mean_f, st_stream_t stream);
/* Channels-last grids: a trunk run in torch.channels_last hands the decoder (B, P, C) = F itself -- no re-layout pass;
 * this is the one thing still needed from the grid: mean_f (B, C) fp32 = the channel means over the P locations
 * (rnn_attn.py:62).  F fp32 or bf16, contiguous. */
int st_grid_mean_bpc(const void* F, int f_bf16, int B, int P, int C, float* mean_f, st_stream_t stream);
int st_attn_step_fwd(int rows, int P, int A, int E, const void* att1, const void* Fe, int in_bf16,
                     const float* att2, const float* wf, const float* bf, const float* b_embed, float* alphas,
                     int alpha_stride, float* S, float* ctx_out, int ld_ctx, void* ctx_out_bf16, int ld_ctx_bf16,
                     int act, st_stream_t stream);
int st_attn_step_bwd(int rows, int P, int A, int E, const void* att1, const void* Fe, int in_bf16,
                     const float* att2, const float* wf, const float* alphas, int alpha_stride,
                     const float* dalpha, int dalpha_stride, const float* dctx, int ld_dctx, float* de_out,
                     float* datt2, void* datt2_bf16, float* gt, int act, st_stream_t stream);
int st_attn_hoist_bwd(int nsteps, const int* batch_sizes_host, int P, int A, const void* att1, int in_bf16,
                      const float* att2, const float* de, const float* wf, void* datt1, void* datt1T, int ldt,
                      int out_bf16, float* dwf, const float* gt, int act, st_stream_t stream);
int st_attn_ctx_all(int nsteps, const int* batch_sizes_host, int P, int C, int T_cap, const void* F, int in_bf16,
                    const float* alphas, void* ctx, void* ctxT, int ldt, int out_bf16, st_stream_t stream);
/* Q[(b,p), e] = sum_t alphas[b,t,p] dctx[(t,b), e] over the steps in which row b is live, bf16 (B*P, E): the A operand
 * of dW_embed = Q^T . F (one tensor-core GEMM over the B*P locations) -- autograd of embed(sum_p alpha_p f_p),
 * rnn_attn.py:29,70 -- instead of rebuilding the (N, C) contexts with st_attn_ctx_all. */
int st_attn_embed_q(int nsteps, const int* batch_sizes_host, int P, int E, int T_cap, const float* alphas,
                    const float* dctx, int ld_dctx, void* Q_bf16, st_stream_t stream);
int st_attn_penalty(int n, const float* S, float coef, float* pen_sum, float* Gpen, st_stream_t stream);
int st_add_rows(float* dst, const float* src, int rows, int cols, st_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Cross-entropy over materialised logits (nn.CrossEntropyLoss(), main.py:94,149).
 *   logits (N, V) ld;  target (N) int64
 *   loss_sum: out, 1 float, sum_n (lse_n - logit_n[target_n])  (caller divides by the global N)
 *   lse (N) out, may be NULL
 *   dlogits: if not NULL, (N, V) ld written with (softmax - onehot) * grad_scale; may alias logits.
 * ------------------------------------------------------------------------------------------ */
int st_ce_fwd_bwd(const float* logits, int ld, const int64_t* target, int N, int V,
                  float* loss_sum, float* lse, float* dlogits, float grad_scale,
                  st_stream_t stream);

/* Row-wise arg-max (lowest index wins on ties, like Tensor.max(1)[1] on CPU; rnn.py:51) and
 * top-K (descending; ties by lower index first; rnn.py:63,90-91).  K <= 32. */
int st_argmax_rows(const float* X, int ld, int rows, int cols, int64_t* idx, int idx_stride,
                   st_stream_t stream);   /* idx[row * idx_stride] */
int st_topk_rows(const float* X, int ld, int rows, int cols, int K, float* val, int32_t* idx,
                 int out_stride, st_stream_t stream);   /* val/idx[row * out_stride + k] */

/* ------------------------------------------------------------------------------------------
 * Decoding.  All loops run inside the library on `stream` with no host synchronisation.
 * Weights: per layer l: Wih_l (g*H, in_l), Whh_l (g*H, H), bih_l, bhh_l (g*H); in_0 = E, in_l = H.
 * Passed as arrays of L device pointers living in HOST memory (*_host).
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  int kind, L, E, H, V;
  const float* emb;           /* (V, E) */
  const float* const* Wih_host;
  const float* const* Whh_host;
  const float* const* bih_host;
  const float* const* bhh_host;
  const float* Wv;            /* (V, H) */
  const float* bv;            /* (V)    */
  int gemm_mode;              /* nn.Linear products of the loops: 0 = fp32 CUDA cores, 1 = 3xTF32 tensor cores
                                 (fp32-accurate, st_gemm_tf32x3; needs E, H multiples of 4, else falls back to 0) */
} st_rnn_weights;

int64_t st_decode_workspace_bytes(const st_rnn_weights* w, int n_img, int K, int max_len);
/* Development / test aid: the projected-embedding table of gemm_mode 1 (EP = emb . W_ih^T + b_ih, built once per call):
 * 0 = by size (rows * max_len >= V), 1 = always, -1 = never.  Set before st_decode_workspace_bytes. */
int st_debug_decode_table(int mode);
/* ... and the vocabulary projection of gemm_mode 1: 0 / 1 = bf16 screening GEMM + exact fp32 re-scoring of the columns
 * that can still be among the top K (default), -1 = the 3xTF32 product with the top-K in its epilogue. */
int st_debug_decode_screen(int mode);

/* RNN.sentence_index(cnn_feature) greedy (rnn.py:44-58, rnn_lstm.py:35-57):
 * feature (n_img, E) -> tokens (n_img, max_len) int64. */
int st_decode_greedy(const st_rnn_weights* w, const float* feature, int n_img, int max_len,
                     int64_t* tokens, void* workspace, int64_t workspace_bytes, st_stream_t stream);

/* RNN.sentence_index(cnn_feature, beam_size=K) "chain" beam (rnn.py:60-108), batched over images
 * (the reference handles one image per call, main.py:81-82; images are independent).
 * tokens (n_img, max_len) int64 = old_beam_sentence[0].  Optional trace for parity checks:
 * trace_scores (max_len, n_img, K) scores of the K surviving sentences of each round (round 0: the
 * top-K logits), trace_words (max_len, n_img, K) surviving words (rnn.py:103 order).  May be NULL. */
int st_decode_beam_chain(const st_rnn_weights* w, const float* feature, int n_img, int K,
                         int max_len, int64_t* tokens, float* trace_scores, int32_t* trace_words,
                         void* workspace, int64_t workspace_bytes, st_stream_t stream);

/* beam_search.beam_search() semantics (beam_search.py:45-97) for the single-layer GRU decoder,
 * batched over images: per-hypothesis state, cumulative -log softmax cost, <end> termination,
 * hypotheses alive after max_length are dropped.
 * out_tokens (n_img, num_hyp, max_length+1) int32 (position 0 = start_id, padded with -1),
 * out_len (n_img, num_hyp) int32 (0 = no such hypothesis), out_cost (n_img, num_hyp) float. */
int st_decode_beam_tree(const st_rnn_weights* w, const float* feature, int n_img, int beam_width,
                        int num_hyp, int max_length, int start_id, int end_id, int32_t* out_tokens,
                        int32_t* out_len, float* out_cost, void* workspace,
                        int64_t workspace_bytes, st_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SHOWTELL_B200_H_ */
