// Persistent recurrent kernels, fp32: the inside of nn.GRU / nn.LSTM over a PackedSequence
// (rnn.py:32, LSTM/rnn_lstm.py:30) and its backward through time.
//
// Decomposition (forward): CTA (x = tile of UT=8 hidden units, y = tile of BT=128 batch rows).
// The CTA keeps the g*8 rows of W_hh that produce its units' gates in shared memory for the whole
// sequence.  Per step it streams its 128 rows of h_{t-1} (written by all unit tiles in the previous
// step, read back through L2) in 32-wide K chunks, accumulates the g*8 dot products per row in
// registers, applies the gate non-linearities and the state update for its (row, unit) pairs and
// writes h_t / c_t / saved gates.  The unit tiles of one batch tile then meet at a device-wide
// barrier (one counter per batch tile).  batch_sizes[t] shrinking is handled by masking rows;
// a batch tile whose rows are all finished leaves the loop.
//
// Backward: same ownership.  Phase 1 turns dh_t (from above + recurrent) into gate gradients for the
// CTA's own units (thread-local).  After the barrier, phase 2 forms
// dh_{t-1}[rows, own units] = dGh_t[rows, :] . W_hh[:, own units] with the W_hh^T slice resident in
// shared memory.  The weight gradients are hoisted GEMMs over all steps (done by the caller).
//
// Launch: cooperative when the grid is co-resident (always at the training shapes); otherwise, and
// for single steps (decoding), one ordinary launch per step / phase.
#include "common.cuh"

namespace st {
namespace {

constexpr int UT = 8, BT = 128, KC = 32, NT = 256, RPT = BT / 32, LDH = KC + 4;

struct SeqFwdParams {
  int H, t_begin, t_end;
  const float *Gx, *Whh, *bhh, *h0, *c0;
  float *Hs, *Cs, *gates, *ghn;
  int* barrier;
};

template <int G>
__global__ void __launch_bounds__(NT)
rnn_seq_fwd_kernel(const __grid_constant__ StepTable tab, const SeqFwdParams p) {
  extern __shared__ __align__(16) float smem[];
  const int H = p.H, ldw = H + 4;
  float* Ws = smem;                 // [G*UT][ldw]   row j = gate (j/UT), unit u0 + j%UT
  float* hS = smem + G * UT * ldw;  // [BT][LDH]
  const int tid = threadIdx.x, ul = tid & (UT - 1), rg = tid >> 3;
  const int u0 = blockIdx.x * UT, u = u0 + ul;
  const int r0 = blockIdx.y * BT;
  const bool uok = u < H;

  for (int idx = tid; idx < G * UT * H; idx += NT) {
    int j = idx / H, k = idx - j * H;
    int uu = u0 + (j % UT);
    Ws[j * ldw + k] = (uu < H) ? p.Whh[(size_t)((j / UT) * H + uu) * H + k] : 0.f;
  }
  float bh[G];
#pragma unroll
  for (int g = 0; g < G; ++g) bh[g] = uok ? p.bhh[g * H + u] : 0.f;
  __syncthreads();

  int nbar = 0;
  for (int t = p.t_begin; t < p.t_end; ++t) {
    const int nr = min(BT, tab.bs[t] - r0);
    if (nr <= 0) break;  // uniform over the CTAs of this batch tile; batch sizes only shrink
    const float* hprev = (t == 0) ? p.h0 : p.Hs + (size_t)tab.off[t - 1] * H;
    const float* cprev = (t == 0) ? p.c0 : p.Cs + (size_t)tab.off[t - 1] * H;

    float acc[RPT][G];
#pragma unroll
    for (int i = 0; i < RPT; ++i)
#pragma unroll
      for (int g = 0; g < G; ++g) acc[i][g] = 0.f;

    if (hprev != nullptr) {
      for (int k0 = 0; k0 < H; k0 += KC) {
        __syncthreads();
        for (int idx = tid; idx < BT * (KC / 4); idx += NT) {
          int r = idx >> 3, k4 = (idx & 7) * 4;
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (r < nr && k0 + k4 < H)
            v = __ldcg(reinterpret_cast<const float4*>(hprev + (size_t)(r0 + r) * H + k0 + k4));
          *reinterpret_cast<float4*>(hS + r * LDH + k4) = v;
        }
        __syncthreads();
        const int kmax = min(KC, H - k0);
        for (int kk = 0; kk < kmax; kk += 4) {
          float4 w[G];
#pragma unroll
          for (int g = 0; g < G; ++g)
            w[g] = *reinterpret_cast<const float4*>(Ws + (g * UT + ul) * ldw + k0 + kk);
#pragma unroll
          for (int i = 0; i < RPT; ++i) {
            float4 hv = *reinterpret_cast<const float4*>(hS + (rg + 32 * i) * LDH + kk);
#pragma unroll
            for (int g = 0; g < G; ++g) {
              acc[i][g] = fmaf(hv.x, w[g].x, acc[i][g]);
              acc[i][g] = fmaf(hv.y, w[g].y, acc[i][g]);
              acc[i][g] = fmaf(hv.z, w[g].z, acc[i][g]);
              acc[i][g] = fmaf(hv.w, w[g].w, acc[i][g]);
            }
          }
        }
      }
    }

#pragma unroll
    for (int i = 0; i < RPT; ++i) {
      const int r = rg + 32 * i;
      if (r >= nr || !uok) continue;
      const size_t n = (size_t)tab.off[t] + r0 + r;
      const float* gx = p.Gx + n * (size_t)(G * H);
      if (G == 4) {
        // nn.LSTM: c' = s(f) c + s(i) tanh(g); h' = s(o) tanh(c')
        const float ig = sigmoidf_(gx[u] + acc[i][0] + bh[0]);
        const float fg = sigmoidf_(gx[H + u] + acc[i][1] + bh[1]);
        const float gg = tanhf(gx[2 * H + u] + acc[i][2] + bh[2]);
        const float og = sigmoidf_(gx[3 * H + u] + acc[i][G - 1] + bh[G - 1]);
        const float cp = cprev ? __ldcg(cprev + (size_t)(r0 + r) * H + u) : 0.f;
        const float c2 = fmaf(fg, cp, ig * gg);
        p.Cs[n * H + u] = c2;
        p.Hs[n * H + u] = og * tanhf(c2);
        if (p.gates) {
          float* gs = p.gates + n * (size_t)(4 * H);
          gs[u] = ig; gs[H + u] = fg; gs[2 * H + u] = gg; gs[3 * H + u] = og;
        }
      } else {
        // nn.GRU: r = s(gi_r + gh_r); z = s(gi_z + gh_z); n = tanh(gi_n + r gh_n); h' = (1-z) n + z h
        const float ghn = acc[i][2] + bh[2];
        const float rr = sigmoidf_(gx[u] + acc[i][0] + bh[0]);
        const float zz = sigmoidf_(gx[H + u] + acc[i][1] + bh[1]);
        const float nn = tanhf(fmaf(rr, ghn, gx[2 * H + u]));
        const float hp = hprev ? __ldcg(hprev + (size_t)(r0 + r) * H + u) : 0.f;
        p.Hs[n * H + u] = fmaf(zz, hp - nn, nn);
        if (p.gates) {
          float* gs = p.gates + n * (size_t)(3 * H);
          gs[u] = rr; gs[H + u] = zz; gs[2 * H + u] = nn;
          p.ghn[n * H + u] = ghn;
        }
      }
    }
    if (t + 1 < p.t_end) {
      ++nbar;
      grid_barrier(p.barrier + blockIdx.y, nbar * (int)gridDim.x);
    }
  }
}

struct SeqBwdParams {
  int H, t_hi, t_lo;  // steps t_hi-1 ... t_lo
  int phase_mask;     // bit0: gate gradients, bit1: recurrent product
  const float *Whh, *h0, *c0, *Hs, *Cs, *gates, *ghn, *dHs;
  float *dG, *dGh, *dhrec, *dcrec;
  int* barrier;
};

template <int G>
__global__ void __launch_bounds__(NT)
rnn_seq_bwd_kernel(const __grid_constant__ StepTable tab, const SeqBwdParams p) {
  extern __shared__ __align__(16) float smem[];
  const int H = p.H, GH = G * H, ldw = GH + 4;
  float* WT = smem;             // [UT][ldw]   WT[ul][j] = Whh[j][u0+ul]
  float* dS = smem + UT * ldw;  // [BT][LDH]
  const int tid = threadIdx.x, ul = tid & (UT - 1), rg = tid >> 3;
  const int u0 = blockIdx.x * UT, u = u0 + ul;
  const int r0 = blockIdx.y * BT;
  const bool uok = u < H;

  if (p.phase_mask & 2) {
    for (int idx = tid; idx < UT * GH; idx += NT) {
      int j = idx / UT, l = idx - j * UT;
      WT[l * ldw + j] = (u0 + l < H) ? p.Whh[(size_t)j * H + u0 + l] : 0.f;
    }
  }
  __syncthreads();

  int nbar = 0;
  for (int t = p.t_hi - 1; t >= p.t_lo; --t) {
    const int nr = min(BT, tab.bs[t] - r0);
    if (nr <= 0) continue;  // this batch tile only becomes live at an earlier step
    const int b_next = tab.bs[t + 1];  // rows that were still live at t+1 (0 past the end)

    if (p.phase_mask & 1) {
#pragma unroll
      for (int i = 0; i < RPT; ++i) {
        const int r = rg + 32 * i;
        if (r >= nr || !uok) continue;
        const int b = r0 + r;
        const size_t n = (size_t)tab.off[t] + b;
        const bool has_rec = b < b_next;
        const float dh = p.dHs[n * H + u] + (has_rec ? p.dhrec[(size_t)b * H + u] : 0.f);
        if (G == 4) {
          const float* gs = p.gates + n * (size_t)(4 * H);
          const float ig = gs[u], fg = gs[H + u], gg = gs[2 * H + u], og = gs[3 * H + u];
          const float ct = p.Cs[n * H + u];
          const float cp = (t == 0) ? (p.c0 ? p.c0[(size_t)b * H + u] : 0.f)
                                    : p.Cs[((size_t)tab.off[t - 1] + b) * H + u];
          const float tc = tanhf(ct);
          const float dc = fmaf(dh * og, 1.f - tc * tc, has_rec ? p.dcrec[(size_t)b * H + u] : 0.f);
          const float da_i = dc * gg * ig * (1.f - ig);
          const float da_f = dc * cp * fg * (1.f - fg);
          const float da_g = dc * ig * (1.f - gg * gg);
          const float da_o = dh * tc * og * (1.f - og);
          float* d = p.dG + n * (size_t)(4 * H);
          d[u] = da_i; d[H + u] = da_f; d[2 * H + u] = da_g; d[3 * H + u] = da_o;
          if (p.dGh != p.dG) {
            float* e = p.dGh + n * (size_t)(4 * H);
            e[u] = da_i; e[H + u] = da_f; e[2 * H + u] = da_g; e[3 * H + u] = da_o;
          }
          p.dcrec[(size_t)b * H + u] = dc * fg;
          p.dhrec[(size_t)b * H + u] = 0.f;
        } else {
          const float* gs = p.gates + n * (size_t)(3 * H);
          const float rr = gs[u], zz = gs[H + u], nn = gs[2 * H + u];
          const float ghn = p.ghn[n * H + u];
          const float hp = (t == 0) ? (p.h0 ? p.h0[(size_t)b * H + u] : 0.f)
                                    : p.Hs[((size_t)tab.off[t - 1] + b) * H + u];
          const float da_n = dh * (1.f - zz) * (1.f - nn * nn);
          const float da_z = dh * (hp - nn) * zz * (1.f - zz);
          const float da_r = da_n * ghn * rr * (1.f - rr);
          float* d = p.dG + n * (size_t)(3 * H);
          d[u] = da_r; d[H + u] = da_z; d[2 * H + u] = da_n;
          float* e = p.dGh + n * (size_t)(3 * H);
          e[u] = da_r; e[H + u] = da_z; e[2 * H + u] = da_n * rr;
          p.dhrec[(size_t)b * H + u] = dh * zz;
        }
      }
    }
    if ((p.phase_mask & 3) == 3) {
      ++nbar;
      grid_barrier(p.barrier + blockIdx.y, nbar * (int)gridDim.x);
    }
    if (p.phase_mask & 2) {
      float acc[RPT];
#pragma unroll
      for (int i = 0; i < RPT; ++i) acc[i] = 0.f;
      const float* src = p.dGh + (size_t)tab.off[t] * GH;
      for (int k0 = 0; k0 < GH; k0 += KC) {
        __syncthreads();
        for (int idx = tid; idx < BT * (KC / 4); idx += NT) {
          int r = idx >> 3, k4 = (idx & 7) * 4;
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (r < nr && k0 + k4 < GH)
            v = __ldcg(reinterpret_cast<const float4*>(src + (size_t)(r0 + r) * GH + k0 + k4));
          *reinterpret_cast<float4*>(dS + r * LDH + k4) = v;
        }
        __syncthreads();
        const int kmax = min(KC, GH - k0);
        for (int kk = 0; kk < kmax; kk += 4) {
          const float4 w = *reinterpret_cast<const float4*>(WT + ul * ldw + k0 + kk);
#pragma unroll
          for (int i = 0; i < RPT; ++i) {
            const float4 dv = *reinterpret_cast<const float4*>(dS + (rg + 32 * i) * LDH + kk);
            acc[i] = fmaf(dv.x, w.x, acc[i]);
            acc[i] = fmaf(dv.y, w.y, acc[i]);
            acc[i] = fmaf(dv.z, w.z, acc[i]);
            acc[i] = fmaf(dv.w, w.w, acc[i]);
          }
        }
      }
#pragma unroll
      for (int i = 0; i < RPT; ++i) {
        const int r = rg + 32 * i;
        if (r < nr && uok) p.dhrec[(size_t)(r0 + r) * H + u] += acc[i];
      }
    }
  }
}

size_t fwd_smem(int G, int H) { return sizeof(float) * ((size_t)G * UT * (H + 4) + BT * LDH); }
size_t bwd_smem(int G, int H) { return sizeof(float) * ((size_t)UT * (G * H + 4) + BT * LDH); }

template <typename Kern>
int max_coresident(Kern kern, size_t smem, int* out) {
  int dev = 0, sms = 0, per_sm = 0;
  ST_CUDA_TRY(cudaGetDevice(&dev));
  ST_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  ST_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NT, smem));
  *out = sms * per_sm;
  return ST_OK;
}

template <int G>
int launch_fwd(const StepTable& tab, SeqFwdParams p, cudaStream_t s) {
  auto kern = rnn_seq_fwd_kernel<G>;
  const size_t smem = fwd_smem(G, p.H);
  int64_t optin = 0;
  ST_TRY(st_device_info(nullptr, nullptr, nullptr, &optin));
  ST_REQUIRE((int64_t)smem <= optin, ST_ERR_BAD_SHAPE,
             "rnn_seq_fwd: H=%d needs %zu B of shared memory for the resident W_hh slice (> %lld)",
             p.H, smem, (long long)optin);
  ST_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int b0 = tab.bs[p.t_begin];
  dim3 grid((p.H + UT - 1) / UT, (b0 + BT - 1) / BT);
  ST_REQUIRE(grid.y <= 65535, ST_ERR_BAD_SHAPE, "rnn_seq_fwd: batch %d too large", b0);
  int cores = 0;
  ST_TRY(max_coresident(kern, smem, &cores));
  const bool multi = p.t_end - p.t_begin > 1;
  if (multi && grid.y <= 64 && (int)(grid.x * grid.y) <= cores) {
    ST_CUDA_TRY(cudaMemsetAsync(p.barrier, 0, sizeof(int) * 64, s));
    void* args[] = {(void*)&tab, (void*)&p};
    ST_CUDA_TRY(cudaLaunchCooperativeKernel((const void*)kern, grid, dim3(NT), args, smem, s));
    note_launch();
  } else {
    for (int t = p.t_begin; t < p.t_end; ++t) {
      SeqFwdParams q = p;
      q.t_begin = t;
      q.t_end = t + 1;
      dim3 g((p.H + UT - 1) / UT, (tab.bs[t] + BT - 1) / BT);
      kern<<<g, NT, smem, s>>>(tab, q);
      ST_LAUNCH_TRY("rnn_seq_fwd_kernel");
    }
  }
  return ST_OK;
}

template <int G>
int launch_bwd(const StepTable& tab, SeqBwdParams p, cudaStream_t s) {
  auto kern = rnn_seq_bwd_kernel<G>;
  const size_t smem = bwd_smem(G, p.H);
  int64_t optin = 0;
  ST_TRY(st_device_info(nullptr, nullptr, nullptr, &optin));
  ST_REQUIRE((int64_t)smem <= optin, ST_ERR_BAD_SHAPE,
             "rnn_seq_bwd: H=%d needs %zu B of shared memory for the resident W_hh^T slice (> %lld)",
             p.H, smem, (long long)optin);
  ST_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int b0 = tab.bs[p.t_lo];
  dim3 grid((p.H + UT - 1) / UT, (b0 + BT - 1) / BT);
  ST_REQUIRE(grid.y <= 65535, ST_ERR_BAD_SHAPE, "rnn_seq_bwd: batch %d too large", b0);
  int cores = 0;
  ST_TRY(max_coresident(kern, smem, &cores));
  if (grid.y <= 64 && (int)(grid.x * grid.y) <= cores) {
    p.phase_mask = 3;
    ST_CUDA_TRY(cudaMemsetAsync(p.barrier, 0, sizeof(int) * 64, s));
    void* args[] = {(void*)&tab, (void*)&p};
    ST_CUDA_TRY(cudaLaunchCooperativeKernel((const void*)kern, grid, dim3(NT), args, smem, s));
    note_launch();
  } else {
    for (int t = p.t_hi - 1; t >= p.t_lo; --t) {
      dim3 g((p.H + UT - 1) / UT, (tab.bs[t] + BT - 1) / BT);
      for (int phase = 1; phase <= 2; ++phase) {
        SeqBwdParams q = p;
        q.t_hi = t + 1;
        q.t_lo = t;
        q.phase_mask = phase;
        kern<<<g, NT, smem, s>>>(tab, q);
        ST_LAUNCH_TRY("rnn_seq_bwd_kernel");
      }
    }
  }
  return ST_OK;
}

}  // namespace
}  // namespace st

extern "C" {

int st_rnn_seq_fwd(int kind, int H, int nsteps, const int* batch_sizes_host, int t_begin, int t_end,
                   const float* Gx, const float* Whh, const float* bhh, const float* h0,
                   const float* c0, float* Hs, float* Cs, float* gates, float* ghn, int* barrier,
                   st_stream_t stream) {
  using namespace st;
  StepTable tab;
  ST_TRY(make_step_table(tab, nsteps, batch_sizes_host));
  ST_REQUIRE(kind == ST_GRU || kind == ST_LSTM, ST_ERR_UNSUPPORTED, "st_rnn_seq_fwd: kind=%d", kind);
  ST_REQUIRE(H >= 4 && H % 4 == 0, ST_ERR_BAD_SHAPE, "st_rnn_seq_fwd: H=%d must be a multiple of 4", H);
  ST_REQUIRE(0 <= t_begin && t_begin < t_end && t_end <= nsteps, ST_ERR_BAD_SHAPE,
             "st_rnn_seq_fwd: step range [%d,%d) outside [0,%d)", t_begin, t_end, nsteps);
  ST_REQUIRE(Gx && Whh && bhh && Hs && barrier, ST_ERR_NULL, "st_rnn_seq_fwd: NULL pointer");
  ST_REQUIRE(kind == ST_GRU || Cs, ST_ERR_NULL, "st_rnn_seq_fwd: LSTM needs Cs");
  ST_REQUIRE(kind == ST_LSTM || !gates || ghn, ST_ERR_NULL, "st_rnn_seq_fwd: GRU gates need ghn");
  SeqFwdParams p{H, t_begin, t_end, Gx, Whh, bhh, h0, c0, Hs, Cs, gates, ghn, barrier};
  return kind == ST_LSTM ? launch_fwd<4>(tab, p, as_stream(stream))
                         : launch_fwd<3>(tab, p, as_stream(stream));
}

int st_rnn_seq_bwd(int kind, int H, int nsteps, const int* batch_sizes_host, int t_hi, int t_lo,
                   const float* Whh, const float* h0, const float* c0, const float* Hs,
                   const float* Cs, const float* gates, const float* ghn, const float* dHs,
                   float* dG, float* dGh, float* dstate, int* barrier, st_stream_t stream) {
  using namespace st;
  StepTable tab;
  ST_TRY(make_step_table(tab, nsteps, batch_sizes_host));
  ST_REQUIRE(kind == ST_GRU || kind == ST_LSTM, ST_ERR_UNSUPPORTED, "st_rnn_seq_bwd: kind=%d", kind);
  ST_REQUIRE(H >= 4 && H % 4 == 0, ST_ERR_BAD_SHAPE, "st_rnn_seq_bwd: H=%d must be a multiple of 4", H);
  ST_REQUIRE(0 <= t_lo && t_lo < t_hi && t_hi <= nsteps, ST_ERR_BAD_SHAPE,
             "st_rnn_seq_bwd: step range [%d,%d) outside [0,%d)", t_lo, t_hi, nsteps);
  ST_REQUIRE(Whh && Hs && gates && dHs && dG && dGh && dstate && barrier, ST_ERR_NULL,
             "st_rnn_seq_bwd: NULL pointer");
  ST_REQUIRE(kind == ST_GRU || Cs, ST_ERR_NULL, "st_rnn_seq_bwd: LSTM needs Cs");
  ST_REQUIRE(kind == ST_LSTM || ghn, ST_ERR_NULL, "st_rnn_seq_bwd: GRU needs ghn");
  SeqBwdParams p{H, t_hi, t_lo, 3, Whh, h0, c0, Hs, Cs, gates, ghn, dHs, dG, dGh,
                 dstate, dstate + (size_t)tab.bs[0] * H, barrier};
  return kind == ST_LSTM ? launch_bwd<4>(tab, p, as_stream(stream))
                         : launch_bwd<3>(tab, p, as_stream(stream));
}

}  // extern "C"
