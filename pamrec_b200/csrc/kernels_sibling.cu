// Kernels of the sibling multi-task baselines: MMoEModel_original (models/sequential/mmoe.py),
// PLEModel (ple.py), ShareBottomModel (sharebottom.py).  What differs from PAMRec is the front end - two DIN-style attention
// poolings (`_attention_fcn`, mmoe.py:299-337) instead of the time-aware encoder - and the mixing layer; every dense layer,
// batch norm, the clip / Adam step and the sparse backward reuse the kernels of kernels_head.cu / kernels_optim.cu.
//
// Branch 0 = long_term (satisfied-only history, mmoe.py:199-207), branch 1 = short_term (full history, mmoe.py:208-216).
// Arrays indexed [2][N][.] use the ACTUAL token count N = B * T of the batch as the branch stride, so that the lookups of one
// table are one contiguous key list for the sparse plan.
#include "head_tiles.cuh"

namespace pamrec {

// ------------------------------------------------------------------------------------------
// gather: h[r][n] = item_w[id] | cate_w[cid] for both histories, tgt[b] = item_w[item] | cate_w[cate]; the ids are copied into
// branch order for the sparse plan.  One thread per 16-byte chunk (5 per token).
__global__ void __launch_bounds__(256)
k_sib_gather(const int* __restrict__ sat_item, const int* __restrict__ sat_cate, const int* __restrict__ hist_item,
             const int* __restrict__ hist_cate, const int* __restrict__ items, const int* __restrict__ cates,
             const float* __restrict__ item_w, const float* __restrict__ cate_w, float* __restrict__ h, float* __restrict__ tgt,
             int* __restrict__ ids_item, int* __restrict__ ids_cate, int64_t N, int B) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t n_hist = 2 * N * 5;
  if (i < n_hist) {
    const int64_t tok = i / 5;                // r * N + n
    const int c = (int)(i % 5);
    const int r = tok >= N ? 1 : 0;
    const int64_t n = tok - (int64_t)r * N;
    const int id = r == 0 ? sat_item[n] : hist_item[n];
    const int cid = r == 0 ? sat_cate[n] : hist_cate[n];
    if (c == 0) { if (ids_item) ids_item[tok] = id; if (ids_cate) ids_cate[tok] = cid; }
    const float4 v = c < 4 ? ld4(item_w + (int64_t)id * kI + 4 * c) : ld4(cate_w + (int64_t)cid * kC);
    st4(h + tok * kE + 4 * c, v);
  } else if (i < n_hist + (int64_t)B * 5) {
    const int64_t j = i - n_hist;
    const int b = (int)(j / 5), c = (int)(j % 5);
    const float4 v = c < 4 ? ld4(item_w + (int64_t)items[b] * kI + 4 * c) : ld4(cate_w + (int64_t)cates[b] * kC);
    st4(tgt + (int64_t)b * kE + 4 * c, v);
  }
}
void launch_sib_gather(const int* sat_item, const int* sat_cate, const int* hist_item, const int* hist_cate, const int* items,
                       const int* cates, const float* item_w, const float* cate_w, float* h, float* tgt, int* ids_item,
                       int* ids_cate, int B, int T, cudaStream_t st) { PAMREC_PROF("sib_gather", 1, st);
  if (B == 0) return;
  const int64_t N = (int64_t)B * T, total = 2 * N * 5 + (int64_t)B * 5;
  k_sib_gather<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(sat_item, sat_cate, hist_item, hist_cate, items, cates, item_w, cate_w,
                                                               h, tgt, ids_item, ids_cate, N, B);
}

// has0[t] = 1 when id 0 is among the full-history or target ids of table t (0 item, 1 category): the rows that carry the L2
// term are tf.unique(history ids, target ids) (sequential_base_model.py:640-664); the satisfied-only history is a subset of the
// history EXCEPT for its padding id 0, so row 0 may be looked up (and receive a gradient) without being an L2 row.
__global__ void k_sib_has0(const int* __restrict__ hist_item, const int* __restrict__ hist_cate, const int* __restrict__ items,
                           const int* __restrict__ cates, int64_t N, int B, int* __restrict__ has0) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) {
    if (hist_item[i] == 0) has0[0] = 1;
    if (hist_cate[i] == 0) has0[1] = 1;
  } else if (i < N + B) {
    if (items[i - N] == 0) has0[0] = 1;
    if (cates[i - N] == 0) has0[1] = 1;
  }
}
void launch_sib_has0(const int* hist_item, const int* hist_cate, const int* items, const int* cates, int B, int T, int* has0,
                     cudaStream_t st) { PAMREC_PROF("sib_has0", 1, st);
  cudaMemsetAsync(has0, 0, 4 * sizeof(int), st);
  if (B == 0) return;
  const int64_t N = (int64_t)B * T;
  k_sib_has0<<<(unsigned)((N + B + 255) / 256), 256, 0, st>>>(hist_item, hist_cate, items, cates, N, B, has0);
}

// ------------------------------------------------------------------------------------------
// feature row of the attention MLP (mmoe.py:320-326): a = h A is already in feat[n, 80 r + 0:20]; fill q | a - q | a * q
__global__ void __launch_bounds__(256) k_sib_feat_fwd(float* __restrict__ feat, const float* __restrict__ tgt, int64_t N, int T) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * 40) return;
  const int64_t n = i / 40;
  const int r = (int)((i % 40) / kE), j = (int)(i % kE);
  const int64_t b = n / T;
  float* f = feat + n * 160 + r * 80;
  const float a = f[j], q = tgt[b * kE + j];
  f[kE + j] = q;
  f[2 * kE + j] = a - q;
  f[3 * kE + j] = a * q;
}
void launch_sib_feat_fwd(float* feat, const float* tgt, int B, int T, cudaStream_t st) { PAMREC_PROF("sib_feat_fwd", 1, st);
  if (B == 0) return;
  const int64_t N = (int64_t)B * T;
  k_sib_feat_fwd<<<(unsigned)((N * 40 + 255) / 256), 256, 0, st>>>(feat, tgt, N, T);
}

// gradient of the feature row: d_a = df[0:20] + df[40:60] + df[60:80] * q  (-> d_att[r][n]);
// d_q[b][r] = sum_t df[20:40] - df[40:60] + df[60:80] * a.   CTA = (sample, branch), thread = (t mod 6, feature).
__global__ void __launch_bounds__(128)
k_sib_feat_bwd(const float* __restrict__ d_feat, const float* __restrict__ feat, const float* __restrict__ tgt,
               float* __restrict__ d_att, float* __restrict__ dq, int64_t N, int T) {
  __shared__ float red[6][kE];
  const int b = blockIdx.x, r = blockIdx.y, tid = threadIdx.x;
  const int ty = tid / kE, j = tid % kE;
  float acc = 0.f;
  if (tid < 6 * kE) {
    const float q = tgt[(int64_t)b * kE + j];
    for (int t = ty; t < T; t += 6) {
      const int64_t n = (int64_t)b * T + t;
      const float* df = d_feat + n * 160 + r * 80;
      const float a = feat[n * 160 + r * 80 + j];
      const float d3 = df[3 * kE + j], d2 = df[2 * kE + j];
      d_att[((int64_t)r * N + n) * kE + j] = df[j] + d2 + d3 * q;
      acc += df[kE + j] - d2 + d3 * a;
    }
    red[ty][j] = acc;
  }
  __syncthreads();
  if (tid < kE) {
    float s = 0.f;
#pragma unroll
    for (int y = 0; y < 6; ++y) s += red[y][tid];
    dq[((int64_t)b * 2 + r) * kE + tid] = s;
  }
}
void launch_sib_feat_bwd(const float* d_feat, const float* feat, const float* tgt, float* d_att, float* dq, int B, int T,
                         cudaStream_t st) { PAMREC_PROF("sib_feat_bwd", 1, st);
  if (B == 0) return;
  k_sib_feat_bwd<<<dim3(B, 2), 128, 0, st>>>(d_feat, feat, tgt, d_att, dq, (int64_t)B * T, T);
}

// ------------------------------------------------------------------------------------------
// masked softmax over the history + weighted sum (mmoe.py:330-337, then reduce_sum over the sequence, mmoe.py:205-216):
// w = softmax_t(mask == 1 ? score : -(2^32)+1);  x[b, 20 r + j] = sum_t w_t h[r][b,t,j];  x[b, 40:60] = target.  Warp = (b, r).
__device__ __forceinline__ void sib_weights(const float* __restrict__ score, const int* __restrict__ mask, int64_t base, int T, int r,
                                            int lane, float* aw) {
  float s[8];
  float m = -INFINITY;
#pragma unroll
  for (int jj = 0; jj < 8; ++jj) {
    const int t = jj * 32 + lane;
    float val = -INFINITY;
    if (t < T) val = (mask[base + t] == 1) ? score[(base + t) * 2 + r] : kMaskNeg;
    s[jj] = val;
    m = fmaxf(m, val);
  }
  m = warp_max(m);
  float sum = 0.f;
#pragma unroll
  for (int jj = 0; jj < 8; ++jj) {
    const int t = jj * 32 + lane;
    const float e = (t < T) ? expf(s[jj] - m) : 0.f;
    s[jj] = e;
    sum += e;
  }
  sum = warp_sum(sum);
#pragma unroll
  for (int jj = 0; jj < 8; ++jj) {
    const int t = jj * 32 + lane;
    if (t < T) aw[t] = s[jj] / sum;
  }
  __syncwarp();
}

__global__ void __launch_bounds__(128)
k_sib_pool_fwd(const float* __restrict__ h, const float* __restrict__ score, const int* __restrict__ sat_mask,
               const int* __restrict__ mask, const float* __restrict__ tgt, float* __restrict__ aw_out, float* __restrict__ x,
               int B, int T) {
  __shared__ float aws[4][PAMREC_MAX_T];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int unit = blockIdx.x * 4 + w;
  if (unit >= 2 * B) return;
  const int b = unit >> 1, r = unit & 1;
  const int64_t N = (int64_t)B * T, base = (int64_t)b * T;
  float* aw = aws[w];
  sib_weights(score, r == 0 ? sat_mask : mask, base, T, r, lane, aw);
  for (int t = lane; t < T; t += 32) aw_out[(int64_t)r * N + base + t] = aw[t];
  if (lane < kE) {
    const float* hr = h + ((int64_t)r * N + base) * kE + lane;
    float acc = 0.f;
    for (int t = 0; t < T; ++t) acc = fmaf(aw[t], hr[(int64_t)t * kE], acc);
    x[(int64_t)b * 60 + r * kE + lane] = acc;
    if (r == 0) x[(int64_t)b * 60 + 2 * kE + lane] = tgt[(int64_t)b * kE + lane];
  }
}
void launch_sib_pool_fwd(const float* h, const float* score, const int* sat_mask, const int* mask, const float* tgt, float* aw,
                         float* x, int B, int T, cudaStream_t st) { PAMREC_PROF("sib_pool_fwd", 1, st);
  if (B == 0) return;
  k_sib_pool_fwd<<<(2 * B + 3) / 4, 128, 0, st>>>(h, score, sat_mask, mask, tgt, aw, x, B, T);
}

// backward: d_out = d_x[b, 20 r : 20 r + 20];  dw_t = d_out . h_t;  d_score_t = mask ? w_t (dw_t - sum_s w_s dw_s) : 0;
// dh[r][n] = w_t d_out  (the attention_mat path is added afterwards by the dX GEMM of d_att)
__global__ void __launch_bounds__(128)
k_sib_pool_bwd(const float* __restrict__ h, const float* __restrict__ aw_in, const int* __restrict__ sat_mask,
               const int* __restrict__ mask, const float* __restrict__ d_x, float* __restrict__ d_score, float* __restrict__ dh,
               int B, int T) {
  __shared__ __align__(16) float dout[4][kE];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int unit = blockIdx.x * 4 + w;
  if (unit >= 2 * B) return;
  const int b = unit >> 1, r = unit & 1;
  const int64_t N = (int64_t)B * T, base = (int64_t)b * T;
  const int* mk = r == 0 ? sat_mask : mask;
  if (lane < kE) dout[w][lane] = d_x[(int64_t)b * 60 + r * kE + lane];
  __syncwarp();
  float4 d[5];
#pragma unroll
  for (int c = 0; c < 5; ++c) d[c] = ld4(&dout[w][4 * c]);
  float dw[8], a[8];
  float dot = 0.f;
#pragma unroll
  for (int jj = 0; jj < 8; ++jj) {
    const int t = jj * 32 + lane;
    float v = 0.f, at = 0.f;
    if (t < T) {
      const float* hr = h + ((int64_t)r * N + base + t) * kE;
#pragma unroll
      for (int c = 0; c < 5; ++c) v += f4_dot(d[c], ld4(hr + 4 * c));
      at = aw_in[(int64_t)r * N + base + t];
      dot = fmaf(at, v, dot);
    }
    dw[jj] = v; a[jj] = at;
  }
  dot = warp_sum(dot);
#pragma unroll
  for (int jj = 0; jj < 8; ++jj) {
    const int t = jj * 32 + lane;
    if (t < T) {
      d_score[(base + t) * 2 + r] = (mk[base + t] == 1) ? a[jj] * (dw[jj] - dot) : 0.f;
      float* o = dh + ((int64_t)r * N + base + t) * kE;
#pragma unroll
      for (int c = 0; c < 5; ++c) st4(o + 4 * c, make_float4(a[jj] * d[c].x, a[jj] * d[c].y, a[jj] * d[c].z, a[jj] * d[c].w));
    }
  }
}
void launch_sib_pool_bwd(const float* h, const float* aw, const int* sat_mask, const int* mask, const float* d_x, float* d_score,
                         float* dh, int B, int T, cudaStream_t st) { PAMREC_PROF("sib_pool_bwd", 1, st);
  if (B == 0) return;
  k_sib_pool_bwd<<<(2 * B + 3) / 4, 128, 0, st>>>(h, aw, sat_mask, mask, d_x, d_score, dh, B, T);
}

// gradient of the target rows: towers' target columns (d_tgt, MMoE / PLE) + x's target columns + the query of both branches
__global__ void k_sib_tgt_total(const float* __restrict__ d_tgt, const float* __restrict__ d_x, const float* __restrict__ dq,
                                float* __restrict__ out, int B) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * kE) return;
  const int b = i / kE, j = i % kE;
  float s = d_x[(int64_t)b * 60 + 2 * kE + j] + dq[((int64_t)b * 2) * kE + j] + dq[((int64_t)b * 2 + 1) * kE + j];
  if (d_tgt != nullptr) s += d_tgt[i];
  out[i] = s;
}
void launch_sib_tgt_total(const float* d_tgt, const float* d_x, const float* dq, float* out, int B, cudaStream_t st) {
  PAMREC_PROF("sib_tgt_total", 1, st);
  if (B == 0) return;
  k_sib_tgt_total<<<(B * kE + 255) / 256, 256, 0, st>>>(d_tgt, d_x, dq, out, B);
}

// ------------------------------------------------------------------------------------------
// mixing (mmoe.py:43-50, ple.py:51-58): a gate is a BN + ReLU MLP like any other (no softmax);
// out_g = sum_j gate_g[j] * expert_{sel[g][j]};  U = main | tgt | sub | tgt
struct MixSel { int s[2][5]; };
__global__ void __launch_bounds__(64)
k_sib_mix_fwd(const float* __restrict__ ZE1, const float* __restrict__ ZG1, BnSet e1, BnSet g1, const float* __restrict__ tgt,
              float* __restrict__ U, int ldE, MixSel sel) {
  __shared__ float gt[10];
  const int b = blockIdx.x, c = threadIdx.x;
  if (c < 10) gt[c] = bn_relu(ZG1[(int64_t)b * 10 + c], g1.stat, g1.gamma, g1.beta, c);
  __syncthreads();
  float mn = 0.f, sb = 0.f;
#pragma unroll
  for (int j = 0; j < 5; ++j) {
    const int em = sel.s[0][j] * 64 + c, es = sel.s[1][j] * 64 + c;
    mn = fmaf(gt[j], bn_relu(ZE1[(int64_t)b * ldE + em], e1.stat, e1.gamma, e1.beta, em), mn);
    sb = fmaf(gt[5 + j], bn_relu(ZE1[(int64_t)b * ldE + es], e1.stat, e1.gamma, e1.beta, es), sb);
  }
  float* u = U + (int64_t)b * 168;
  u[c] = mn;
  u[84 + c] = sb;
  if (c < kE) { const float t = tgt[(int64_t)b * kE + c]; u[64 + c] = t; u[148 + c] = t; }
}
void launch_sib_mix_fwd(const float* ZE1, const float* ZG1, const BnSet& e1, const BnSet& g1, const float* tgt, float* U, int n_expert,
                        const int sel[2][5], int B, cudaStream_t st) { PAMREC_PROF("sib_mix_fwd", 1, st);
  if (B == 0) return;
  MixSel s;
  for (int g = 0; g < 2; ++g) for (int j = 0; j < 5; ++j) s.s[g][j] = sel[g][j];
  k_sib_mix_fwd<<<B, 64, 0, st>>>(ZE1, ZG1, e1, g1, tgt, U, n_expert * 64, s);
}

// dE1[e] = sum over (gate g, slot j) with sel[g][j] == e of gate_g[j] * d_out_g;  dG1[g][j] = expert_{sel[g][j]} . d_out_g;
// dTgt = dU[64:84] + dU[148:168]
__global__ void __launch_bounds__(64)
k_sib_mix_bwd(const float* __restrict__ ZE1, const float* __restrict__ ZG1, BnSet e1, BnSet g1, const float* __restrict__ dU,
              float* __restrict__ dE1, float* __restrict__ dG1, float* __restrict__ dTgt, int n_expert, MixSel sel) {
  __shared__ float gt[10];
  __shared__ float red[2][10];
  const int b = blockIdx.x, c = threadIdx.x, lane = c & 31, w = c >> 5;
  const int ldE = n_expert * 64;
  if (c < 10) gt[c] = bn_relu(ZG1[(int64_t)b * 10 + c], g1.stat, g1.gamma, g1.beta, c);
  __syncthreads();
  const float* du = dU + (int64_t)b * 168;
  const float dm = du[c], ds = du[84 + c];
  if (c < kE) dTgt[(int64_t)b * kE + c] = du[64 + c] + du[148 + c];
  for (int e = 0; e < n_expert; ++e) {
    float v = 0.f;
#pragma unroll
    for (int j = 0; j < 5; ++j) {
      if (sel.s[0][j] == e) v = fmaf(gt[j], dm, v);
      if (sel.s[1][j] == e) v = fmaf(gt[5 + j], ds, v);
    }
    dE1[(int64_t)b * ldE + e * 64 + c] = v;
  }
#pragma unroll
  for (int j = 0; j < 5; ++j) {
    const int em = sel.s[0][j] * 64 + c, es = sel.s[1][j] * 64 + c;
    const float pm = warp_sum(bn_relu(ZE1[(int64_t)b * ldE + em], e1.stat, e1.gamma, e1.beta, em) * dm);
    const float ps = warp_sum(bn_relu(ZE1[(int64_t)b * ldE + es], e1.stat, e1.gamma, e1.beta, es) * ds);
    if (lane == 0) { red[w][j] = pm; red[w][5 + j] = ps; }
  }
  __syncthreads();
  if (c < 10) dG1[(int64_t)b * 10 + c] = red[0][c] + red[1][c];
}
void launch_sib_mix_bwd(const float* ZE1, const float* ZG1, const BnSet& e1, const BnSet& g1, const float* dU, float* dE1, float* dG1,
                        float* dTgt, int n_expert, const int sel[2][5], int B, cudaStream_t st) { PAMREC_PROF("sib_mix_bwd", 1, st);
  if (B == 0) return;
  MixSel s;
  for (int g = 0; g < 2; ++g) for (int j = 0; j < 5; ++j) s.s[g][j] = sel[g][j];
  k_sib_mix_bwd<<<B, 64, 0, st>>>(ZE1, ZG1, e1, g1, dU, dE1, dG1, dTgt, n_expert, s);
}

// ------------------------------------------------------------------------------------------
// loss = mean xent(logit, labels) + aux_w * mean xent(valid_logit, labels_play)   (mmoe.py:52-82; aux_w is the literal 0.5)
__device__ __forceinline__ float sib_sigmoid(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float sib_xent(float x, float y) { return fmaxf(x, 0.f) - x * y + log1pf(expf(-fabsf(x))); }
__global__ void __launch_bounds__(1024)
k_sib_loss(const float* __restrict__ logits, const float* __restrict__ y_sat, const float* __restrict__ y_play,
           float* __restrict__ d_logits, double* __restrict__ loss_acc, int B, float aux_w, int heads) {
  __shared__ double sh[2][32];
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const float inv_b = 1.0f / (float)B;
  double a0 = 0.0, a1 = 0.0;
  for (int b = tid; b < B; b += blockDim.x) {
    const float x0 = logits[heads * b], y0 = y_sat[b];
    a0 += (double)sib_xent(x0, y0);
    d_logits[heads * b] = (sib_sigmoid(x0) - y0) * inv_b;
    if (heads == 2) {
      const float x1 = logits[2 * b + 1], y1 = y_play[b];
      a1 += (double)sib_xent(x1, y1);
      d_logits[2 * b + 1] = aux_w * (sib_sigmoid(x1) - y1) * inv_b;
    }
  }
  a0 = warp_sum_d(a0); a1 = warp_sum_d(a1);
  if (lane == 0) { sh[0][w] = a0; sh[1][w] = a1; }
  __syncthreads();
  if (tid == 0) {
    double t0 = 0.0, t1 = 0.0;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) { t0 += sh[0][k]; t1 += sh[1][k]; }
    loss_acc[0] = t0 / (double)B;
    loss_acc[1] = (double)aux_w * t1 / (double)B;
    loss_acc[2] = 0.0;
  }
}
void launch_sib_loss(const float* logits, const float* y_sat, const float* y_play, float* d_logits, double* loss_acc, int B, float aux_w,
                     int heads, cudaStream_t st) { PAMREC_PROF("sib_loss", 1, st);
  k_sib_loss<<<1, 1024, 0, st>>>(logits, y_sat, y_play, d_logits, loss_acc, B, aux_w, heads);
}

__global__ void k_sib_pred(const float* __restrict__ logits, float* __restrict__ pred, int B, int heads) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B) pred[b] = sib_sigmoid(logits[heads * b]);
}
void launch_sib_pred(const float* logits, float* pred, int B, int heads, cudaStream_t st) { PAMREC_PROF("sigmoid", 1, st);
  if (B == 0) return;
  k_sib_pred<<<(B + 255) / 256, 256, 0, st>>>(logits, pred, B, heads);
}

}  // namespace pamrec
