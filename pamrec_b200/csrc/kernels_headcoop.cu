// The head (attention pooling, MMoE, towers, losses) as ONE persistent cooperative kernel per direction.
//
// Round 1 ran the head as 27 dense launches + 11 batch-norm launches per step: 5-8 us kernels below one wave, serialised
// by the batch-norm reductions (every BN layer is a reduction over the whole batch, pamrec.py:366-372, base_model.py:680-686).
// Here a grid of (SMs x kHeadCtasPerSm) CTAs stays resident and walks a device-resident PROGRAM of phases; a phase is a
// list of independent tile items (the same tile bodies as the stand-alone kernels, head_tiles.cuh) that the CTAs pick
// round-robin, and where the next phase needs this one's output - or a batch statistic - a grid-wide barrier separates them.
// The barrier has a LEADER section: CTA 0 waits for every arrival, then
//   * data parallel: all-reduces the fp64 batch-norm sums over the ranks through the NVLink peer mailboxes (the body of
//     kernels_p2p.cu, no extra launch, no NCCL call),
//   * forward / training: turns the sums into (mean, invstd) and updates the moving statistics (k_bn_finalize's arithmetic),
//   * forward / scoring: loads (mean, invstd) from the moving statistics,
// and releases the grid.  Launched with cudaLaunchCooperativeKernel so that co-residency of the grid is guaranteed; every
// spin is bounded (a lost CTA or peer turns into an error word, not a hang).
#include <cstdio>

#include "head_tiles.cuh"
#include "headcoop.h"

namespace pamrec {

// Spin loads are RELAXED (an acquire load makes the compiler invalidate the SM's whole L1 - CCTL.IVALL - after every poll, which
// starved the CTAs of the same SM that were still working: ncu showed 45 % of the samples at the barrier and 10 % in CCTL); the
// acquire happens once, by a fence, after the awaited value has been seen.
__device__ __forceinline__ void st_release_u32(unsigned* p, unsigned v) { asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned ld_relaxed_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
__device__ __forceinline__ void st_flag_sys(uint32_t* p, uint32_t v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t ld_flag_sys_relaxed(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void fence_acq_rel_sys() { asm volatile("fence.acq_rel.sys;" ::: "memory"); }

constexpr int kHeadSmemBytes = kDenseFwdSmem > kDenseDxSmem ? kDenseFwdSmem : kDenseDxSmem;
static_assert(kHeadSmemBytes >= kDenseDwSmem && kHeadSmemBytes >= 4 * PAMREC_MAX_T * 4 + 4 * kD * 4 + 1024, "shared memory union");

// ------------------------------------------------------------------------------------------ item bodies (kHT = 128 threads)
__device__ __forceinline__ float bn_xhat(float z, const float* stat, int col) { return (z - stat[2 * col]) * stat[2 * col + 1]; }

// P1 (pamrec.py:272-282), warp per sample: a = softmax_t(mask ? relu(BN(z2)) : -(2^32)+1)
__device__ __forceinline__ void pool_weights_w(const float* __restrict__ Z2, const BnSet& s1, const int* __restrict__ mask, int64_t base,
                                               int T, int lane, float* aw) {
  const float mean = s1.stat[0], inv = s1.stat[1], g = s1.gamma[0], be = s1.beta[0];
  float s[8];
  float m = -INFINITY;
#pragma unroll
  for (int jj = 0; jj < 8; ++jj) {
    const int t = jj * 32 + lane;
    float val = -INFINITY;
    if (t < T) {
      const float sv = fmaxf(fmaf(g, (Z2[base + t] - mean) * inv, be), 0.f);
      val = (mask[base + t] == 1) ? sv : kMaskNeg;
    }
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

__device__ __noinline__ void pool_fwd_item(const PoolP& p, const BnSet& s1, const int* __restrict__ mask, int B, int T, int item,
                                              unsigned char* smem) {
  float* aws = reinterpret_cast<float*>(smem);              // [4][PAMREC_MAX_T]
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = item * 4 + w;
  if (b >= B) return;                                       // warp-uniform; no block-wide barrier below
  const int64_t base = (int64_t)b * T;
  float* aw = aws + w * PAMREC_MAX_T;
  pool_weights_w(p.Z2, s1, mask, base, T, lane, aw);
  const int tg = lane / 10, c = lane % 10;
  float4 acc = f4_zero();
  if (lane < 30)
    for (int t = tg; t < T; t += 3) f4_fma(acc, aw[t], ld4(p.H + (base + t) * kD + 4 * c));
  float4 a1 = f4_shfl_down(acc, 10), a2 = f4_shfl_down(acc, 20);
  if (lane < 10) {
    acc.x += a1.x + a2.x; acc.y += a1.y + a2.y; acc.z += a1.z + a2.z; acc.w += a1.w + a2.w;
    st4(p.new_long + (int64_t)b * kD + 4 * c, acc);
  }
  __syncwarp();
}

// backward of the pooling + the two column sums of the score layer-1 batch norm (C = 1) of the gradient it produces
__device__ __noinline__ void pool_bwd_item(const PoolP& p, const BnSet& s1, const int* __restrict__ mask, int B, int T, int item,
                                              unsigned char* smem) {
  float* aws = reinterpret_cast<float*>(smem);              // [4][PAMREC_MAX_T]
  float* dnls = aws + 4 * PAMREC_MAX_T;                     // [4][kD]
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = item * 4 + w;
  if (b >= B) return;
  const int64_t base = (int64_t)b * T;
  float* aw = aws + w * PAMREC_MAX_T;
  float* dnl = dnls + w * kD;
  pool_weights_w(p.Z2, s1, mask, base, T, lane, aw);
  for (int i = lane; i < kD; i += 32) dnl[i] = p.dNL[(int64_t)b * kD + i];
  __syncwarp();
  float4 dn[10];
#pragma unroll
  for (int i = 0; i < 10; ++i) dn[i] = ld4(dnl + 4 * i);
  float da[8];
  float dot = 0.f;
#pragma unroll
  for (int jj = 0; jj < 8; ++jj) {
    const int t = jj * 32 + lane;
    float v = 0.f;
    if (t < T) {
      const float* h = p.H + (base + t) * kD;
#pragma unroll
      for (int i = 0; i < 10; ++i) v += f4_dot(dn[i], ld4(h + 4 * i));
      dot = fmaf(aw[t], v, dot);
    }
    da[jj] = v;
  }
  dot = warp_sum(dot);
  const float mean = s1.stat[0], inv = s1.stat[1], g = s1.gamma[0], be = s1.beta[0];
  double b1 = 0.0, b2 = 0.0;
#pragma unroll
  for (int jj = 0; jj < 8; ++jj) {
    const int t = jj * 32 + lane;
    if (t < T) {
      const float d = (mask[base + t] == 1) ? aw[t] * (da[jj] - dot) : 0.f;
      p.dA2[base + t] = d;
      const float xh = (p.Z2[base + t] - mean) * inv;
      const float dy = fmaf(g, xh, be) > 0.f ? d : 0.f;
      b1 += (double)dy;
      b2 += (double)dy * (double)xh;
    }
  }
  b1 = warp_sum_d(b1); b2 = warp_sum_d(b2);
  if (lane == 0) { atomicAdd(s1.bsums, b1); atomicAdd(s1.bsums + 1, b2); }
  const int tg = lane / 10, c = lane % 10;
  if (lane < 30)
    for (int t = tg; t < T; t += 3) {
      const float a = aw[t];
      const float4 d = ld4(dnl + 4 * c);
      st4(p.dH + (base + t) * kD + 4 * c, make_float4(a * d.x, a * d.y, a * d.z, a * d.w));
    }
  __syncwarp();
}

// M1 mixing (pamrec.py:46-50, 315-316): two samples per item, 64 threads each
constexpr int kCombineRows = 8;                              // samples per combine item (4 iterations of 2)
__device__ __noinline__ void combine_fwd_item(const CombineP& p, const BnSet& e1, const BnSet& g1, int B, int item, unsigned char* smem) {
  float* gt = reinterpret_cast<float*>(smem);               // [2][10]
  const int sub = threadIdx.x >> 6, c = threadIdx.x & 63;
  for (int it = 0; it < kCombineRows / 2; ++it) {
    const int b = item * kCombineRows + 2 * it + sub;
    const bool ok = b < B;
    __syncthreads();
    if (ok && c < 10) gt[sub * 10 + c] = bn_relu(p.ZG1[(int64_t)b * 10 + c], g1.stat, g1.gamma, g1.beta, c);
    __syncthreads();
    if (!ok) continue;
    float mn = 0.f, sb = 0.f;
#pragma unroll
    for (int j = 0; j < 5; ++j) {
      const float e = bn_relu(p.ZE1[(int64_t)b * 320 + j * 64 + c], e1.stat, e1.gamma, e1.beta, j * 64 + c);
      mn = fmaf(gt[sub * 10 + j], e, mn);
      sb = fmaf(gt[sub * 10 + 5 + j], e, sb);
    }
    float* u = p.U + (int64_t)b * 168;
    u[c] = mn;
    u[84 + c] = sb;
    if (c < kE) { const float t = p.tgt[(int64_t)b * kE + c]; u[64 + c] = t; u[148 + c] = t; }
  }
}

// backward of the mixing + the column sums of the expert / gate layer-1 batch norms of the gradients it produces
__device__ __noinline__ void combine_bwd_item(const CombineP& p, const BnSet& e1, const BnSet& g1, int B, int item, unsigned char* smem) {
  float* gt = reinterpret_cast<float*>(smem);               // [2][10]
  float* red = gt + 32;                                     // [2 samples][2 warps][10]
  const int sub = threadIdx.x >> 6, c = threadIdx.x & 63, lane = c & 31, w = c >> 5;
  double se1[5] = {0, 0, 0, 0, 0}, se2[5] = {0, 0, 0, 0, 0}, sg1 = 0.0, sg2 = 0.0;
  for (int it = 0; it < kCombineRows / 2; ++it) {
    const int b = item * kCombineRows + 2 * it + sub;
    const bool ok = b < B;
    __syncthreads();
    if (ok && c < 10) gt[sub * 10 + c] = bn_relu(p.ZG1[(int64_t)b * 10 + c], g1.stat, g1.gamma, g1.beta, c);
    __syncthreads();
    float dm = 0.f, ds = 0.f;
    if (ok) {
      const float* du = p.dU + (int64_t)b * 168;
      dm = du[c]; ds = du[84 + c];
      if (c < kE) p.dTgt[(int64_t)b * kE + c] = du[64 + c] + du[148 + c];
    }
#pragma unroll
    for (int j = 0; j < 5; ++j) {
      float e = 0.f;
      if (ok) {
        const int col = j * 64 + c;
        const float xh = bn_xhat(p.ZE1[(int64_t)b * 320 + col], e1.stat, col);
        const float y = fmaf(e1.gamma[col], xh, e1.beta[col]);
        e = fmaxf(y, 0.f);
        const float d = gt[sub * 10 + j] * dm + gt[sub * 10 + 5 + j] * ds;
        p.dE1[(int64_t)b * 320 + col] = d;
        const float dy = y > 0.f ? d : 0.f;
        se1[j] += (double)dy;
        se2[j] += (double)dy * (double)xh;
      }
      const float pm = warp_sum(e * dm), ps = warp_sum(e * ds);
      if (lane == 0) { red[(sub * 2 + w) * 10 + j] = pm; red[(sub * 2 + w) * 10 + 5 + j] = ps; }
    }
    __syncthreads();
    if (ok && c < 10) {
      const float d = red[(sub * 2) * 10 + c] + red[(sub * 2 + 1) * 10 + c];
      p.dG1[(int64_t)b * 10 + c] = d;
      const float xh = bn_xhat(p.ZG1[(int64_t)b * 10 + c], g1.stat, c);
      const float dy = fmaf(g1.gamma[c], xh, g1.beta[c]) > 0.f ? d : 0.f;
      sg1 += (double)dy;
      sg2 += (double)dy * (double)xh;
    }
  }
  // the two halves of the CTA (sub = 0 / 1) hold partial sums of the same columns: merge through shared memory, one atomic each
  __syncthreads();
  double* sh = reinterpret_cast<double*>(smem + 1024);      // [2][330]
  if (sub == 1) {
#pragma unroll
    for (int j = 0; j < 5; ++j) { sh[j * 64 + c] = se1[j]; sh[330 + j * 64 + c] = se2[j]; }
    if (c < 10) { sh[320 + c] = sg1; sh[330 + 320 + c] = sg2; }
  }
  __syncthreads();
  if (sub == 0) {
#pragma unroll
    for (int j = 0; j < 5; ++j) {
      const int col = j * 64 + c;
      atomicAdd(e1.bsums + 2 * col, se1[j] + sh[col]);
      atomicAdd(e1.bsums + 2 * col + 1, se2[j] + sh[330 + col]);
    }
    if (c < 10) { atomicAdd(g1.bsums + 2 * c, sg1 + sh[320 + c]); atomicAdd(g1.bsums + 2 * c + 1, sg2 + sh[330 + 320 + c]); }
  }
}

// L1 + L2 + L3 and their gradients (base_model.py:196-242, pamrec.py:70-106), distributed: one item = kLossUnits listwise
// groups (and as many cross-entropy rows / softmax units).  ApproxNDCG restated from TensorFlow-Ranking 0.3.x
// (oracle/pamrec_oracle.py:approx_ndcg_loss).  loss_acc is zeroed by the caller; items add their shares.
constexpr int kLossUnits = kHT;
__device__ __forceinline__ float sigmoidf2_(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float xent2_(float x, float y) { return fmaxf(x, 0.f) - x * y + log1pf(expf(-fabsf(x))); }
__device__ __forceinline__ double block_sum_d2(double v, double* sh) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  v = warp_sum_d(v);
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  double r = 0.0;
#pragma unroll
  for (int k = 0; k < kHT / 32; ++k) r += sh[k];
  return r;
}
__device__ __noinline__ void loss_item(const LossP& p, const HeadDyn& d, int item, unsigned char* smem) {
  double* sh = reinterpret_cast<double*>(smem);
  const int tid = threadIdx.x, B = d.B;
  const int G = B / PAMREC_GROUP;
  const float* logits = p.logits;
  float* d_logits = p.d_logits;
  double nval;
  if (d.world > 1) {
    nval = *p.n_valid_global;                                // count over all ranks (all-reduced in the forward pass)
  } else {
    double cnt = 0.0;
    for (int g = tid; g < G; g += kHT) {
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < PAMREC_GROUP; ++i) s += d.plays[g * PAMREC_GROUP + i];
      cnt += (s > 0.f) ? 1.0 : 0.0;
    }
    nval = block_sum_d2(cnt, sh);
  }
  const float inv_b = 1.0f / (float)d.Bg;
  double a0 = 0.0, a1 = 0.0, a2 = 0.0;
  if (d.sm_group == 0) {
    for (int k = 0; k < PAMREC_GROUP; ++k) {
      const int b = (item * kLossUnits) * PAMREC_GROUP + k * kLossUnits + tid;      // the item's 5 * kLossUnits rows, coalesced
      if (b < B) {
        const float x0 = logits[3 * b], x1 = logits[3 * b + 1];
        const float y0 = d.y_sat[b], y1 = d.y_play[b];
        a0 += (double)xent2_(x0, y0);
        a1 += (double)xent2_(x1, y1);
        d_logits[3 * b] = (sigmoidf2_(x0) - y0) * inv_b;
        d_logits[3 * b + 1] = d.fuzhu_w * (sigmoidf2_(x1) - y1) * inv_b;
      }
    }
  } else {
    // hparams.loss == "softmax":  -group * mean(log(where(y == 1, softmax, 1))) over all B elements
    const int sm = d.sm_group;
    const float scale = (float)sm * inv_b;
    const int u = item * kLossUnits + tid;
    // units are (softmax group, head) pairs; an item covers kLossUnits of them
    if (u < 2 * (B / sm)) {
      const int head = u & 1, r0 = (u >> 1) * sm;
      const float* y = head ? d.y_play : d.y_sat;
      float mx = -INFINITY;
      for (int i = 0; i < sm; ++i) mx = fmaxf(mx, logits[3 * (r0 + i) + head]);
      float se = 0.f;
      int n_pos = 0;
      for (int i = 0; i < sm; ++i) { se += expf(logits[3 * (r0 + i) + head] - mx); n_pos += y[r0 + i] == 1.0f; }
      const float lse = mx + logf(se), wgt = head ? d.fuzhu_w : 1.0f;
      double acc = 0.0;
      for (int i = 0; i < sm; ++i) {
        const float x = logits[3 * (r0 + i) + head];
        const bool pos = y[r0 + i] == 1.0f;
        if (pos) acc += (double)(lse - x);
        d_logits[3 * (r0 + i) + head] = wgt * scale * ((float)n_pos * expf(x - lse) - (pos ? 1.0f : 0.f));
      }
      if (head) a1 += acc * (double)sm; else a0 += acc * (double)sm;
    }
  }
  const float alpha = 10.0f;
  const int g = item * kLossUnits + tid;
  if (g < G) {
    float o[5], s[5], y[5], gain[5], rank[5], dLr[5];
    float lsum = 0.f;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      o[i] = logits[3 * (g * 5 + i) + 2];
      s[i] = sigmoidf2_(o[i]);                      // pamrec.py:74
      y[i] = d.plays[g * 5 + i];
      lsum += y[i];
    }
    const bool valid = lsum > 0.f;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      const float yy = valid ? y[i] : 1e-10f;
      y[i] = yy;
      gain[i] = exp2f(yy) - 1.0f;
    }
    float dcg = 0.f;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      float r = 0.5f;
#pragma unroll
      for (int j = 0; j < 5; ++j) r += sigmoidf2_(alpha * (s[j] - s[i]));
      rank[i] = r;
      dcg += gain[i] / log1pf(r);
    }
    float ys[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) ys[i] = y[i];
#pragma unroll
    for (int i = 0; i < 4; ++i)                    // sort descending (5 elements)
#pragma unroll
      for (int j = 0; j < 4 - i; ++j)
        if (ys[j] < ys[j + 1]) { const float tmp = ys[j]; ys[j] = ys[j + 1]; ys[j + 1] = tmp; }
    float idcg = 0.f;
#pragma unroll
    for (int r = 0; r < 5; ++r) idcg += (exp2f(ys[r]) - 1.0f) / log1pf((float)(r + 1));
    const float inv = idcg > 0.f ? 1.0f / idcg : 0.f;
    const float w = valid ? 1.0f : 0.f;
    a2 += (double)(w * -(dcg * inv));
    const float coef = (nval > 0.0) ? d.order_w * w / (float)nval : 0.f;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      const float l1p = log1pf(rank[i]);
      dLr[i] = gain[i] * inv / (l1p * l1p * (1.0f + rank[i]));
    }
#pragma unroll
    for (int j = 0; j < 5; ++j) {
      float acc = 0.f;
#pragma unroll
      for (int i = 0; i < 5; ++i) {
        if (i == j) continue;
        const float sij = sigmoidf2_(alpha * (s[j] - s[i]));   // d rank_i / d s_j
        const float sji = sigmoidf2_(alpha * (s[i] - s[j]));   // d rank_j / d s_j (negative sign)
        acc += dLr[i] * alpha * sij * (1.0f - sij) - dLr[j] * alpha * sji * (1.0f - sji);
      }
      d_logits[3 * (g * 5 + j) + 2] = coef * acc * s[j] * (1.0f - s[j]);
    }
  }
  if (item == 0)
    for (int b = G * 5 + tid; b < B; b += kHT) d_logits[3 * b + 2] = 0.f;
  a0 = block_sum_d2(a0, sh);
  a1 = block_sum_d2(a1, sh);
  a2 = block_sum_d2(a2, sh);
  if (tid == 0) {
    if (a0 != 0.0) atomicAdd(p.loss_acc, a0 / (double)d.Bg);
    if (a1 != 0.0) atomicAdd(p.loss_acc + 1, (double)d.fuzhu_w * a1 / (double)d.Bg);
    if (a2 != 0.0 && nval > 0.0) atomicAdd(p.loss_acc + 2, (double)d.order_w * a2 / nval);
  }
}

// ------------------------------------------------------------------------------------------ leader work
// all-reduce of fp64 vectors over the ranks through the peer mailboxes (same protocol as k_p2p_allreduce, kernels_p2p.cu)
__device__ __forceinline__ void leader_p2p(const HeadDyn& d, int slot, double* const* buf, const int* n, int nbuf) {
  const int tid = threadIdx.x;
  int total = 0;
  for (int b = 0; b < nbuf; ++b) total += n[b];
  for (int i = tid; i < total; i += kHT) {
    int b = 0, o = i;
    while (o >= n[b]) { o -= n[b]; ++b; }
    const double v = buf[b][o];
    for (int p = 0; p < d.world; ++p) d.peer_slots[p][(size_t)(slot * d.world + d.rank) * kP2PMaxDoubles + i] = v;
  }
  __threadfence_system();
  __syncthreads();
  if (tid < d.world) st_flag_sys(d.peer_flags[tid] + slot * d.world + d.rank, d.p2p_epoch);
  if (tid < d.world) {
    const uint32_t* f = d.peer_flags[d.rank] + slot * d.world + tid;
    uint32_t spins = 0;
    while ((int32_t)(ld_flag_sys_relaxed(f) - d.p2p_epoch) < 0) {
      if (++spins > (1u << 24)) { *d.p2p_err = 1u + (uint32_t)slot; break; }
      __nanosleep(64);
    }
    fence_acq_rel_sys();
  }
  __syncthreads();
  const double* mine = d.peer_slots[d.rank] + (size_t)slot * d.world * kP2PMaxDoubles;
  for (int i = tid; i < total; i += kHT) {
    double s = 0.0;
    for (int p = 0; p < d.world; ++p) s += __ldcg(mine + (size_t)p * kP2PMaxDoubles + i);
    int b = 0, o = i;
    while (o >= n[b]) { o -= n[b]; ++b; }
    buf[b][o] = s;
  }
  __syncthreads();
}

__device__ __noinline__ void leader_work(const HeadProgram* prog, const HeadPhase& ph, const HeadDyn& d, int barrier_index) {
  const int tid = threadIdx.x;
  if (d.world > 1 && d.training && (ph.n_sync > 0 || ph.sync_scalars)) {
    double* buf[3]; int n[3]; int nb = 0;
    for (int k = 0; k < ph.n_sync; ++k) {
      const BnSet& s = prog->bn[ph.sync[k]];
      buf[nb] = ph.sync_bwd ? s.bsums : s.sums; n[nb++] = 2 * s.C;
    }
    if (ph.sync_scalars) { buf[nb] = prog->dp_scalars; n[nb++] = 8; }
    leader_p2p(d, d.p2p_slot0 + barrier_index, buf, n, nb);
  }
  if (d.training) {
    for (int k = 0; k < ph.n_fin; ++k) {
      const BnSet& s = prog->bn[ph.fin[k]];
      const double count = ph.fin_rows_n ? d.cntN : d.cntB;
      for (int c = tid; c < s.C; c += kHT) {                 // k_bn_finalize's arithmetic
        const double mean = s.sums[2 * c] / count;
        double var = s.sums[2 * c + 1] / count - mean * mean;
        if (var < 0.0) var = 0.0;
        s.stat[2 * c] = (float)mean;
        s.stat[2 * c + 1] = (float)(1.0 / sqrt(var + (double)kBnEps));
        s.mmean[c] -= (s.mmean[c] - (float)mean) * kBnDecay;
        s.mvar[c] -= (s.mvar[c] - (float)var) * kBnDecay;
        s.sums[2 * c] = 0.0;
        s.sums[2 * c + 1] = 0.0;
      }
    }
  } else if (ph.eval_stats) {
    for (int k = 0; k < BN_COUNT_; ++k) {
      const BnSet& s = prog->bn[k];
      for (int c = tid; c < s.C; c += kHT) { s.stat[2 * c] = s.mmean[c]; s.stat[2 * c + 1] = 1.0f / sqrtf(s.mvar[c] + kBnEps); }
    }
  }
}

// grid-wide barrier with a leader section.  bar[0] = arrivals (reset by the leader), bar[32] = release epoch (monotonic across
// launches, on its own 128-byte line so that polls do not queue behind the arrival atomics), bar[2] = error word.  `epoch` is this CTA's copy of the release epoch.
__device__ __forceinline__ void grid_barrier(unsigned* bar, unsigned& epoch, const HeadProgram* prog, const HeadPhase& ph,
                                             const HeadDyn& d, int barrier_index) {
  __syncthreads();
  if (blockIdx.x == 0) {
    if (threadIdx.x == 0) {
      unsigned spins = 0;
      while (ld_relaxed_u32(bar) < gridDim.x - 1) {
        if (++spins > (1u << 24)) { bar[2] = 1u + (unsigned)barrier_index; break; }
        __nanosleep(64);
      }
      fence_acq_rel_gpu();                                  // every arrival's release (and what its CTA wrote) is visible from here
    }
    __syncthreads();
    leader_work(prog, ph, d, barrier_index);
    __syncthreads();
    if (threadIdx.x == 0) {
      if (d.trace != nullptr && barrier_index < 29) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); d.trace[barrier_index] = t; }
      bar[0] = 0u;
      __threadfence();                                      // cumulative: orders the whole CTA's writes (seen through the CTA barrier)
      st_release_u32(bar + 32, epoch + 1u);
    }
  } else {
    if (threadIdx.x == 0) {
      __threadfence();                                      // this CTA's writes before its arrival
      atomicAdd(bar, 1u);
      unsigned spins = 0;
      while ((int)(ld_relaxed_u32(bar + 32) - (epoch + 1u)) < 0) {
        if (++spins > (1u << 24)) { bar[2] = 1000u + (unsigned)barrier_index; break; }
        __nanosleep(64);
      }
      fence_acq_rel_gpu();
    }
    __syncthreads();
  }
  epoch += 1u;
}

// ------------------------------------------------------------------------------------------ the persistent kernel
__global__ void __launch_bounds__(kHT, kHeadCtasPerSm) k_head_program(const HeadProgram* __restrict__ prog, const HeadDyn d) {
  __shared__ __align__(16) unsigned char smem[kHeadSmemBytes];
  unsigned* bar = d.bar;
  __shared__ unsigned s_epoch;
  if (threadIdx.x == 0 && blockIdx.x == 0 && d.trace != nullptr) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); d.trace[31] = t; }
  if (threadIdx.x == 0) s_epoch = ld_relaxed_u32(bar + 32);   // the previous launch on this stream has completed: every CTA reads the same value
  __syncthreads();
  unsigned epoch = s_epoch;
  const int n_phase = prog->n;
  const int N = d.B * d.T;
  int n_barrier = 0;
  const int cap = (int)gridDim.x;                            // items per phase ~ one per resident CTA (one wave)
  // Phase descriptors are read from SHARED memory: every field access from the device-resident program was an L2 round trip
  // in a chain of dependent loads (ncu: 30 % of the samples on the long scoreboard).  The next descriptor is prefetched while
  // the current phase runs.
  static_assert(sizeof(HeadPhase) % 16 == 0, "HeadPhase is copied as uint4");
  __shared__ __align__(16) unsigned char s_phase[2][sizeof(HeadPhase)];
  auto fetch = [&](int k, int buf) {
    const uint4* src = reinterpret_cast<const uint4*>(&prog->ph[k]);
    uint4* dst = reinterpret_cast<uint4*>(s_phase[buf]);
    for (int i = threadIdx.x; i < (int)(sizeof(HeadPhase) / 16); i += kHT) dst[i] = __ldg(src + i);
  };
  fetch(0, 0);
  for (int k = 0; k < n_phase; ++k) {
    __syncthreads();                                         // descriptor k is complete; the other buffer (phase k - 1) is free
    if (k + 1 < n_phase) fetch(k + 1, (k + 1) & 1);
    const HeadPhase& ph = *reinterpret_cast<const HeadPhase*>(s_phase[k & 1]);
    const bool skip = (ph.only == 1 && !d.training) || (ph.only == 2 && d.training);
    if (!skip) {
      const int M = ph.rows_n ? N : d.B;
      // rotate the first item over the CTAs from phase to phase: phases without a barrier in between then fill different CTAs
      const int rot = (int)((blockIdx.x + (unsigned)k * 37u) % gridDim.x);
      switch (ph.op) {
        case HEAD_OP_DENSE_FWD: {
          const DenseP p = ph.u.f;                           // small: a private copy keeps its fields in registers
          const FwdGrid g = dense_fwd_grid(M, p.n_groups, p.N, cap);
          double* sums = d.training ? p.out_sums : nullptr;
          for (int i = rot; i < g.gx * g.ny; i += gridDim.x) { __syncthreads(); dense_fwd_item(p, M, sums, g.tiles_m, i % g.gx, g.gx, i / g.gx, smem); }
        } break;
        case HEAD_OP_DENSE_DX: {
          const DenseDxP& p = ph.u.x;
          const DxGrid g = dense_dx_grid(M, p.n_slices, p.K);
          const double cnt = ph.rows_n ? d.cntN : d.cntB;
          for (int i = rot; i < g.gx * g.ny; i += gridDim.x) { __syncthreads(); dense_dx_item(p, M, cnt, i % g.gx, i / g.gx, smem); }
        } break;
        case HEAD_OP_DENSE_DW: {
          const DenseDwP& p = ph.u.w;
          const DwGrid g = dense_dw_grid(M, p.n_groups, p.K, p.N, cap);
          const double cnt = ph.rows_n ? d.cntN : d.cntB;
          for (int i = rot; i < g.chunks * g.ny; i += gridDim.x) { __syncthreads(); dense_dw_item(p, M, cnt, g.rows_per_cta, i % g.chunks, i / g.chunks, smem); }
        } break;
        case HEAD_OP_POOL_FWD:
          for (int i = rot; i < (d.B + 3) / 4; i += gridDim.x) { __syncthreads(); pool_fwd_item(ph.u.pl, prog->bn[BN_S1_], d.mask, d.B, d.T, i, smem); }
          break;
        case HEAD_OP_POOL_BWD:
          for (int i = rot; i < (d.B + 3) / 4; i += gridDim.x) { __syncthreads(); pool_bwd_item(ph.u.pl, prog->bn[BN_S1_], d.mask, d.B, d.T, i, smem); }
          break;
        case HEAD_OP_COMBINE_FWD:
          for (int i = rot; i < (d.B + kCombineRows - 1) / kCombineRows; i += gridDim.x) { __syncthreads(); combine_fwd_item(ph.u.cb, prog->bn[BN_E1_], prog->bn[BN_G1_], d.B, i, smem); }
          break;
        case HEAD_OP_COMBINE_BWD:
          for (int i = rot; i < (d.B + kCombineRows - 1) / kCombineRows; i += gridDim.x) { __syncthreads(); combine_bwd_item(ph.u.cb, prog->bn[BN_E1_], prog->bn[BN_G1_], d.B, i, smem); }
          break;
        case HEAD_OP_LOSS: {
          const int G = d.B / PAMREC_GROUP;
          int units = G;
          if (d.sm_group > 0) { const int u2 = 2 * (d.B / d.sm_group); units = u2 > G ? u2 : G; }
          int n_items = (units + kLossUnits - 1) / kLossUnits;
          if (n_items < 1) n_items = 1;
          for (int i = rot; i < n_items; i += gridDim.x) { __syncthreads(); loss_item(ph.u.ls, d, i, smem); }
        } break;
        default: break;
      }
    }
    if (ph.barrier && !((ph.only == 1 && !d.training) || (ph.only == 2 && d.training))) {
      grid_barrier(bar, epoch, prog, ph, d, n_barrier);
      ++n_barrier;
    }
  }
  if (threadIdx.x == 0 && blockIdx.x == 0 && d.trace != nullptr) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); d.trace[30] = t; d.trace[29] = (unsigned long long)n_barrier; }
}

// ------------------------------------------------------------------------------------------ host side
int head_program_grid(int* ctas_per_sm_out) {
  int dev = 0, sms = 0, occ = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_head_program, kHT, 0) != cudaSuccess || occ < 1) return -1;
  int coop = 0;
  cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
  if (!coop) return -1;
  if (occ > kHeadCtasPerSm) occ = kHeadCtasPerSm;
  if (ctas_per_sm_out) *ctas_per_sm_out = occ;
  return sms * occ;
}

int launch_head_program(const HeadProgram* dev_prog, const HeadDyn& d, int grid, const char* name, cudaStream_t st) {
  PAMREC_PROF(name, 1, st);
  if (d.B == 0 && d.world == 1) return 0;
  void* args[2] = {(void*)&dev_prog, (void*)&d};
  cudaError_t e = cudaLaunchCooperativeKernel((const void*)k_head_program, dim3(grid), dim3(kHT), args, 0, st);
  return e == cudaSuccess ? 0 : -1;
}

}  // namespace pamrec
