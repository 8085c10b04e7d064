// Shared device helpers and the host-side model description.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/pamrec_b200.h"

namespace pamrec {

constexpr int kI = PAMREC_ITEM_DIM;   // 16
constexpr int kC = PAMREC_CATE_DIM;   // 4
constexpr int kE = PAMREC_EMB_DIM;    // 20
constexpr int kD = PAMREC_D;          // 40
constexpr int kNB = PAMREC_NBUCKET;   // 10
constexpr int kDD = kD * kD;          // 1600
constexpr int kMT = 1;                // 16-row MMA tiles per warp in the token-tile kernels
constexpr int kWarpRows = 16 * kMT;   // token rows owned by one warp
constexpr int kTokThreads = 128;      // threads per CTA of the token-tile kernels (4 warps)
constexpr int kTokTile = 4 * kWarpRows;  // tokens per CTA tile (64: three CTAs per SM, ~1.8 waves at the bench workload)
constexpr int kRowPad = 41;           // smem row stride (floats) for thread==token tiles: conflict-free column walks
constexpr int kAttnStride = 44;       // smem row stride for float4 row reads in the attention kernels
constexpr float kMaskNeg = -4294967295.0f;  // -(2**32)+1, pamrec.py:276,780
constexpr float kLnEps = 1e-8f;             // pamrec.py:639
constexpr float kBnEps = 1e-4f;             // pamrec.py:370
constexpr float kBnDecay = 0.05f;           // 1 - momentum(0.95), pamrec.py:369

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, const float4& v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ float4 f4_zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ void f4_fma(float4& a, float s, const float4& b) {
  a.x = fmaf(s, b.x, a.x); a.y = fmaf(s, b.y, a.y); a.z = fmaf(s, b.z, a.z); a.w = fmaf(s, b.w, a.w);
}
__device__ __forceinline__ float f4_dot(const float4& a, const float4& b) {
  return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, a.w * b.w)));
}
__device__ __forceinline__ float4 f4_shfl_down(const float4& v, int d) {
  float4 r;
  r.x = __shfl_down_sync(0xffffffffu, v.x, d); r.y = __shfl_down_sync(0xffffffffu, v.y, d);
  r.z = __shfl_down_sync(0xffffffffu, v.z, d); r.w = __shfl_down_sync(0xffffffffu, v.w, d);
  return r;
}

// One batch-normalisation "set": a run of channels that is normalised by one kernel pass.
struct BnSet {
  int C;
  const float* gamma; const float* beta;  // dense params
  float* dgamma; float* dbeta;            // dense grads
  float* mmean; float* mvar;              // moving stats (bn pool)
  double* sums;                           // [C][2] forward  sum z, sum z^2
  float* stat;                            // [C][2] mean, invstd
  double* bsums;                          // [C][2] backward sum dy, sum dy*xhat
};

// Grouped dense layer: Z[m, z_off[g]+n] = sum_k act(X[m, x_off[g]+k]) * W_g[k][n] + b_g[n]
struct DenseP {
  const float* X; int ldx; int M;
  int n_groups, K, N;
  int x_off[8], z_off[8];
  const float* W; int w_stride;
  const float* bias; int b_stride;
  float* Z; int ldz;
  const float* in_stat; const float* in_gamma; const float* in_beta;  // BN+ReLU on load (indexed by input column) or null
  double* out_sums;                                                   // [.][2] by output column, or null
};

// dX[m, out_off+k] (+)= sum over contributions c: sum_n dZ[m, dz_off[c]+n] * W_c[k][n]
// Batch-norm backward folded into the consumers of a gradient buffer.  The buffer holds dA, the gradient wrt the
// layer's post-BN-ReLU activation, laid out exactly like the layer's pre-BN output Z; the gradient wrt Z,
//   dz = gamma * invstd * (dy - S1/n - xhat * S2/n),  dy = relu'(gamma * xhat + beta) * dA,
// is evaluated while a consumer loads its operand (S1 = sum dy, S2 = sum dy * xhat: bsums, complete before the launch).
struct BnGrad {
  const float* Z;            // null: the buffer already holds dz
  const float* stat; const float* gamma; const float* beta;
  const double* bsums; double count;
};
// Column sums S1, S2 of a gradient buffer that a kernel PRODUCES (the batch norm of the previous layer), or Z == null
struct BnGradOut {
  const float* Z; const float* stat; const float* gamma; const float* beta;
  double* bsums;
};
struct DenseDxP {
  const float* dZ; int lddz; int M;
  int n_slices; int K;
  int out_off[8]; int n_contrib[8];
  int dz_off[8][8]; int64_t w_off[8][8]; int Ncon[8][8];
  const float* Wbase;
  float* dX; int lddx; int accumulate;
  BnGrad g;                  // of dZ
  BnGradOut o;               // of dX (same layout as o.Z, leading dimension lddx)
};

// dW_g[k][n] += sum_m act(X[m, x_off[g]+k]) * dZ[m, z_off[g]+n];  db_g[n] += sum_m dZ[m, z_off[g]+n]
struct DenseDwP {
  const float* X; int ldx; int M;
  int n_groups, K, N;
  int x_off[8], z_off[8];
  const float* dZ; int lddz;
  const float* in_stat; const float* in_gamma; const float* in_beta;
  float* dW; int w_stride; float* db; int b_stride;
  BnGrad g;                  // of dZ
  // the CTA (0, 0) of this launch also emits the BN parameter gradients of dZ's layer: dbeta += S1 * scale, dgamma += S2 * scale
  float* g_dgamma; float* g_dbeta; int g_C; float g_scale;
};

}  // namespace pamrec
