// Encoder kernels: embedding gather, bucket planning, time-aware Q/K/V projection,
// key-masked attention, LayerNorm + point-wise FFN — forward and backward.
//
// Mapping used by every token-parallel kernel: ONE THREAD == ONE TOKEN.  A CTA stages a
// tile of 128 token rows (40 floats each) in shared memory with row stride 41 (so the 32
// lanes of a warp walking the same column hit 32 different banks), keeps the 40 output
// accumulators of a row in registers, and reads the 40x40 weight matrix from shared memory
// with warp-uniform float4 loads (one broadcast wavefront feeds 128 FMAs).  Tokens are
// bucket-sorted first so that a CTA needs exactly one (Wq,Wk,Wv)[bucket] triple:
// the reference instead materialises a [B,T,40,40] gather per matrix (pamrec.py:714-728).
#include "kernels.h"

namespace pamrec {

// ------------------------------------------------------------------------------------------
// G1+G2+G3+X1 (sequential_base_model.py:603-616,666-668; pamrec.py:155-159,251-257):
// x0[b,t,:] = item[ih[b,t]] | cate[ch[b,t]] | item[items[b]] | cate[cates[b]]  +  pos[t]
// One thread per 16-byte chunk, kEmbU tokens per thread.  A CTA of 320 threads covers 32 x kEmbU tokens: phase 1
// stages the four ids and the position of every token in shared memory (coalesced 4-byte loads), phase 2 issues all
// kEmbU row loads of a thread back to back (they are independent, so kEmbU x 16 B per thread are in flight while
// HBM answers: the kernel is bound by bytes in flight, not by instruction issue), phase 3 adds the L1-resident
// position row and writes; a warp's stores are contiguous 512 B.
constexpr int kEmbU = 8;
constexpr int kEmbTok = 32;
__global__ void __launch_bounds__(kEmbTok * 10, 3)
k_embed_fwd(const int* __restrict__ ih, const int* __restrict__ ch, const int* __restrict__ items,
            const int* __restrict__ cates, const float* __restrict__ item_w, const float* __restrict__ cate_w,
            const float* __restrict__ pos, float* __restrict__ x0, float* __restrict__ tgt, int64_t n_tok, int T) {
  constexpr int TOK = kEmbTok * kEmbU;
  __shared__ int s_id[4][TOK];      // history item, history cate, target item, target cate
  __shared__ int s_t[TOK];
  const int tid = threadIdx.x;
  const int64_t tok0 = (int64_t)blockIdx.x * TOK;
  for (int i = tid; i < TOK; i += kEmbTok * 10) {
    const int64_t tok = tok0 + i;
    if (tok < n_tok) {
      const int64_t b = tok / T;
      s_t[i] = (int)(tok - b * T);
      s_id[0][i] = __ldg(ih + tok);
      s_id[1][i] = __ldg(ch + tok);
      s_id[2][i] = __ldg(items + b);
      s_id[3][i] = __ldg(cates + b);
    }
  }
  __syncthreads();
  const int c = tid % 10, tl = tid / 10;
  const int which = c < 4 ? 0 : (c == 4 ? 1 : (c < 9 ? 2 : 3));
  const bool is_item = (which & 1) == 0;
  const int sub = c < 4 ? c : (c < 9 && c > 4 ? c - 5 : 0);
  const float* table = is_item ? item_w : cate_w;
  const int width = is_item ? kI : kC;
  float4 v[kEmbU];
#pragma unroll
  for (int j = 0; j < kEmbU; ++j) {
    const int i = j * kEmbTok + tl;
    if (tok0 + i < n_tok) v[j] = __ldg(reinterpret_cast<const float4*>(table + (int64_t)s_id[which][i] * width) + sub);
  }
#pragma unroll
  for (int j = 0; j < kEmbU; ++j) {
    const int i = j * kEmbTok + tl;
    const int64_t tok = tok0 + i;
    if (tok < n_tok) {
      const int t = s_t[i];
      if (tgt != nullptr && t == 0 && c >= 5) st4(tgt + (tok / T) * kE + 4 * (c - 5), v[j]);
      const float4 p = __ldg(reinterpret_cast<const float4*>(pos + t * kD) + c);
      float4 o = v[j];
      o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w;
      st4(x0 + tok * kD + 4 * c, o);
    }
  }
}

void launch_embed_fwd(const int* ih, const int* ch, const int* items, const int* cates, const float* item_w,
                      const float* cate_w, const float* pos, float* x0, float* tgt, int64_t n_rows, int T,
                      cudaStream_t st) {
  PAMREC_PROF("embed_fwd", 1, st);
  const int64_t n_tok = n_rows * T;
  if (n_tok == 0) return;
  const int64_t per_cta = (int64_t)kEmbU * kEmbTok;
  k_embed_fwd<<<(unsigned)((n_tok + per_cta - 1) / per_cta), kEmbTok * 10, 0, st>>>(ih, ch, items, cates, item_w, cate_w, pos, x0,
                                                                                  tgt, n_tok, T);
}

// ------------------------------------------------------------------------------------------
// Bucket planning: counting sort of tokens by play-ratio bucket (10 bins) and a tile table.
// ctl: [0..15] counts, [16..31] cursors, [32] number of tiles.
__global__ void k_bucket_hist(const float* __restrict__ lt, int n, int* __restrict__ bucket, int* __restrict__ ctl) {
  __shared__ int h[16];
  if (threadIdx.x < 16) h[threadIdx.x] = 0;
  __syncthreads();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    int k = (int)lt[i];                 // tf.cast(float -> int32) truncates, pamrec.py:716
    k = min(max(k, 0), kNB - 1);
    bucket[i] = k;
    atomicAdd(&h[k], 1);
  }
  __syncthreads();
  if (threadIdx.x < kNB && h[threadIdx.x]) atomicAdd(&ctl[threadIdx.x], h[threadIdx.x]);
}

__global__ void k_bucket_plan(int* __restrict__ ctl, int* __restrict__ tile_bucket, int* __restrict__ tile_begin,
                              int* __restrict__ tile_count) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  int start = 0, nt = 0;
  for (int k = 0; k < kNB; ++k) {
    int cnt = ctl[k];
    ctl[16 + k] = start;
    for (int o = 0; o < cnt; o += kTokTile) {
      tile_bucket[nt] = k;
      tile_begin[nt] = start + o;
      tile_count[nt] = min(kTokTile, cnt - o);
      ++nt;
    }
    start += cnt;
  }
  ctl[32] = nt;
}

__global__ void k_bucket_scatter(const int* __restrict__ bucket, int n, int* __restrict__ ctl, int* __restrict__ perm) {
  __shared__ int h[16], base[16];
  if (threadIdx.x < 16) h[threadIdx.x] = 0;
  __syncthreads();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int k = -1, r = 0;
  if (i < n) { k = bucket[i]; r = atomicAdd(&h[k], 1); }
  __syncthreads();
  if (threadIdx.x < kNB) base[threadIdx.x] = h[threadIdx.x] ? atomicAdd(&ctl[16 + threadIdx.x], h[threadIdx.x]) : 0;
  __syncthreads();
  if (i < n) perm[base[k] + r] = i;
}

void launch_bucket_plan(const float* lt, int n, int* bucket, int* perm, int* ctl, int* tile_bucket, int* tile_begin,
                        int* tile_count, cudaStream_t st) {
  PAMREC_PROF("bucket_plan", 3, st);
  cudaMemsetAsync(ctl, 0, 64 * sizeof(int), st);
  if (n == 0) return;
  int g = (n + 255) / 256;
  k_bucket_hist<<<g, 256, 0, st>>>(lt, n, bucket, ctl);
  k_bucket_plan<<<1, 32, 0, st>>>(ctl, tile_bucket, tile_begin, tile_count);
  k_bucket_scatter<<<g, 256, 0, st>>>(bucket, n, ctl, perm);
}

// ------------------------------------------------------------------------------------------
// register-tile helpers (thread == token)
// acc[j] += sum_i in[i] * W[i][j]
__device__ __forceinline__ void mv_fwd(const float* __restrict__ in, const float* __restrict__ W, float (&acc)[kD]) {
#pragma unroll 4
  for (int i = 0; i < kD; ++i) {
    float a = in[i];
#pragma unroll
    for (int j = 0; j < kD / 4; ++j) {
      float4 w = ld4(W + i * kD + 4 * j);
      acc[4 * j + 0] = fmaf(a, w.x, acc[4 * j + 0]);
      acc[4 * j + 1] = fmaf(a, w.y, acc[4 * j + 1]);
      acc[4 * j + 2] = fmaf(a, w.z, acc[4 * j + 2]);
      acc[4 * j + 3] = fmaf(a, w.w, acc[4 * j + 3]);
    }
  }
}
// acc[i] += sum_j g[j] * W[i][j]
__device__ __forceinline__ void mv_bwd(const float* __restrict__ g, const float* __restrict__ W, float (&acc)[kD]) {
#pragma unroll 2
  for (int j = 0; j < kD / 4; ++j) {
    float4 gv = make_float4(g[4 * j], g[4 * j + 1], g[4 * j + 2], g[4 * j + 3]);
#pragma unroll
    for (int i = 0; i < kD; ++i) acc[i] += f4_dot(gv, ld4(W + i * kD + 4 * j));
  }
}
// dW[40][40] += A^T G over `cnt` rows of two stride-41 tiles; 100 threads own a 2x8 patch each.
__device__ __forceinline__ void tile_outer_atomic(const float* __restrict__ A, const float* __restrict__ G, int cnt,
                                                  float* __restrict__ dW, int tid) {
  if (tid >= 100) return;
  int i0 = 2 * (tid / 5), j0 = 8 * (tid % 5);
  float a0[8], a1[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) a0[q] = a1[q] = 0.f;
  for (int r = 0; r < cnt; ++r) {
    float x0 = A[r * kRowPad + i0], x1 = A[r * kRowPad + i0 + 1];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      float g = G[r * kRowPad + j0 + q];
      a0[q] = fmaf(x0, g, a0[q]);
      a1[q] = fmaf(x1, g, a1[q]);
    }
  }
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    atomicAdd(dW + i0 * kD + j0 + q, a0[q]);
    atomicAdd(dW + (i0 + 1) * kD + j0 + q, a1[q]);
  }
}
__device__ __forceinline__ void tile_colsum_atomic(const float* __restrict__ G, int cnt, float* __restrict__ db, int tid) {
  if (tid >= kD) return;
  float s = 0.f;
  for (int r = 0; r < cnt; ++r) s += G[r * kRowPad + tid];
  atomicAdd(db + tid, s);
}
// cooperative load of `cnt` token rows (global, 40 floats each) into a stride-41 tile
__device__ __forceinline__ void load_tile_perm(float* __restrict__ dst, const float* __restrict__ src, const int* toks,
                                               int cnt, int tid) {
  for (int i = tid; i < cnt * 10; i += kTokTile) {
    int r = i / 10, c = i % 10;
    float4 v = ld4(src + (int64_t)toks[r] * kD + 4 * c);
    float* d = dst + r * kRowPad + 4 * c;
    d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
  }
}
__device__ __forceinline__ void load_tile_lin(float* __restrict__ dst, const float* __restrict__ src, int cnt, int tid) {
  for (int i = tid; i < cnt * 10; i += kTokTile) {
    int r = i / 10, c = i % 10;
    float4 v = ld4(src + (int64_t)r * kD + 4 * c);
    float* d = dst + r * kRowPad + 4 * c;
    d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
  }
}
__device__ __forceinline__ void store_row(float* __restrict__ dst, const float (&acc)[kD]) {
#pragma unroll
  for (int j = 0; j < kD / 4; ++j) st4(dst + 4 * j, make_float4(acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]));
}
// population mean / rstd of a 40-wide smem row (pamrec.py:659-662)
__device__ __forceinline__ void row_stats(const float* __restrict__ x, float& mean, float& rstd) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kD; ++i) s += x[i];
  mean = s * (1.0f / kD);
  float v = 0.f;
#pragma unroll
  for (int i = 0; i < kD; ++i) { float d = x[i] - mean; v = fmaf(d, d, v); }
  rstd = 1.0f / sqrtf(v * (1.0f / kD) + kLnEps);
}

// ------------------------------------------------------------------------------------------
// N1 + A1 forward: qin = LN_a(x); Q = qin Wq[k], K = x Wk[k], V = x Wv[k]   (pamrec.py:521-522,714-728)
constexpr int kProjFwdSmem = (3 * kDD + 2 * kTokTile * kRowPad + 2 * kD) * 4;
__global__ void __launch_bounds__(kTokTile)
k_proj_fwd(const float* __restrict__ X, const int* __restrict__ perm, const int* __restrict__ ctl,
           const int* __restrict__ tile_bucket, const int* __restrict__ tile_begin, const int* __restrict__ tile_count,
           const float* __restrict__ Wq, const float* __restrict__ Wk, const float* __restrict__ Wv,
           const float* __restrict__ ln_beta, const float* __restrict__ ln_gamma, float* __restrict__ QIN,
           float* __restrict__ Q, float* __restrict__ K, float* __restrict__ V) {
  int tile = blockIdx.x;
  if (tile >= ctl[32]) return;
  extern __shared__ __align__(16) float sm[];
  float* Ws = sm;
  float* xs = Ws + 3 * kDD;
  float* qs = xs + kTokTile * kRowPad;
  float* lnp = qs + kTokTile * kRowPad;
  __shared__ int toks[kTokTile];
  const int tid = threadIdx.x;
  const int k = tile_bucket[tile], begin = tile_begin[tile], cnt = tile_count[tile];
  for (int i = tid; i < 3 * (kDD / 4); i += kTokTile) {
    int m = i / (kDD / 4), j = i % (kDD / 4);
    const float* src = (m == 0 ? Wq : (m == 1 ? Wk : Wv)) + (int64_t)k * kDD;
    st4(Ws + m * kDD + 4 * j, ld4(src + 4 * j));
  }
  if (tid < kD) { lnp[tid] = ln_beta[tid]; lnp[kD + tid] = ln_gamma[tid]; }
  if (tid < cnt) toks[tid] = perm[begin + tid];
  __syncthreads();
  load_tile_perm(xs, X, toks, cnt, tid);
  __syncthreads();
  if (tid >= cnt) return;
  const float* xr = xs + tid * kRowPad;
  float* qr = qs + tid * kRowPad;
  float mean, rstd;
  row_stats(xr, mean, rstd);
#pragma unroll
  for (int i = 0; i < kD; ++i) qr[i] = fmaf(lnp[kD + i], (xr[i] - mean) * rstd, lnp[i]);
  const int64_t tok = toks[tid];
  {
    float acc[kD];
#pragma unroll
    for (int i = 0; i < kD; ++i) acc[i] = qr[i];
    store_row(QIN + tok * kD, acc);
  }
  {
    float acc[kD];
#pragma unroll
    for (int i = 0; i < kD; ++i) acc[i] = 0.f;
    mv_fwd(qr, Ws, acc);
    store_row(Q + tok * kD, acc);
  }
  {
    float acc[kD];
#pragma unroll
    for (int i = 0; i < kD; ++i) acc[i] = 0.f;
    mv_fwd(xr, Ws + kDD, acc);
    store_row(K + tok * kD, acc);
  }
  {
    float acc[kD];
#pragma unroll
    for (int i = 0; i < kD; ++i) acc[i] = 0.f;
    mv_fwd(xr, Ws + 2 * kDD, acc);
    store_row(V + tok * kD, acc);
  }
}

void launch_proj_fwd(const float* X, const int* perm, const int* ctl, const int* tile_bucket, const int* tile_begin,
                     const int* tile_count, int max_tiles, const float* Wq, const float* Wk, const float* Wv,
                     const float* ln_beta, const float* ln_gamma, float* QIN, float* Q, float* K, float* V,
                     cudaStream_t st) {
  PAMREC_PROF("proj_fwd", 1, st);
  static bool once = false;
  if (!once) { cudaFuncSetAttribute(k_proj_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, kProjFwdSmem); once = true; }
  k_proj_fwd<<<max_tiles, kTokTile, kProjFwdSmem, st>>>(X, perm, ctl, tile_bucket, tile_begin, tile_count, Wq, Wk, Wv,
                                                        ln_beta, ln_gamma, QIN, Q, K, V);
}

// ------------------------------------------------------------------------------------------
// A2-A4 forward: P = softmax_j(mask ? Q K^T / sqrt(40) : -(2^32)+1);  y = P V + qin  (pamrec.py:768-810)
// CTA per sample, warp per query row, lanes over keys; P.V with a (3 key-groups x 10 chunks) lane map.
__host__ __device__ inline int attn_tp(int T) { return (T + 3) & ~3; }
inline size_t attn_fwd_smem(int T) { return (size_t)(3 * T * kAttnStride + 8 * attn_tp(T)) * 4 + (size_t)T * 4; }

template <int NJ>
__global__ void __launch_bounds__(256)
k_attn_fwd(const float* __restrict__ Q, const float* __restrict__ K, const float* __restrict__ V,
           const float* __restrict__ QIN, const int* __restrict__ mask, float* __restrict__ Y, int T) {
  extern __shared__ __align__(16) float sm[];
  const int Tp = attn_tp(T);
  float* Ks = sm;
  float* Vs = Ks + T * kAttnStride;
  float* Qs = Vs + T * kAttnStride;
  float* Ps = Qs + T * kAttnStride;
  int* mk = reinterpret_cast<int*>(Ps + 8 * Tp);
  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
  const int64_t base = (int64_t)blockIdx.x * T;
  for (int i = tid; i < T * 10; i += 256) {
    int r = i / 10, c = i % 10;
    st4(Ks + r * kAttnStride + 4 * c, ld4(K + (base + r) * kD + 4 * c));
    st4(Vs + r * kAttnStride + 4 * c, ld4(V + (base + r) * kD + 4 * c));
    st4(Qs + r * kAttnStride + 4 * c, ld4(Q + (base + r) * kD + 4 * c));
  }
  for (int i = tid; i < T; i += 256) mk[i] = mask[base + i];
  __syncthreads();
  const float scale = sqrtf((float)kD);
  float* pw = Ps + w * Tp;
  const int jg = lane / 10, c = lane % 10;
  for (int t = w; t < T; t += 8) {
    float4 q[10];
#pragma unroll
    for (int i = 0; i < 10; ++i) q[i] = ld4(Qs + t * kAttnStride + 4 * i);
    float s[NJ];
    float m = -INFINITY;
#pragma unroll
    for (int jj = 0; jj < NJ; ++jj) {
      int j = jj * 32 + lane;
      float val = -INFINITY;
      if (j < T) {
        float d = 0.f;
#pragma unroll
        for (int i = 0; i < 10; ++i) d += f4_dot(q[i], ld4(Ks + j * kAttnStride + 4 * i));
        val = mk[j] ? d / scale : kMaskNeg;
      }
      s[jj] = val;
      m = fmaxf(m, val);
    }
    m = warp_max(m);
    float sum = 0.f;
#pragma unroll
    for (int jj = 0; jj < NJ; ++jj) {
      int j = jj * 32 + lane;
      float e = (j < T) ? expf(s[jj] - m) : 0.f;
      s[jj] = e;
      sum += e;
    }
    sum = warp_sum(sum);
#pragma unroll
    for (int jj = 0; jj < NJ; ++jj) {
      int j = jj * 32 + lane;
      if (j < T) pw[j] = s[jj] / sum;
    }
    __syncwarp();
    float4 acc = f4_zero();
    if (lane < 30)
      for (int j = jg; j < T; j += 3) f4_fma(acc, pw[j], ld4(Vs + j * kAttnStride + 4 * c));
    float4 a1 = f4_shfl_down(acc, 10), a2 = f4_shfl_down(acc, 20);
    if (lane < 10) {
      float4 r = ld4(QIN + (base + t) * kD + 4 * c);
      acc.x += a1.x + a2.x + r.x; acc.y += a1.y + a2.y + r.y; acc.z += a1.z + a2.z + r.z; acc.w += a1.w + a2.w + r.w;
      st4(Y + (base + t) * kD + 4 * c, acc);
    }
    __syncwarp();
  }
}

template <int NJ>
static void attn_fwd_nj(const float* Q, const float* K, const float* V, const float* QIN, const int* mask, float* Y, int B,
                        int T, size_t smem, cudaStream_t st) {
  static size_t smem_set = 0;
  if (smem > smem_set) { cudaFuncSetAttribute(k_attn_fwd<NJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); smem_set = smem; }
  k_attn_fwd<NJ><<<B, 256, smem, st>>>(Q, K, V, QIN, mask, Y, T);
}
void launch_attn_fwd(const float* Q, const float* K, const float* V, const float* QIN, const int* mask, float* Y, int B,
                     int T, cudaStream_t st) {
  PAMREC_PROF("attn_fwd", 1, st);
  if (B == 0) return;
  size_t smem = attn_fwd_smem(T);
  switch ((T + 31) / 32) {
    case 1: attn_fwd_nj<1>(Q, K, V, QIN, mask, Y, B, T, smem, st); break;
    case 2: attn_fwd_nj<2>(Q, K, V, QIN, mask, Y, B, T, smem, st); break;
    case 3: attn_fwd_nj<3>(Q, K, V, QIN, mask, Y, B, T, smem, st); break;
    case 4: attn_fwd_nj<4>(Q, K, V, QIN, mask, Y, B, T, smem, st); break;
    case 5: attn_fwd_nj<5>(Q, K, V, QIN, mask, Y, B, T, smem, st); break;
    case 6: attn_fwd_nj<6>(Q, K, V, QIN, mask, Y, B, T, smem, st); break;
    case 7: attn_fwd_nj<7>(Q, K, V, QIN, mask, Y, B, T, smem, st); break;
    default: attn_fwd_nj<8>(Q, K, V, QIN, mask, Y, B, T, smem, st); break;
  }
}

// ------------------------------------------------------------------------------------------
// N1 + F1 forward: f = LN_b(y); out = relu(f W1 + b1) W2 + b2 + f     (pamrec.py:534-535,565-577)
constexpr int kFfnFwdSmem = (2 * kDD + 2 * kTokTile * kRowPad + 4 * kD) * 4;
__global__ void __launch_bounds__(kTokTile)
k_ffn_fwd(const float* __restrict__ Y, const float* __restrict__ W1, const float* __restrict__ b1,
          const float* __restrict__ W2, const float* __restrict__ b2, const float* __restrict__ ln_beta,
          const float* __restrict__ ln_gamma, float* __restrict__ OUT, int n_tok) {
  extern __shared__ __align__(16) float sm[];
  float* W1s = sm;
  float* W2s = W1s + kDD;
  float* ys = W2s + kDD;
  float* fs = ys + kTokTile * kRowPad;
  float* pr = fs + kTokTile * kRowPad;   // b1 | b2 | beta | gamma
  const int tid = threadIdx.x;
  const int64_t tok0 = (int64_t)blockIdx.x * kTokTile;
  const int cnt = (int)min((int64_t)kTokTile, (int64_t)n_tok - tok0);
  for (int i = tid; i < kDD / 4; i += kTokTile) {
    st4(W1s + 4 * i, ld4(W1 + 4 * i));
    st4(W2s + 4 * i, ld4(W2 + 4 * i));
  }
  if (tid < kD) { pr[tid] = b1[tid]; pr[kD + tid] = b2[tid]; pr[2 * kD + tid] = ln_beta[tid]; pr[3 * kD + tid] = ln_gamma[tid]; }
  load_tile_lin(ys, Y + tok0 * kD, cnt, tid);
  __syncthreads();
  if (tid >= cnt) return;
  float* yr = ys + tid * kRowPad;
  float* fr = fs + tid * kRowPad;
  float mean, rstd;
  row_stats(yr, mean, rstd);
#pragma unroll
  for (int i = 0; i < kD; ++i) fr[i] = fmaf(pr[3 * kD + i], (yr[i] - mean) * rstd, pr[2 * kD + i]);
  {
    float acc[kD];
#pragma unroll
    for (int i = 0; i < kD; ++i) acc[i] = pr[i];
    mv_fwd(fr, W1s, acc);
#pragma unroll
    for (int i = 0; i < kD; ++i) yr[i] = fmaxf(acc[i], 0.f);
  }
  {
    float acc[kD];
#pragma unroll
    for (int i = 0; i < kD; ++i) acc[i] = pr[kD + i] + fr[i];
    mv_fwd(yr, W2s, acc);
    store_row(OUT + (tok0 + tid) * kD, acc);
  }
}

void launch_ffn_fwd(const float* Y, const float* W1, const float* b1, const float* W2, const float* b2,
                    const float* ln_beta, const float* ln_gamma, float* OUT, int n_tok, cudaStream_t st) {
  PAMREC_PROF("ffn_fwd", 1, st);
  if (n_tok == 0) return;
  static bool once = false;
  if (!once) { cudaFuncSetAttribute(k_ffn_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, kFfnFwdSmem); once = true; }
  k_ffn_fwd<<<(n_tok + kTokTile - 1) / kTokTile, kTokTile, kFfnFwdSmem, st>>>(Y, W1, b1, W2, b2, ln_beta, ln_gamma, OUT, n_tok);
}

// ------------------------------------------------------------------------------------------
// FFN + LN_b backward.  In: dOUT (grad of block output), Y.  Out: dY, and atomically
// accumulated dW1, db1, dW2, db2, dbeta, dgamma.
constexpr int kFfnBwdSmem = (2 * kDD + 4 * kTokTile * kRowPad + 4 * kD + 2 * kD) * 4;
__global__ void __launch_bounds__(kTokTile)
k_ffn_bwd(const float* __restrict__ Y, const float* __restrict__ dOUT, const float* __restrict__ W1,
          const float* __restrict__ b1, const float* __restrict__ W2, const float* __restrict__ ln_beta,
          const float* __restrict__ ln_gamma, float* __restrict__ dY, float* __restrict__ dW1, float* __restrict__ db1,
          float* __restrict__ dW2, float* __restrict__ db2, float* __restrict__ dbeta, float* __restrict__ dgamma,
          int n_tok) {
  extern __shared__ __align__(16) float sm[];
  float* W1s = sm;
  float* W2s = W1s + kDD;
  float* fs = W2s + kDD;                       // y, then f
  float* hs = fs + kTokTile * kRowPad;         // relu(f W1 + b1)
  float* gs = hs + kTokTile * kRowPad;         // dOUT
  float* ps = gs + kTokTile * kRowPad;         // d(pre-activation)
  float* pr = ps + kTokTile * kRowPad;         // b1 | - | beta | gamma
  float* red = pr + 4 * kD;                    // dbeta | dgamma partials
  const int tid = threadIdx.x, lane = tid & 31;
  const int64_t tok0 = (int64_t)blockIdx.x * kTokTile;
  const int cnt = (int)min((int64_t)kTokTile, (int64_t)n_tok - tok0);
  for (int i = tid; i < kDD / 4; i += kTokTile) {
    st4(W1s + 4 * i, ld4(W1 + 4 * i));
    st4(W2s + 4 * i, ld4(W2 + 4 * i));
  }
  if (tid < kD) { pr[tid] = b1[tid]; pr[2 * kD + tid] = ln_beta[tid]; pr[3 * kD + tid] = ln_gamma[tid]; }
  if (tid < 2 * kD) red[tid] = 0.f;
  load_tile_lin(fs, Y + tok0 * kD, cnt, tid);
  load_tile_lin(gs, dOUT + tok0 * kD, cnt, tid);
  __syncthreads();
  float df[kD];
  float mean = 0.f, rstd = 0.f;
  const bool active = tid < cnt;
  float* fr = fs + tid * kRowPad;
  float* hr = hs + tid * kRowPad;
  float* gr = gs + tid * kRowPad;
  float* pp = ps + tid * kRowPad;
#pragma unroll
  for (int i = 0; i < kD; ++i) df[i] = 0.f;
  if (active) {
    row_stats(fr, mean, rstd);
#pragma unroll
    for (int i = 0; i < kD; ++i) fr[i] = fmaf(pr[3 * kD + i], (fr[i] - mean) * rstd, pr[2 * kD + i]);
    {
      float acc[kD];
#pragma unroll
      for (int i = 0; i < kD; ++i) acc[i] = pr[i];
      mv_fwd(fr, W1s, acc);
#pragma unroll
      for (int i = 0; i < kD; ++i) hr[i] = fmaxf(acc[i], 0.f);
    }
    {
      float dh[kD];
#pragma unroll
      for (int i = 0; i < kD; ++i) dh[i] = 0.f;
      mv_bwd(gr, W2s, dh);
#pragma unroll
      for (int i = 0; i < kD; ++i) pp[i] = hr[i] > 0.f ? dh[i] : 0.f;
    }
#pragma unroll
    for (int i = 0; i < kD; ++i) df[i] = gr[i];
    mv_bwd(pp, W1s, df);
  }
  // LN backward (needs xhat = (y - mean) * rstd: y re-read from global, L2-resident)
  float m1 = 0.f, m2 = 0.f;
  float xh[kD];
#pragma unroll
  for (int i = 0; i < kD; ++i) xh[i] = 0.f;
  if (active) {
    const float* yg = Y + (tok0 + tid) * kD;
#pragma unroll
    for (int j = 0; j < kD / 4; ++j) {
      float4 v = ld4(yg + 4 * j);
      xh[4 * j] = (v.x - mean) * rstd; xh[4 * j + 1] = (v.y - mean) * rstd;
      xh[4 * j + 2] = (v.z - mean) * rstd; xh[4 * j + 3] = (v.w - mean) * rstd;
    }
#pragma unroll
    for (int i = 0; i < kD; ++i) {
      float dxh = df[i] * pr[3 * kD + i];
      m1 += dxh;
      m2 = fmaf(dxh, xh[i], m2);
    }
    m1 *= (1.0f / kD);
    m2 *= (1.0f / kD);
    float out[kD];
#pragma unroll
    for (int i = 0; i < kD; ++i) out[i] = rstd * (df[i] * pr[3 * kD + i] - m1 - xh[i] * m2);
    store_row(dY + (tok0 + tid) * kD, out);
  }
  // dbeta = sum df, dgamma = sum df * xhat  (all lanes take part; inactive lanes carry zeros)
#pragma unroll
  for (int i = 0; i < kD; ++i) {
    float sb = warp_sum(df[i]);
    float sg = warp_sum(df[i] * xh[i]);
    if (lane == 0) { atomicAdd(red + i, sb); atomicAdd(red + kD + i, sg); }
  }
  __syncthreads();
  tile_outer_atomic(hs, gs, cnt, dW2, tid);
  tile_outer_atomic(fs, ps, cnt, dW1, tid);
  tile_colsum_atomic(gs, cnt, db2, tid);
  tile_colsum_atomic(ps, cnt, db1, tid);
  if (tid < kD) { atomicAdd(dbeta + tid, red[tid]); atomicAdd(dgamma + tid, red[kD + tid]); }
}

void launch_ffn_bwd(const float* Y, const float* dOUT, const float* W1, const float* b1, const float* W2,
                    const float* ln_beta, const float* ln_gamma, float* dY, float* dW1, float* db1, float* dW2,
                    float* db2, float* dbeta, float* dgamma, int n_tok, cudaStream_t st) {
  PAMREC_PROF("ffn_bwd", 1, st);
  if (n_tok == 0) return;
  static bool once = false;
  if (!once) { cudaFuncSetAttribute(k_ffn_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, kFfnBwdSmem); once = true; }
  k_ffn_bwd<<<(n_tok + kTokTile - 1) / kTokTile, kTokTile, kFfnBwdSmem, st>>>(Y, dOUT, W1, b1, W2, ln_beta, ln_gamma, dY, dW1,
                                                                             db1, dW2, db2, dbeta, dgamma, n_tok);
}

// ------------------------------------------------------------------------------------------
// Attention backward.  In: Q, K, V, dY (grad of y = P V + qin).  Out: dQ, dK, dV.
// Pass A (warp per query row) recomputes the softmax row, forms dS and dQ; pass B (warp per
// key) recomputes the column from the saved row statistics and forms dK, dV.
inline size_t attn_bwd_smem(int T) { return (size_t)(4 * T * kAttnStride + 16 * attn_tp(T) + 3 * T) * 4 + (size_t)T * 4; }

template <int NJ>
__global__ void __launch_bounds__(256)
k_attn_bwd(const float* __restrict__ Q, const float* __restrict__ K, const float* __restrict__ V,
           const float* __restrict__ dY, const int* __restrict__ mask, float* __restrict__ dQ, float* __restrict__ dK,
           float* __restrict__ dV, int T) {
  extern __shared__ __align__(16) float sm[];
  const int Tp = attn_tp(T);
  float* Ks = sm;
  float* Vs = Ks + T * kAttnStride;
  float* Qs = Vs + T * kAttnStride;
  float* Gs = Qs + T * kAttnStride;
  float* Ps = Gs + T * kAttnStride;   // 8 x Tp
  float* Ds = Ps + 8 * Tp;            // 8 x Tp
  float* rowm = Ds + 8 * Tp;
  float* rowl = rowm + T;
  float* rowD = rowl + T;
  int* mk = reinterpret_cast<int*>(rowD + T);
  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
  const int64_t base = (int64_t)blockIdx.x * T;
  for (int i = tid; i < T * 10; i += 256) {
    int r = i / 10, c = i % 10;
    st4(Ks + r * kAttnStride + 4 * c, ld4(K + (base + r) * kD + 4 * c));
    st4(Vs + r * kAttnStride + 4 * c, ld4(V + (base + r) * kD + 4 * c));
    st4(Qs + r * kAttnStride + 4 * c, ld4(Q + (base + r) * kD + 4 * c));
    st4(Gs + r * kAttnStride + 4 * c, ld4(dY + (base + r) * kD + 4 * c));
  }
  for (int i = tid; i < T; i += 256) mk[i] = mask[base + i];
  __syncthreads();
  const float scale = sqrtf((float)kD);
  float* pw = Ps + w * Tp;
  float* dw = Ds + w * Tp;
  const int jg = lane / 10, c = lane % 10;
  // ---- pass A
  for (int t = w; t < T; t += 8) {
    float4 q[10], g[10];
#pragma unroll
    for (int i = 0; i < 10; ++i) { q[i] = ld4(Qs + t * kAttnStride + 4 * i); g[i] = ld4(Gs + t * kAttnStride + 4 * i); }
    float s[NJ], dp[NJ];
    float m = -INFINITY;
#pragma unroll
    for (int jj = 0; jj < NJ; ++jj) {
      int j = jj * 32 + lane;
      float val = -INFINITY, dpv = 0.f;
      if (j < T) {
        float d = 0.f;
#pragma unroll
        for (int i = 0; i < 10; ++i) d += f4_dot(q[i], ld4(Ks + j * kAttnStride + 4 * i));
        val = mk[j] ? d / scale : kMaskNeg;
#pragma unroll
        for (int i = 0; i < 10; ++i) dpv += f4_dot(g[i], ld4(Vs + j * kAttnStride + 4 * i));
      }
      s[jj] = val; dp[jj] = dpv;
      m = fmaxf(m, val);
    }
    m = warp_max(m);
    float sum = 0.f;
#pragma unroll
    for (int jj = 0; jj < NJ; ++jj) {
      int j = jj * 32 + lane;
      float e = (j < T) ? expf(s[jj] - m) : 0.f;
      s[jj] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    float Dv = 0.f;
#pragma unroll
    for (int jj = 0; jj < NJ; ++jj) {
      int j = jj * 32 + lane;
      float p = (j < T) ? s[jj] / sum : 0.f;
      s[jj] = p;
      Dv = fmaf(p, dp[jj], Dv);
    }
    Dv = warp_sum(Dv);
    if (lane == 0) { rowm[t] = m; rowl[t] = sum; rowD[t] = Dv; }
#pragma unroll
    for (int jj = 0; jj < NJ; ++jj) {
      int j = jj * 32 + lane;
      if (j < T) pw[j] = mk[j] ? s[jj] * (dp[jj] - Dv) / scale : 0.f;
    }
    __syncwarp();
    float4 acc = f4_zero();
    if (lane < 30)
      for (int j = jg; j < T; j += 3) f4_fma(acc, pw[j], ld4(Ks + j * kAttnStride + 4 * c));
    float4 a1 = f4_shfl_down(acc, 10), a2 = f4_shfl_down(acc, 20);
    if (lane < 10) {
      acc.x += a1.x + a2.x; acc.y += a1.y + a2.y; acc.z += a1.z + a2.z; acc.w += a1.w + a2.w;
      st4(dQ + (base + t) * kD + 4 * c, acc);
    }
    __syncwarp();
  }
  __syncthreads();
  // ---- pass B
  for (int j = w; j < T; j += 8) {
    float4 kj[10], vj[10];
#pragma unroll
    for (int i = 0; i < 10; ++i) { kj[i] = ld4(Ks + j * kAttnStride + 4 * i); vj[i] = ld4(Vs + j * kAttnStride + 4 * i); }
    const int mkj = mk[j];
#pragma unroll
    for (int tt = 0; tt < NJ; ++tt) {
      int t = tt * 32 + lane;
      if (t < T) {
        float d = 0.f, dpv = 0.f;
#pragma unroll
        for (int i = 0; i < 10; ++i) d += f4_dot(ld4(Qs + t * kAttnStride + 4 * i), kj[i]);
#pragma unroll
        for (int i = 0; i < 10; ++i) dpv += f4_dot(ld4(Gs + t * kAttnStride + 4 * i), vj[i]);
        float sv = mkj ? d / scale : kMaskNeg;
        float p = expf(sv - rowm[t]) / rowl[t];
        pw[t] = p;
        dw[t] = mkj ? p * (dpv - rowD[t]) / scale : 0.f;
      }
    }
    __syncwarp();
    float4 aK = f4_zero(), aV = f4_zero();
    if (lane < 30)
      for (int t = jg; t < T; t += 3) {
        f4_fma(aK, dw[t], ld4(Qs + t * kAttnStride + 4 * c));
        f4_fma(aV, pw[t], ld4(Gs + t * kAttnStride + 4 * c));
      }
    float4 k1 = f4_shfl_down(aK, 10), k2 = f4_shfl_down(aK, 20);
    float4 v1 = f4_shfl_down(aV, 10), v2 = f4_shfl_down(aV, 20);
    if (lane < 10) {
      aK.x += k1.x + k2.x; aK.y += k1.y + k2.y; aK.z += k1.z + k2.z; aK.w += k1.w + k2.w;
      aV.x += v1.x + v2.x; aV.y += v1.y + v2.y; aV.z += v1.z + v2.z; aV.w += v1.w + v2.w;
      st4(dK + (base + j) * kD + 4 * c, aK);
      st4(dV + (base + j) * kD + 4 * c, aV);
    }
    __syncwarp();
  }
}

template <int NJ>
static void attn_bwd_nj(const float* Q, const float* K, const float* V, const float* dY, const int* mask, float* dQ, float* dK,
                        float* dV, int B, int T, size_t smem, cudaStream_t st) {
  static size_t smem_set = 0;
  if (smem > smem_set) { cudaFuncSetAttribute(k_attn_bwd<NJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); smem_set = smem; }
  k_attn_bwd<NJ><<<B, 256, smem, st>>>(Q, K, V, dY, mask, dQ, dK, dV, T);
}
void launch_attn_bwd(const float* Q, const float* K, const float* V, const float* dY, const int* mask, float* dQ,
                     float* dK, float* dV, int B, int T, cudaStream_t st) {
  PAMREC_PROF("attn_bwd", 1, st);
  if (B == 0) return;
  size_t smem = attn_bwd_smem(T);
  switch ((T + 31) / 32) {
    case 1: attn_bwd_nj<1>(Q, K, V, dY, mask, dQ, dK, dV, B, T, smem, st); break;
    case 2: attn_bwd_nj<2>(Q, K, V, dY, mask, dQ, dK, dV, B, T, smem, st); break;
    case 3: attn_bwd_nj<3>(Q, K, V, dY, mask, dQ, dK, dV, B, T, smem, st); break;
    case 4: attn_bwd_nj<4>(Q, K, V, dY, mask, dQ, dK, dV, B, T, smem, st); break;
    case 5: attn_bwd_nj<5>(Q, K, V, dY, mask, dQ, dK, dV, B, T, smem, st); break;
    case 6: attn_bwd_nj<6>(Q, K, V, dY, mask, dQ, dK, dV, B, T, smem, st); break;
    case 7: attn_bwd_nj<7>(Q, K, V, dY, mask, dQ, dK, dV, B, T, smem, st); break;
    default: attn_bwd_nj<8>(Q, K, V, dY, mask, dQ, dK, dV, B, T, smem, st); break;
  }
}

// ------------------------------------------------------------------------------------------
// Projection + LN_a backward (bucket-sorted tiles).  In: X (block input), dY (residual path
// into qin), dQ, dK, dV.  Out: dX; atomically accumulated dWq/dWk/dWv[bucket], dbeta, dgamma.
constexpr int kProjBwdSmem = (3 * kDD + 3 * kTokTile * kRowPad + 2 * kD + 2 * kD) * 4;
__global__ void __launch_bounds__(kTokTile)
k_proj_bwd(const float* __restrict__ X, const float* __restrict__ dY, const float* __restrict__ dQ,
           const float* __restrict__ dK, const float* __restrict__ dV, const int* __restrict__ perm,
           const int* __restrict__ ctl, const int* __restrict__ tile_bucket, const int* __restrict__ tile_begin,
           const int* __restrict__ tile_count, const float* __restrict__ Wq, const float* __restrict__ Wk,
           const float* __restrict__ Wv, const float* __restrict__ ln_beta, const float* __restrict__ ln_gamma,
           float* __restrict__ dX, float* __restrict__ dWq, float* __restrict__ dWk, float* __restrict__ dWv,
           float* __restrict__ dbeta, float* __restrict__ dgamma) {
  int tile = blockIdx.x;
  if (tile >= ctl[32]) return;
  extern __shared__ __align__(16) float sm[];
  float* Ws = sm;
  float* xs = Ws + 3 * kDD;
  float* qs = xs + kTokTile * kRowPad;
  float* gs = qs + kTokTile * kRowPad;
  float* lnp = gs + kTokTile * kRowPad;
  float* red = lnp + 2 * kD;
  __shared__ int toks[kTokTile];
  const int tid = threadIdx.x, lane = tid & 31;
  const int k = tile_bucket[tile], begin = tile_begin[tile], cnt = tile_count[tile];
  for (int i = tid; i < 3 * (kDD / 4); i += kTokTile) {
    int m = i / (kDD / 4), j = i % (kDD / 4);
    const float* src = (m == 0 ? Wq : (m == 1 ? Wk : Wv)) + (int64_t)k * kDD;
    st4(Ws + m * kDD + 4 * j, ld4(src + 4 * j));
  }
  if (tid < kD) { lnp[tid] = ln_beta[tid]; lnp[kD + tid] = ln_gamma[tid]; }
  if (tid < 2 * kD) red[tid] = 0.f;
  if (tid < cnt) toks[tid] = perm[begin + tid];
  __syncthreads();
  load_tile_perm(xs, X, toks, cnt, tid);
  load_tile_perm(gs, dQ, toks, cnt, tid);
  __syncthreads();
  const bool active = tid < cnt;
  const float* xr = xs + tid * kRowPad;
  float* qr = qs + tid * kRowPad;
  const float* gr = gs + tid * kRowPad;
  float mean = 0.f, rstd = 0.f;
  float dqin[kD], dx[kD];
#pragma unroll
  for (int i = 0; i < kD; ++i) { dqin[i] = 0.f; dx[i] = 0.f; }
  if (active) {
    row_stats(xr, mean, rstd);
#pragma unroll
    for (int i = 0; i < kD; ++i) qr[i] = fmaf(lnp[kD + i], (xr[i] - mean) * rstd, lnp[i]);
    mv_bwd(gr, Ws, dqin);
  }
  __syncthreads();                                   // qs complete
  tile_outer_atomic(qs, gs, cnt, dWq + (int64_t)k * kDD, tid);
  __syncthreads();
  load_tile_perm(gs, dK, toks, cnt, tid);
  __syncthreads();
  if (active) mv_bwd(gr, Ws + kDD, dx);
  tile_outer_atomic(xs, gs, cnt, dWk + (int64_t)k * kDD, tid);
  __syncthreads();
  load_tile_perm(gs, dV, toks, cnt, tid);
  __syncthreads();
  if (active) mv_bwd(gr, Ws + 2 * kDD, dx);
  tile_outer_atomic(xs, gs, cnt, dWv + (int64_t)k * kDD, tid);
  // LN_a backward
  float xh[kD];
#pragma unroll
  for (int i = 0; i < kD; ++i) xh[i] = 0.f;
  if (active) {
    const int64_t tok = toks[tid];
    const float* yg = dY + tok * kD;
#pragma unroll
    for (int j = 0; j < kD / 4; ++j) {
      float4 v = ld4(yg + 4 * j);
      dqin[4 * j] += v.x; dqin[4 * j + 1] += v.y; dqin[4 * j + 2] += v.z; dqin[4 * j + 3] += v.w;
    }
    float m1 = 0.f, m2 = 0.f;
#pragma unroll
    for (int i = 0; i < kD; ++i) {
      xh[i] = (xr[i] - mean) * rstd;
      float dxh = dqin[i] * lnp[kD + i];
      m1 += dxh;
      m2 = fmaf(dxh, xh[i], m2);
    }
    m1 *= (1.0f / kD);
    m2 *= (1.0f / kD);
#pragma unroll
    for (int i = 0; i < kD; ++i) dx[i] += rstd * (dqin[i] * lnp[kD + i] - m1 - xh[i] * m2);
    store_row(dX + tok * kD, dx);
  }
#pragma unroll
  for (int i = 0; i < kD; ++i) {
    float sb = warp_sum(dqin[i]);
    float sg = warp_sum(dqin[i] * xh[i]);
    if (lane == 0) { atomicAdd(red + i, sb); atomicAdd(red + kD + i, sg); }
  }
  __syncthreads();
  if (tid < kD) { atomicAdd(dbeta + tid, red[tid]); atomicAdd(dgamma + tid, red[kD + tid]); }
}

void launch_proj_bwd(const float* X, const float* dY, const float* dQ, const float* dK, const float* dV, const int* perm,
                     const int* ctl, const int* tile_bucket, const int* tile_begin, const int* tile_count, int max_tiles,
                     const float* Wq, const float* Wk, const float* Wv, const float* ln_beta, const float* ln_gamma,
                     float* dX, float* dWq, float* dWk, float* dWv, float* dbeta, float* dgamma, cudaStream_t st) {
  PAMREC_PROF("proj_bwd", 1, st);
  static bool once = false;
  if (!once) { cudaFuncSetAttribute(k_proj_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, kProjBwdSmem); once = true; }
  k_proj_bwd<<<max_tiles, kTokTile, kProjBwdSmem, st>>>(X, dY, dQ, dK, dV, perm, ctl, tile_bucket, tile_begin, tile_count, Wq,
                                                        Wk, Wv, ln_beta, ln_gamma, dX, dWq, dWk, dWv, dbeta, dgamma);
}

}  // namespace pamrec
