// Sparse embedding backward, second generation: ONE sort for the three id spaces, a sequential run walk without shuffles,
// row-wise Adam fused into the walk.  (The first generation - three CUB pipelines, shuffle-scan segmented reduction, accumulator
// memset of one slot per POSITION, separate L2 / Adam / slot-reset launches - stays for the row-sharded exchange path.)
//
// TF semantics reproduced (base_model.py:288-304, sequential_base_model.py:640-664, TF 2.4 AdamOptimizer._apply_sparse_shared):
//  * the gradient of a table is the concatenation of every lookup's rows (history, target) and of the L2 rows of tf.unique(ids);
//    tf.clip_by_norm takes its norm over those un-deduplicated rows;
//  * duplicate rows are summed, then Adam moves every row (DENSE_EXACT) or the looked-up rows (LAZY).
//
// Plan (depends on the ids only; runs on the side stream beside the forward pass):
//   k_sp2_keys      keys of all lookups in one id space: item ids | n_items + cate ids | n_items + n_cates + user ids
//   cub radix sort  (key, position) pairs; cub scan of the run-head flags (computed on the fly) -> rank of every position
//   k_sp2_fill      per-table unique lists / counts, row -> unique index map (DENSE_EXACT), |w|^2 of the unique rows (the L2 rows'
//                   share of the clip norm and of the regularisation loss), zero of the accumulator rows of runs that cross a
//                   warp block
// Gradient (after the backward pass produced dX0):
//   k_sp2_walk      a group of 4 lanes (item, 64-byte rows) or one lane (category, 16-byte rows) walks 16 consecutive sorted
//                   positions and sums runs of equal ids in registers; the first / last run of each group are merged across the
//                   warp through shared memory; a run that lies inside the warp's block is FINAL: its sum is stored to the compact
//                   accumulator (DENSE_EXACT), to the dense gradient table (replicated data parallel) or goes straight through
//                   Adam into the table row (LAZY) - no atomics, no zero-fill; only runs that cross a warp block (hot rows: one
//                   partial per 128 / 512 positions) use red.global.add.v4.f32.
//   the un-deduplicated squared norm comes from the kernel that already reads all of dX0 (k_dtgt_total, kernels_optim.cu), so the
//   clip scale is known before the walk and Adam can be applied inside it.
// Apply:
//   k_sp2_adam_sweep   DENSE_EXACT: one launch over all four tables (item, cate, user_long, user_short)
//   k_sp2_lazy_finish  LAZY: user rows (L2 only) and the few rows whose runs crossed a warp block
#include <cub/cub.cuh>
#include <thrust/iterator/counting_iterator.h>
#include <thrust/iterator/transform_iterator.h>

#include "kernels.h"

namespace pamrec {

constexpr int kSp2Sub = 16;                          // sorted positions walked by one lane group
constexpr int kSp2Warps = 4;                         // warps per CTA of the walk
__host__ __device__ constexpr int sp2_wb(int ch) { return (32 / ch) * kSp2Sub; }   // sorted positions per warp: 128 (item) / 512 (cate)

__device__ __forceinline__ void red_add4(float* p, const float4& v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void f4_add(float4& a, const float4& b) { a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }
__device__ __forceinline__ int sp2_table_of_pos(int64_t p, int64_t NI) { return p < NI ? 0 : (p < 2 * NI ? 1 : 2); }
__device__ __forceinline__ int sp2_base(const Sp2& s, int t) { return t == 0 ? 0 : (t == 1 ? s.n_items : s.n_items + s.n_cates); }

__device__ __forceinline__ float sp2_clip_scale(const Sp2& s, const AdamP& a, int t) {
  if (!a.is_clip) return 1.0f;
  const double nsq = s.undedup[t] + (double)a.l2 * (double)a.l2 * s.l2sq[t];
  const float norm = (float)sqrt(nsq);
  return a.clip / fmaxf(norm, a.clip);             // tf.clip_by_norm: t * clip / max(norm, clip)
}
__device__ __forceinline__ void sp2_adam4(float4& w, float4& m, float4& v, const float4& g, const AdamP& a) {
  // TF _apply_sparse_shared: m = m*b1 + g*(1-b1); v = v*b2 + g*g*(1-b2); w -= lr*m/(sqrt(v)+eps)
  const float b1 = a.b1, b2 = a.b2, lr = a.lr_dev ? __ldg(a.lr_dev) : a.lr, eps = a.eps;
  m.x = m.x * b1 + g.x * (1.f - b1); m.y = m.y * b1 + g.y * (1.f - b1);
  m.z = m.z * b1 + g.z * (1.f - b1); m.w = m.w * b1 + g.w * (1.f - b1);
  v.x = v.x * b2 + (g.x * g.x) * (1.f - b2); v.y = v.y * b2 + (g.y * g.y) * (1.f - b2);
  v.z = v.z * b2 + (g.z * g.z) * (1.f - b2); v.w = v.w * b2 + (g.w * g.w) * (1.f - b2);
  w.x -= lr * m.x / (sqrtf(v.x) + eps); w.y -= lr * m.y / (sqrtf(v.y) + eps);
  w.z -= lr * m.z / (sqrtf(v.z) + eps); w.w -= lr * m.w / (sqrtf(v.w) + eps);
}
// one 16-byte chunk of a touched row: g = scale * (a + l2 * w), then Adam
__device__ __forceinline__ void sp2_row_chunk(float* tw, float* tm, float* tv, int64_t o, const float4& acc, float scale, const AdamP& a) {
  float4 w = ld4(tw + o), m = ld4(tm + o), v = ld4(tv + o);
  float4 g;
  g.x = scale * (acc.x + a.l2 * w.x); g.y = scale * (acc.y + a.l2 * w.y);
  g.z = scale * (acc.z + a.l2 * w.z); g.w = scale * (acc.w + a.l2 * w.w);
  sp2_adam4(w, m, v, g, a);
  st4(tw + o, w); st4(tm + o, m); st4(tv + o, v);
}

// ------------------------------------------------------------------------------------------ plan
__global__ void __launch_bounds__(256) k_sp2_keys(const Sp2 s) {
  const int64_t NI = s.N + s.B, n = 2 * NI + s.B;
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p == 0) { s.l2sq[0] = 0.0; s.l2sq[1] = 0.0; s.l2sq[2] = 0.0; s.l2sq[3] = 0.0; }
  if (s.slot != nullptr) {
    // rows marked by the previous plan (its unique lists are still in place: k_sp2_fill of THIS plan runs later on the stream)
    const int u1 = s.meta[1], u2 = s.meta[2], u3 = s.meta[3];
    for (int64_t q = p; q < u3; q += (int64_t)gridDim.x * blockDim.x) {
      const int t = q >= u2 ? 2 : (q >= u1 ? 1 : 0);
      const int ul = (int)q - (t == 2 ? u2 : (t == 1 ? u1 : 0));
      s.slot[sp2_base(s, t) + s.ukeys[t][ul]] = -1;
    }
  }
  if (p >= n) return;
  int key;
  if (p < NI) key = p < s.N ? s.item_hist[p] : s.items[p - s.N];
  else if (p < 2 * NI) { const int64_t q = p - NI; key = s.n_items + (q < s.N ? s.cate_hist[q] : s.cates[q - s.N]); }
  else key = s.n_items + s.n_cates + s.users[p - 2 * NI];
  s.keys[p] = key;
  s.idx[p] = (int)p;
}

struct Sp2HeadFlag {
  const int* sk;
  __host__ __device__ __forceinline__ int operator()(int p) const { return (p == 0 || sk[p] != sk[p - 1]) ? 1 : 0; }
};
using Sp2FlagIter = thrust::transform_iterator<Sp2HeadFlag, thrust::counting_iterator<int>, int>;

__global__ void __launch_bounds__(256) k_sp2_fill(const Sp2 s) {
  __shared__ double sh[4][8];
  const int64_t NI = s.N + s.B, n = 2 * NI + s.B;
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double ss[4] = {0.0, 0.0, 0.0, 0.0};
  if (p < n) {
    const int t = sp2_table_of_pos(p, NI);
    const int64_t r0 = (int64_t)t * NI;
    const int key = s.skeys[p];
    const bool head = p == 0 || s.skeys[p - 1] != key;
    const int u = s.uidx[p] - 1;                               // global unique index
    const int uf = t == 0 ? 0 : s.uidx[r0] - 1;                // first unique index of this table (its first position is a head)
    const int ul = u - uf;
    const int id = key - sp2_base(s, t);
    if (p == r0) s.meta[t] = uf;
    if (p == n - 1) s.meta[3] = u + 1;
    const int64_t r1 = t == 2 ? n : r0 + NI;
    if (p == r1 - 1) s.nuniq[t] = ul + 1;
    if (head) {
      s.ukeys[t][ul] = id;
      if (s.slot != nullptr) s.slot[key] = ul;
      if (s.mode == SP2_DENSE && t == 2) s.rep_grad[(int64_t)s.n_items * kI + (int64_t)s.n_cates * kC + key] = 1.0f;   // users have no walk
      if (s.mode != SP2_DENSE) {
        // L2 rows of tf.unique(ids): |w|^2 (before the update of this step)
        if (t == 0) { const float* w = s.w[0] + (int64_t)id * kI; float q = 0.f;
#pragma unroll
          for (int c = 0; c < 4; ++c) { const float4 x = ld4(w + 4 * c); q += f4_dot(x, x); } ss[0] = (double)q; }
        else if (t == 1) { const float4 x = ld4(s.w[1] + (int64_t)id * kC); ss[1] = (double)f4_dot(x, x); }
        else {
          const float* a = s.w[2] + (int64_t)id * PAMREC_USER_DIM; const float* b = s.w[3] + (int64_t)id * PAMREC_USER_DIM;
          float qa = 0.f, qb = 0.f;
#pragma unroll
          for (int c = 0; c < 5; ++c) { const float4 x = ld4(a + 4 * c), y = ld4(b + 4 * c); qa += f4_dot(x, x); qb += f4_dot(y, y); }
          ss[2] = (double)qa; ss[3] = (double)qb;
        }
      }
    } else if (t < 2 && s.mode != SP2_DENSE) {
      // a run that crosses a warp block of the walk is accumulated with atomics: its accumulator row starts at zero
      const int wb = t == 0 ? sp2_wb(4) : sp2_wb(1);
      if ((p - r0) % wb == 0) {
        if (t == 0) {
#pragma unroll
          for (int c = 0; c < 4; ++c) st4(s.accum[0] + (int64_t)ul * kI + 4 * c, f4_zero());
        } else st4(s.accum[1] + (int64_t)ul * kC, f4_zero());
      }
    }
  }
  if (s.mode == SP2_DENSE) return;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 4; ++k) { const double v = warp_sum_d(ss[k]); if (lane == 0) sh[k][w] = v; }
  __syncthreads();
  if (threadIdx.x < 4) {
    double tot = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) tot += sh[threadIdx.x][k];
    if (tot != 0.0) atomicAdd(s.l2sq + threadIdx.x, tot);
  }
}

size_t sp2_temp_bytes(int64_t n_keys) {
  size_t a = 0, b = 0;
  int* p = nullptr;
  cub::DeviceRadixSort::SortPairs(nullptr, a, p, p, p, p, (int)n_keys, 0, 32, (cudaStream_t)0);
  Sp2FlagIter it(thrust::counting_iterator<int>(0), Sp2HeadFlag{p});
  cub::DeviceScan::InclusiveSum(nullptr, b, it, p, (int)n_keys, (cudaStream_t)0);
  return a > b ? a : b;
}

int launch_sp2_plan(const Sp2& s, void* cub_temp, size_t cub_bytes, cudaStream_t st) {
  PAMREC_PROF("sparse_plan", 9, st);
  const int64_t n = 2 * (s.N + s.B) + s.B;
  // replicated tables: the gradient table and the touch counts start at zero (the previous step's sweep has read them: this
  // stream was forked from the caller's after that step)
  if (s.mode == SP2_DENSE) cudaMemsetAsync(s.rep_grad, 0, (size_t)sp2_rep_floats(s.n_items, s.n_cates, s.n_users) * sizeof(float), st);
  if (n == 0) { cudaMemsetAsync(s.nuniq, 0, 3 * sizeof(int), st); cudaMemsetAsync(s.meta, 0, 4 * sizeof(int), st); cudaMemsetAsync(s.l2sq, 0, 4 * sizeof(double), st); return 0; }
  const unsigned g = (unsigned)((n + 255) / 256);
  k_sp2_keys<<<g, 256, 0, st>>>(s);
  const int64_t range = (int64_t)s.n_items + s.n_cates + s.n_users;
  int bits = 1;
  while (bits < 31 && ((int64_t)1 << bits) < range) ++bits;
  size_t bytes = cub_bytes;
  if (cub::DeviceRadixSort::SortPairs(cub_temp, bytes, s.keys, s.skeys, s.idx, s.sidx, (int)n, 0, bits, st) != cudaSuccess) return -1;
  Sp2FlagIter it(thrust::counting_iterator<int>(0), Sp2HeadFlag{s.skeys});
  bytes = cub_bytes;
  if (cub::DeviceScan::InclusiveSum(cub_temp, bytes, it, s.uidx, (int)n, st) != cudaSuccess) return -1;
  k_sp2_fill<<<g, 256, 0, st>>>(s);
  return 0;
}

// ------------------------------------------------------------------------------------------ walk
template <int CH, int MODE>
__device__ __forceinline__ void sp2_flush(const Sp2& s, const AdamP& a, int t, int uf, float scale, int rank, int key, const float4& acc,
                                          int c, bool shared) {
  constexpr int W = CH * 4;
  const int ul = rank - 1 - uf;
  if (MODE == SP2_COMPACT) {
    float* p = s.accum[t] + (int64_t)ul * W + 4 * c;
    if (shared) red_add4(p, acc); else st4(p, acc);
  } else if (MODE == SP2_DENSE) {
    const int id = key - sp2_base(s, t);
    float* p = s.rep_grad + (t == 0 ? 0 : (int64_t)s.n_items * kI) + (int64_t)id * W + 4 * c;
    if (shared) red_add4(p, acc); else st4(p, acc);
    if (c == 0) s.rep_grad[(int64_t)s.n_items * kI + (int64_t)s.n_cates * kC + key] = 1.0f;
  } else {
    if (shared) {
      red_add4(s.accum[t] + (int64_t)ul * W + 4 * c, acc);
      if (c == 0) s.dflag[t][ul] = 1;
    } else {
      const int id = key - sp2_base(s, t);
      sp2_row_chunk(s.w[t], s.m[t], s.v[t], (int64_t)id * W + 4 * c, acc, scale, a);
    }
  }
}

template <int CH, int MODE>
__device__ __forceinline__ void sp2_walk_warp(const Sp2& s, const AdamP& a, int t, int64_t wb, float4* s_acc, int* s_rank, int* s_key) {
  constexpr int NG = 32 / CH, WB = NG * kSp2Sub;
  const int lane = threadIdx.x & 31, grp = lane / CH, c = lane % CH;
  const int64_t NI = s.N + s.B;
  const int64_t r0 = (int64_t)t * NI, r1 = r0 + NI;              // table t's positions, sorted and unsorted alike
  const int64_t w0 = r0 + wb * WB;
  if (w0 >= r1) return;                                          // warp-uniform
  const int64_t w1 = w0 + WB < r1 ? w0 + WB : r1;
  const int64_t p0 = w0 + (int64_t)grp * kSp2Sub;
  const int64_t p1 = p0 + kSp2Sub < w1 ? p0 + kSp2Sub : w1;
  const int col0 = t == 0 ? 0 : kI;
  // every index of the group's 16 positions first (coalesced), then the gradient rows in two batches of 8 independent loads
  int rk[kSp2Sub], ky[kSp2Sub];
  const float* row[kSp2Sub];
#pragma unroll
  for (int i = 0; i < kSp2Sub; ++i) {
    const int64_t p = p0 + i;
    const bool ok = p < p1;
    rk[i] = ok ? s.uidx[p] : -1;
    ky[i] = (MODE != SP2_COMPACT && ok) ? s.skeys[p] : 0;
    const int64_t src = ok ? (int64_t)s.sidx[p] - r0 : 0;
    row[i] = (src < s.N ? s.dX0 + src * kD + col0 : s.dT + (src - s.N) * kE + col0) + 4 * c;
  }
  const bool left = w0 > r0 && s.uidx[w0 - 1] == s.uidx[w0];
  const bool right = w1 < r1 && s.uidx[w1] == s.uidx[w1 - 1];
  const int uf = s.meta[t];
  const float scale = MODE == SP2_FUSED ? sp2_clip_scale(s, a, t) : 1.0f;
  int cur_rank = -1, cur_key = 0, n_runs = 0;
  float4 acc = f4_zero();
#pragma unroll
  for (int h = 0; h < kSp2Sub; h += 8) {
    float4 g[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) g[j] = rk[h + j] >= 0 ? ld4(row[h + j]) : f4_zero();
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int r = rk[h + j];
      if (r < 0) continue;
      if (r != cur_rank) {
        if (cur_rank >= 0) {
          if (n_runs == 0) {                                     // first run of the group: may continue the previous group's
            s_acc[(2 * grp) * CH + c] = acc;
            if (c == 0) { s_rank[2 * grp] = cur_rank; s_key[2 * grp] = cur_key; }
          } else {
            sp2_flush<CH, MODE>(s, a, t, uf, scale, cur_rank, cur_key, acc, c, false);   // strictly inside the group: final
          }
          ++n_runs;
        }
        cur_rank = r; cur_key = ky[h + j]; acc = f4_zero();
      }
      f4_add(acc, g[j]);
    }
  }
  {
    const int sl = n_runs == 0 ? 2 * grp : 2 * grp + 1;          // the last run may continue into the next group
    if (cur_rank >= 0) s_acc[sl * CH + c] = acc;
    if (c == 0) {
      s_rank[sl] = cur_rank; s_key[sl] = cur_key;
      if (n_runs == 0) s_rank[2 * grp + 1] = -1;
    }
  }
  __syncwarp();
  if (lane < CH) {
    int mr = -1, mk = 0;
    bool first = true;
    float4 ma = f4_zero();
#pragma unroll 4
    for (int sl = 0; sl < 2 * NG; ++sl) {
      const int r = s_rank[sl];
      if (r < 0) continue;
      const float4 e = s_acc[sl * CH + c];
      if (r != mr) {
        if (mr >= 0) { sp2_flush<CH, MODE>(s, a, t, uf, scale, mr, mk, ma, c, first && left); first = false; }
        mr = r; mk = s_key[sl]; ma = e;
      } else {
        f4_add(ma, e);
      }
    }
    if (mr >= 0) sp2_flush<CH, MODE>(s, a, t, uf, scale, mr, mk, ma, c, (first && left) || right);
  }
  __syncwarp();
}

template <int MODE>
__global__ void __launch_bounds__(kSp2Warps * 32) k_sp2_walk(const Sp2 s, const AdamP a, int item_ctas) {
  __shared__ float4 s_acc[kSp2Warps][64];                        // 2 * NG * CH = 64 float4 per warp for both tables
  __shared__ int s_rank[kSp2Warps][64];
  __shared__ int s_key[kSp2Warps][64];
  const int w = threadIdx.x >> 5;
  if ((int)blockIdx.x < item_ctas)
    sp2_walk_warp<4, MODE>(s, a, 0, (int64_t)blockIdx.x * kSp2Warps + w, s_acc[w], s_rank[w], s_key[w]);
  else
    sp2_walk_warp<1, MODE>(s, a, 1, (int64_t)(blockIdx.x - item_ctas) * kSp2Warps + w, s_acc[w], s_rank[w], s_key[w]);
}

void launch_sp2_walk(const Sp2& s, const AdamP& a, cudaStream_t st) {
  PAMREC_PROF("sparse_walk", 1, st);
  const int64_t NI = s.N + s.B;
  if (NI == 0) return;
  const int item_ctas = (int)((NI + (int64_t)sp2_wb(4) * kSp2Warps - 1) / ((int64_t)sp2_wb(4) * kSp2Warps));
  const int cate_ctas = (int)((NI + (int64_t)sp2_wb(1) * kSp2Warps - 1) / ((int64_t)sp2_wb(1) * kSp2Warps));
  const unsigned grid = (unsigned)(item_ctas + cate_ctas);
  if (s.mode == SP2_COMPACT) k_sp2_walk<SP2_COMPACT><<<grid, kSp2Warps * 32, 0, st>>>(s, a, item_ctas);
  else if (s.mode == SP2_FUSED) k_sp2_walk<SP2_FUSED><<<grid, kSp2Warps * 32, 0, st>>>(s, a, item_ctas);
  else k_sp2_walk<SP2_DENSE><<<grid, kSp2Warps * 32, 0, st>>>(s, a, item_ctas);
}

// ------------------------------------------------------------------------------------------ apply
// LAZY: the rows the walk could not finish (runs across warp blocks) and the user rows (L2 gradient only)
__global__ void __launch_bounds__(256) k_sp2_lazy_finish(const Sp2 s, const AdamP a) {
  const int u = blockIdx.x * blockDim.x + threadIdx.x;
  const int u1 = s.meta[1], u2 = s.meta[2], u3 = s.meta[3];
  if (u >= u3) return;
  const int t = u >= u2 ? 2 : (u >= u1 ? 1 : 0);
  const int ul = u - (t == 2 ? u2 : (t == 1 ? u1 : 0));
  const int id = s.ukeys[t][ul];
  if (t == 2) {
    const float sl = sp2_clip_scale(s, a, 2), ss = sp2_clip_scale(s, a, 3);
#pragma unroll
    for (int c = 0; c < 5; ++c) {
      sp2_row_chunk(s.w[2], s.m[2], s.v[2], (int64_t)id * PAMREC_USER_DIM + 4 * c, f4_zero(), sl, a);
      sp2_row_chunk(s.w[3], s.m[3], s.v[3], (int64_t)id * PAMREC_USER_DIM + 4 * c, f4_zero(), ss, a);
    }
    return;
  }
  if (!s.dflag[t][ul]) return;
  s.dflag[t][ul] = 0;
  const float scale = sp2_clip_scale(s, a, t);
  if (t == 0) {
#pragma unroll
    for (int c = 0; c < 4; ++c) sp2_row_chunk(s.w[0], s.m[0], s.v[0], (int64_t)id * kI + 4 * c, ld4(s.accum[0] + (int64_t)ul * kI + 4 * c), scale, a);
  } else {
    sp2_row_chunk(s.w[1], s.m[1], s.v[1], (int64_t)id * kC, ld4(s.accum[1] + (int64_t)ul * kC), scale, a);
  }
}
void launch_sp2_lazy_finish(const Sp2& s, const AdamP& a, cudaStream_t st) {
  PAMREC_PROF("sparse_adam", 1, st);
  const int64_t n = 2 * (s.N + s.B) + s.B;                        // upper bound of the unique count (it lives on the device)
  if (n == 0) return;
  k_sp2_lazy_finish<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(s, a);
}

// Full sweep over the four tables in one launch (the HBM-bound kernel: 6 x row bytes per row + 4 B of slot / touch word).
// COMPACT: gradient of a touched row = accum[slot[row]];  DENSE (replicated data parallel): gradient table and touch counts,
// already summed over the ranks.  LAZY skips untouched rows (replicated LAZY).  Every CTA works on ONE table (its share of the
// grid is proportional to the table's bytes), so the loop body is the plain one-table sweep: w, m, v loads issued first, the
// slot / touch word and the gradient row behind them.
template <int CH, int MODE, bool LAZY>
__device__ __forceinline__ void sp2_sweep_table(const Sp2& s, const AdamP& a, int t, int key_base, int64_t n_rows, const float* grad,
                                                const float* touch, float scale, int64_t cta, int64_t n_cta) {
  const int64_t total = n_rows * CH;
  float* tw = s.w[t]; float* tm = s.m[t]; float* tv = s.v[t];
  for (int64_t i = cta * blockDim.x + threadIdx.x; i < total; i += n_cta * blockDim.x) {
    const int64_t r = i / CH;
    const int c = (int)(i % CH);
    float4 w, m, v;
    if (!LAZY) { w = ld4(tw + 4 * i); m = ld4(tm + 4 * i); v = ld4(tv + 4 * i); }
    bool touched;
    float4 g = f4_zero();
    if (MODE == SP2_DENSE) {
      touched = __ldg(touch + key_base + r) > 0.f;
      if (touched && grad != nullptr) g = ld4(grad + 4 * i);
    } else {
      const int u = __ldg(s.slot + key_base + r);
      touched = u >= 0;
      if (touched && grad != nullptr) g = ld4(grad + ((int64_t)u * CH + c) * 4);
    }
    if (LAZY) {
      if (!touched) continue;
      w = ld4(tw + 4 * i); m = ld4(tm + 4 * i); v = ld4(tv + 4 * i);
    }
    if (touched) {
      g.x = scale * (g.x + a.l2 * w.x); g.y = scale * (g.y + a.l2 * w.y); g.z = scale * (g.z + a.l2 * w.z); g.w = scale * (g.w + a.l2 * w.w);
    }
    sp2_adam4(w, m, v, g, a);
    st4(tw + 4 * i, w); st4(tm + 4 * i, m); st4(tv + 4 * i, v);
  }
}
struct Sp2SweepGrid { int g[4]; };                       // CTAs per table
template <int MODE, bool LAZY>
__global__ void __launch_bounds__(256, 8) k_sp2_adam_sweep(const Sp2 s, const AdamP a, const Sp2SweepGrid sg) {
  int b = blockIdx.x, t = 0;
  while (t < 3 && b >= sg.g[t]) { b -= sg.g[t]; ++t; }
  const float scale = sp2_clip_scale(s, a, t);
  const float* touch = MODE == SP2_DENSE ? s.rep_grad + (int64_t)s.n_items * kI + (int64_t)s.n_cates * kC : nullptr;
  const float* gi = MODE == SP2_DENSE ? s.rep_grad : s.accum[0];
  const float* gc = MODE == SP2_DENSE ? s.rep_grad + (int64_t)s.n_items * kI : s.accum[1];
  if (t == 0) sp2_sweep_table<4, MODE, LAZY>(s, a, 0, 0, s.n_items, gi, touch, scale, b, sg.g[0]);
  else if (t == 1) sp2_sweep_table<1, MODE, LAZY>(s, a, 1, s.n_items, s.n_cates, gc, touch, scale, b, sg.g[1]);
  else sp2_sweep_table<5, MODE, LAZY>(s, a, t, s.n_items + s.n_cates, s.n_users, nullptr, touch, scale, b, sg.g[t]);
}
void launch_sp2_adam_sweep(const Sp2& s, const AdamP& a, int lazy, cudaStream_t st) {
  PAMREC_PROF("sparse_adam", 1, st);
  const int64_t chunks[4] = {(int64_t)s.n_items * 4, (int64_t)s.n_cates, (int64_t)s.n_users * 5, (int64_t)s.n_users * 5};
  const int64_t total = chunks[0] + chunks[1] + chunks[2] + chunks[3];
  const int64_t cap = 148 * 32;                                 // grid-stride above 32 CTAs per SM
  Sp2SweepGrid sg;
  int grid = 0;
  for (int t = 0; t < 4; ++t) {
    const int64_t blocks = (chunks[t] + 255) / 256;
    int64_t share = total > cap * 256 ? (chunks[t] * cap + total - 1) / total : blocks;   // proportional to the table's size
    if (share > blocks) share = blocks;
    if (share < 1) share = 1;
    sg.g[t] = (int)share;
    grid += sg.g[t];
  }
  if (s.mode == SP2_DENSE) {
    if (lazy) k_sp2_adam_sweep<SP2_DENSE, true><<<grid, 256, 0, st>>>(s, a, sg);
    else k_sp2_adam_sweep<SP2_DENSE, false><<<grid, 256, 0, st>>>(s, a, sg);
  } else {
    k_sp2_adam_sweep<SP2_COMPACT, false><<<grid, 256, 0, st>>>(s, a, sg);
  }
}

// replicated data parallel: |w|^2 of the rows ANY rank looked up (touch counts summed over the ranks) = the L2 rows of
// tf.unique over the global batch
__global__ void __launch_bounds__(256) k_sp2_rep_l2(const Sp2 s) {
  __shared__ double sh[4][8];
  const int64_t nk = (int64_t)s.n_items + s.n_cates + s.n_users;
  const int64_t key = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const float* touch = s.rep_grad + (int64_t)s.n_items * kI + (int64_t)s.n_cates * kC;
  double ss[4] = {0.0, 0.0, 0.0, 0.0};
  if (key < nk && touch[key] > 0.f) {
    if (key < s.n_items) {
      const float* w = s.w[0] + key * kI; float q = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) { const float4 x = ld4(w + 4 * c); q += f4_dot(x, x); }
      ss[0] = (double)q;
    } else if (key < s.n_items + s.n_cates) {
      const float4 x = ld4(s.w[1] + (key - s.n_items) * kC); ss[1] = (double)f4_dot(x, x);
    } else {
      const int64_t id = key - s.n_items - s.n_cates;
      const float* p = s.w[2] + id * PAMREC_USER_DIM; const float* q = s.w[3] + id * PAMREC_USER_DIM;
      float qa = 0.f, qb = 0.f;
#pragma unroll
      for (int c = 0; c < 5; ++c) { const float4 x = ld4(p + 4 * c), y = ld4(q + 4 * c); qa += f4_dot(x, x); qb += f4_dot(y, y); }
      ss[2] = (double)qa; ss[3] = (double)qb;
    }
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 4; ++k) { const double v = warp_sum_d(ss[k]); if (lane == 0) sh[k][w] = v; }
  __syncthreads();
  if (threadIdx.x < 4) {
    double tot = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) tot += sh[threadIdx.x][k];
    if (tot != 0.0) atomicAdd(s.l2sq + threadIdx.x, tot);
  }
}
void launch_sp2_rep_l2(const Sp2& s, cudaStream_t st) {
  PAMREC_PROF("sparse_l2norm", 1, st);
  const int64_t nk = (int64_t)s.n_items + s.n_cates + s.n_users;
  k_sp2_rep_l2<<<(unsigned)((nk + 255) / 256), 256, 0, st>>>(s);
}

}  // namespace pamrec
