// Attention forward on the 5th-generation tensor cores: tcgen05.mma (kind::tf32) with accumulators in tensor memory.
//
//   S = Q K^T / sqrt(40),  P = softmax_j(mask_j ? S : -(2^32)+1),  y = P V + qin        (pamrec.py:768-810)
//
// One CTA of 128 threads owns a UNIT: two samples when T <= 64 (each padded to 64 rows / keys of a 128-row tile, the
// off-diagonal blocks of S are discarded), otherwise one sample with one or two 128-row query tiles.  Per query tile:
//   1. Q, K, V of the unit are staged in shared memory as TF32 hi / lo halves (3xTF32 error-compensated split, see
//      mma.cuh: a*b ~= a_hi*b_hi + a_lo*b_hi + a_hi*b_lo) in the canonical K-major no-swizzle UMMA layout: 16-byte
//      chunks of 4 consecutive k, chunk-major, so that a core matrix (8 rows x 16 bytes) is 128 contiguous bytes;
//   2. one elected thread issues 15 tcgen05.mma (5 k-steps of 8 x 3 split terms): S[128 x NK] lands in TMEM columns
//      [0, NK); tcgen05.commit arrives on an mbarrier;
//   3. thread r reads row r of S from TMEM (tcgen05.ld 32x32b), applies the key mask and an exact two-pass softmax,
//      and writes the un-normalised P back to TMEM as hi (in place of S) and lo (columns [256, 256 + NK)) halves;
//   4. P V as 3 x NK/8 tcgen05.mma with the A operand read FROM TENSOR MEMORY and V^T (K-major over the keys) from
//      shared memory, accumulating O[128 x 48] in TMEM columns [464, 512);
//   5. y = O / l + qin, plus the row statistics (m, l) the backward pass needs.
// The FFMA kernel (kernels_encoder.cu:k_attn_fwd) remains for T > 208 (S, P_lo and O no longer fit 512 TMEM columns)
// and as the PAMREC_ATTN=ffma reference path.
#include "kernels.h"
#include "mma.cuh"

namespace pamrec {

namespace tc {

constexpr int kThreads = 128;
constexpr int kRows = 128;                 // query rows per tile = TMEM lanes
constexpr int kChunks = kD / 4;            // 16-byte k-chunks of a 40-wide row
constexpr int kNV = 48;                    // N of the P V product: 40 padded to a multiple of 16 (M = 128 needs N % 16 == 0)
constexpr int kMaxNK = 208;
constexpr uint32_t kColPlo = 256, kColO = 464, kTmemCols = 512;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
// bounded wait: a tensor-core op that never completes becomes an error word, not a hung GPU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, int* err) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 22)) { if (err) atomicExch(err, 1); break; }
  }
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, both operands K-major
__device__ __forceinline__ void umma_ss(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]^T
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
               ::"r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
// shared-memory matrix descriptor, K-major, no swizzle (cute/arch/mma_sm100_desc.hpp: SmemDescriptor): start address, leading
// (K direction: between the two 16-byte chunks of one k-step) and stride (M / N direction: between 8-row core matrices) byte
// offsets, all in 16-byte units; version 1 (Blackwell); layout type 0 = SWIZZLE_NONE.
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
// instruction descriptor (InstrDescriptor): D = F32, A = B = TF32, both K-major, N >> 3 at bit 17, M >> 4 at bit 24
__device__ __forceinline__ uint32_t make_idesc(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void tmem_ld16(uint32_t addr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                 "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(addr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_st16(uint32_t addr, const uint32_t (&r)[16]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
               ::"r"(addr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
                 "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}

struct Geometry {
  int spt;        // samples per unit (2 when T <= 64)
  int qtiles;     // 128-row query tiles per unit
  int nk;         // key columns per unit (multiple of 16)
};
__host__ __device__ inline Geometry geometry(int T) {
  Geometry g;
  if (T <= 64) { g.spt = 2; g.qtiles = 1; g.nk = 128; }
  else { g.spt = 1; g.qtiles = (T + kRows - 1) / kRows; g.nk = (T + 15) / 16 * 16; }
  return g;
}
inline size_t smem_bytes(int T) {
  const Geometry g = geometry(T);
  // Qh | Ql | Kh | Kl | Vh | Vl | key mask | tmem base, mbarrier
  return (size_t)(2 * kChunks * kRows * 4 + 2 * kChunks * g.nk * 4 + 2 * (g.nk / 4) * kNV * 4) * 4 + (size_t)g.nk * 4 + 64 + 1024;
}

}  // namespace tc

__global__ void __launch_bounds__(tc::kThreads, 1)
k_attn_fwd_tc(const float* __restrict__ Q, const float* __restrict__ K, const float* __restrict__ V, const float* __restrict__ QIN,
              const int* __restrict__ mask, float* __restrict__ Y, float* __restrict__ ML, int B, int T, int n_units, int* __restrict__ err) {
  using namespace tc;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const Geometry g = geometry(T);
  const int NK = g.nk;
  float* Qh = reinterpret_cast<float*>(base);                  // [kChunks][128][4]
  float* Ql = Qh + kChunks * kRows * 4;
  float* Kh = Ql + kChunks * kRows * 4;                        // [kChunks][NK][4]
  float* Kl = Kh + kChunks * NK * 4;
  float* Vh = Kl + kChunks * NK * 4;                           // [NK / 4][48][4]   (V^T: rows = the 40 (+8) output dims, k = keys)
  float* Vl = Vh + (NK / 4) * kNV * 4;
  int* kmask = reinterpret_cast<int*>(Vl + (NK / 4) * kNV * 4);   // [NK]  1 = valid key, 0 = masked key, -1 = no key (padding)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(kmask + NK);
  uint64_t* bar = reinterpret_cast<uint64_t*>(tmem_slot + 2);
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);      // this warp's 32 TMEM lanes
  const uint32_t idesc_s = make_idesc(kRows, NK), idesc_o = make_idesc(kRows, kNV);
  const float rscale = 1.0f / sqrtf((float)kD);
  uint32_t phase = 0;

  for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
    const int b0 = unit * g.spt;
    // ---- stage K (hi / lo), V^T (hi / lo) and the key mask of the unit
    for (int i = tid; i < NK * kChunks; i += kThreads) {
      const int kk = i % NK, c = i / NK;                       // lanes over consecutive keys: conflict-free 16-byte stores
      const int sl = g.spt == 2 ? kk >> 6 : 0, t = g.spt == 2 ? (kk & 63) : kk;
      const bool ok = t < T && b0 + sl < B;
      float4 kv = f4_zero(), vv = f4_zero();
      if (ok) {
        const int64_t gofs = ((int64_t)(b0 + sl) * T + t) * kD + 4 * c;
        kv = ld4(K + gofs);
        vv = ld4(V + gofs);
      }
      uint4 h, l;
      split_tf32(kv.x, h.x, l.x); split_tf32(kv.y, h.y, l.y); split_tf32(kv.z, h.z, l.z); split_tf32(kv.w, h.w, l.w);
      *reinterpret_cast<uint4*>(Kh + (c * NK + kk) * 4) = h;
      *reinterpret_cast<uint4*>(Kl + (c * NK + kk) * 4) = l;
      // V^T: element (n = 4c + e, key kk) at chunk kk / 4, row n, slot kk % 4
      const float ve[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        uint32_t vh, vl;
        split_tf32(ve[e], vh, vl);
        const int o = ((kk >> 2) * kNV + 4 * c + e) * 4 + (kk & 3);
        reinterpret_cast<uint32_t*>(Vh)[o] = vh;
        reinterpret_cast<uint32_t*>(Vl)[o] = vl;
      }
    }
    for (int i = tid; i < (NK / 4) * (kNV - kD) * 4; i += kThreads) {   // rows 40..47 of V^T: zero
      const int ch = i / ((kNV - kD) * 4), rem = i % ((kNV - kD) * 4);
      const int o = (ch * kNV + kD) * 4 + rem;
      Vh[o] = 0.f; Vl[o] = 0.f;
    }
    for (int kk = tid; kk < NK; kk += kThreads) {
      const int sl = g.spt == 2 ? kk >> 6 : 0, t = g.spt == 2 ? (kk & 63) : kk;
      kmask[kk] = (t < T && b0 + sl < B) ? (mask[(int64_t)(b0 + sl) * T + t] != 0 ? 1 : 0) : -1;
    }
    for (int qt = 0; qt < g.qtiles; ++qt) {
      // ---- stage the query tile
      for (int i = tid; i < kRows * kChunks; i += kThreads) {
        const int r = i % kRows, c = i / kRows;
        const int sl = g.spt == 2 ? r >> 6 : 0, t = g.spt == 2 ? (r & 63) : qt * kRows + r;
        float4 qv = f4_zero();
        if (t < T && b0 + sl < B) qv = ld4(Q + ((int64_t)(b0 + sl) * T + t) * kD + 4 * c);
        uint4 h, l;
        split_tf32(qv.x, h.x, l.x); split_tf32(qv.y, h.y, l.y); split_tf32(qv.z, h.z, l.z); split_tf32(qv.w, h.w, l.w);
        *reinterpret_cast<uint4*>(Qh + (c * kRows + r) * 4) = h;
        *reinterpret_cast<uint4*>(Ql + (c * kRows + r) * 4) = l;
      }
      fence_proxy_async_smem();               // generic-proxy writes above -> visible to the tensor core's async proxy
      tc_fence_before();
      __syncthreads();
      tc_fence_after();
      // ---- S = Q K^T : 5 k-steps x (lo*hi, hi*lo, hi*hi)
      if (tid == 0) {
        const uint32_t lbo_q = kRows * 16, lbo_k = (uint32_t)NK * 16, sbo = 128;
        uint32_t acc = 0;
#pragma unroll
        for (int ks = 0; ks < kD / 8; ++ks) {
          const uint64_t dqh = make_desc(smem_u32(Qh) + 2 * ks * lbo_q, lbo_q, sbo), dql = make_desc(smem_u32(Ql) + 2 * ks * lbo_q, lbo_q, sbo);
          const uint64_t dkh = make_desc(smem_u32(Kh) + 2 * ks * lbo_k, lbo_k, sbo), dkl = make_desc(smem_u32(Kl) + 2 * ks * lbo_k, lbo_k, sbo);
          umma_ss(tmem, dql, dkh, idesc_s, acc); acc = 1;
          umma_ss(tmem, dqh, dkl, idesc_s, 1);
          umma_ss(tmem, dqh, dkh, idesc_s, 1);
        }
        tc_commit(bar);
      }
      mbar_wait(bar, phase, err); phase ^= 1;
      tc_fence_after();
      // ---- softmax of row `tid`
      const int r = tid;
      const int sl = g.spt == 2 ? r >> 6 : 0, tq = g.spt == 2 ? (r & 63) : qt * kRows + r;
      const bool row_ok = tq < T && b0 + sl < B;
      const int c_lo = g.spt == 2 ? 64 * sl : 0, c_hi = g.spt == 2 ? 64 * sl + 64 : NK;     // columns that belong to this row's sample
      float m = -INFINITY;
      for (int c0 = c_lo; c0 < c_hi; c0 += 16) {
        float s[16];
        tmem_ld16(lane_base + (uint32_t)c0, s);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int km = kmask[c0 + j];
          if (km >= 0) m = fmaxf(m, km ? s[j] * rscale : kMaskNeg);
        }
      }
      float lsum = 0.f;
      for (int c0 = 0; c0 < NK; c0 += 16) {
        uint32_t ph[16], pl[16];
        if (c0 >= c_lo && c0 < c_hi) {
          float s[16];
          tmem_ld16(lane_base + (uint32_t)c0, s);
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int km = kmask[c0 + j];
            float p = 0.f;
            if (km >= 0) p = expf((km ? s[j] * rscale : kMaskNeg) - m);
            lsum += p;
            split_tf32(p, ph[j], pl[j]);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) { ph[j] = 0u; pl[j] = 0u; }
        }
        tmem_st16(lane_base + (uint32_t)c0, ph);
        tmem_st16(lane_base + kColPlo + (uint32_t)c0, pl);
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      tc_fence_before();
      __syncthreads();
      tc_fence_after();
      // ---- O = P V : NK / 8 k-steps x (lo*hi, hi*lo, hi*hi), A from tensor memory
      if (tid == 0) {
        const uint32_t lbo_v = kNV * 16, sbo = 128;
        uint32_t acc = 0;
        for (int ks = 0; ks < NK / 8; ++ks) {
          const uint64_t dvh = make_desc(smem_u32(Vh) + 2 * ks * lbo_v, lbo_v, sbo), dvl = make_desc(smem_u32(Vl) + 2 * ks * lbo_v, lbo_v, sbo);
          umma_ts(tmem + kColO, tmem + kColPlo + 8 * ks, dvh, idesc_o, acc); acc = 1;
          umma_ts(tmem + kColO, tmem + 8 * ks, dvl, idesc_o, 1);
          umma_ts(tmem + kColO, tmem + 8 * ks, dvh, idesc_o, 1);
        }
        tc_commit(bar);
      }
      mbar_wait(bar, phase, err); phase ^= 1;
      tc_fence_after();
      // ---- y = O / l + qin ; row statistics
      {
        float o[48];
#pragma unroll
        for (int c0 = 0; c0 < kNV; c0 += 16) {
          float v[16];
          tmem_ld16(lane_base + kColO + (uint32_t)c0, v);
#pragma unroll
          for (int j = 0; j < 16; ++j) o[c0 + j] = v[j];
        }
        if (row_ok) {
          const int64_t tok = (int64_t)(b0 + sl) * T + tq;
          const float inv = 1.0f / lsum;
#pragma unroll
          for (int i = 0; i < kChunks; ++i) {
            const float4 q = ld4(QIN + tok * kD + 4 * i);
            st4(Y + tok * kD + 4 * i, make_float4(fmaf(o[4 * i], inv, q.x), fmaf(o[4 * i + 1], inv, q.y), fmaf(o[4 * i + 2], inv, q.z),
                                                  fmaf(o[4 * i + 3], inv, q.w)));
          }
          ML[2 * tok] = m;
          ML[2 * tok + 1] = lsum;
        }
      }
      tc_fence_before();
      __syncthreads();                         // TMEM and the Q tile are free for the next query tile / unit
      tc_fence_after();
    }
  }
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols) : "memory");
}

bool attn_tc_supported(int T) { return T >= 1 && tc::geometry(T).nk <= tc::kMaxNK; }

int init_attn_tc_kernels(int max_T) {
  size_t need = 0;
  for (int T = 1; T <= max_T; ++T)
    if (attn_tc_supported(T)) need = need > tc::smem_bytes(T) ? need : tc::smem_bytes(T);
  if (need == 0) return 0;
  return cudaFuncSetAttribute((const void*)k_attn_fwd_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need) == cudaSuccess ? 0 : -1;
}

void launch_attn_fwd_tc(const float* Q, const float* K, const float* V, const float* QIN, const int* mask, float* Y, float* ML, int B, int T,
                        int n_sm, int* err, cudaStream_t st) {
  PAMREC_PROF("attn_fwd", 1, st);
  if (B == 0) return;
  const tc::Geometry g = tc::geometry(T);
  const int n_units = (B + g.spt - 1) / g.spt;
  const int grid = n_units < n_sm ? n_units : n_sm;          // persistent: one CTA per SM (it owns all 512 TMEM columns)
  k_attn_fwd_tc<<<grid, tc::kThreads, tc::smem_bytes(T), st>>>(Q, K, V, QIN, mask, Y, ML, B, T, n_units, err);
}

}  // namespace pamrec
