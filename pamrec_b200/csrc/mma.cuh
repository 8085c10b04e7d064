// Warp-level tensor-core helpers for the token-parallel encoder kernels.
//
// The contractions of the encoder are [tokens x 40] x [40 x 40] (projection, FFN) and their transposes.  fp32 parity
// with the reference (1e-5 on logits and loss) rules out plain TF32 / BF16 operands, so every product is issued as
// the error-compensated 3xTF32 split   a*b ~= a_hi*b_hi + a_lo*b_hi + a_hi*b_lo   (a_hi = tf32(a), a_lo = tf32(a - a_hi)),
// accumulated in fp32 by mma.sync.m16n8k8.  The dropped a_lo*b_lo term is ~2^-22 relative.
//
// Fragment layout of mma.m16n8k8 (g = lane / 4, t = lane % 4):
//   A (16x8, row major):  a0 = A[g][t]   a1 = A[g+8][t]   a2 = A[g][t+4]   a3 = A[g+8][t+4]
//   B ( 8x8, k x n)     :  b0 = B[t][g]   b1 = B[t+4][g]
//   C (16x8)            :  c0 = C[g][2t]  c1 = C[g][2t+1]  c2 = C[g+8][2t]  c3 = C[g+8][2t+1]
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pamrec {

__device__ __forceinline__ uint32_t to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  hi = to_tf32(x);
  lo = to_tf32(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
// small terms first, then the leading term
__device__ __forceinline__ void mma_3xtf32(float (&d)[4], const uint32_t (&ahi)[4], const uint32_t (&alo)[4],
                                           const uint32_t (&bhi)[2], const uint32_t (&blo)[2]) {
  mma_tf32(d, alo, bhi);
  mma_tf32(d, ahi, blo);
  mma_tf32(d, ahi, bhi);
}

// C[16 MT x 40] += A[16 MT x 40] * B[40 x 40] for one warp.
//   a_elem(r, k): element of the warp's A operand (r in 0 .. 16 MT - 1), evaluated once per element;
//   Bhi / Blo   : the 40 x 40 B operand pre-split into TF32 halves, row major with leading dimension 40;
//                 TRANS_B multiplies by B^T instead (B[n][k] is read where B[k][n] would be).
//   c[mt][nt][4]: accumulator fragments, row tile mt (16 rows), column tile nt (8 columns).
template <int MT, bool TRANS_B, typename AF>
__device__ __forceinline__ void warp_gemm_rows_x40x40(float (&c)[MT][5][4], AF a_elem, const uint32_t* __restrict__ Bhi,
                                                      const uint32_t* __restrict__ Blo, int lane) {
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int k0 = 0; k0 < 5; ++k0) {
    const int k = 8 * k0;
    uint32_t ahi[MT][4], alo[MT][4];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      const int r = 16 * mt + g;
      split_tf32(a_elem(r, k + t), ahi[mt][0], alo[mt][0]);
      split_tf32(a_elem(r + 8, k + t), ahi[mt][1], alo[mt][1]);
      split_tf32(a_elem(r, k + t + 4), ahi[mt][2], alo[mt][2]);
      split_tf32(a_elem(r + 8, k + t + 4), ahi[mt][3], alo[mt][3]);
    }
#pragma unroll
    for (int nt = 0; nt < 5; ++nt) {
      const int n = 8 * nt + g;
      const int i0 = TRANS_B ? n * 40 + k + t : (k + t) * 40 + n;
      const int i1 = TRANS_B ? i0 + 4 : i0 + 4 * 40;
      const uint32_t bh[2] = {Bhi[i0], Bhi[i1]};
      const uint32_t bl[2] = {Blo[i0], Blo[i1]};
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) mma_3xtf32(c[mt][nt], ahi[mt], alo[mt], bh, bl);
    }
  }
}

// C[48 x 40] += A^T * B contracted over `n_rows` (multiple of 8) tile rows, for weight gradients:
//   a_elem(r, i): A[r][i] for i < 40, the constant 1 for i == 40 (row 40 of C becomes the column sum of B: the bias
//                 gradient) and 0 above;   b_elem(r, n): B[r][n].
// c[mt][nt][4]: row tile mt covers output rows 16mt .. 16mt+15 (rows 41..47 are padding).
template <typename AF, typename BF>
__device__ __forceinline__ void warp_gemm_tn_48x40(float (&c)[3][5][4], int n_rows, AF a_elem, BF b_elem, int lane) {
  const int g = lane >> 2, t = lane & 3;
  for (int r0 = 0; r0 < n_rows; r0 += 8) {
    uint32_t bhi[5][2], blo[5][2];
#pragma unroll
    for (int nt = 0; nt < 5; ++nt) {
      split_tf32(b_elem(r0 + t, 8 * nt + g), bhi[nt][0], blo[nt][0]);
      split_tf32(b_elem(r0 + t + 4, 8 * nt + g), bhi[nt][1], blo[nt][1]);
    }
#pragma unroll
    for (int mt = 0; mt < 3; ++mt) {
      uint32_t ahi[4], alo[4];
      const int i0 = 16 * mt + g;
      split_tf32(a_elem(r0 + t, i0), ahi[0], alo[0]);
      split_tf32(a_elem(r0 + t, i0 + 8), ahi[1], alo[1]);
      split_tf32(a_elem(r0 + t + 4, i0), ahi[2], alo[2]);
      split_tf32(a_elem(r0 + t + 4, i0 + 8), ahi[3], alo[3]);
#pragma unroll
      for (int nt = 0; nt < 5; ++nt) mma_3xtf32(c[mt][nt], ahi, alo, bhi[nt], blo[nt]);
    }
  }
}
// add rows 0 .. last_row of such an accumulator into a row-major [last_row + 1][40] shared-memory buffer
__device__ __forceinline__ void tn_flush_smem(const float (&c)[3][5][4], float* __restrict__ acc, int lane, int last_row = 40) {
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int mt = 0; mt < 3; ++mt)
#pragma unroll
    for (int nt = 0; nt < 5; ++nt) {
      const int r = 16 * mt + g, col = 8 * nt + 2 * t;
      if (r <= last_row) { atomicAdd(acc + r * 40 + col, c[mt][nt][0]); atomicAdd(acc + r * 40 + col + 1, c[mt][nt][1]); }
      if (r + 8 <= last_row) { atomicAdd(acc + (r + 8) * 40 + col, c[mt][nt][2]); atomicAdd(acc + (r + 8) * 40 + col + 1, c[mt][nt][3]); }
    }
}

}  // namespace pamrec
