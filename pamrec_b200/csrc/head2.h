// Row-stationary head kernels (kernels_head2.cu): pointers that never change after pamrec_bind (Head2); per-call values travel
// in HeadDyn (headcoop.h).
#pragma once
#include "headcoop.h"
#include "layout.h"

namespace pamrec {

constexpr int kHead2Threads = 512;
constexpr int kHead2Groups = 8;            // copies of the batch-norm sums (atomic contention: 148 CTAs / 8 per address)
constexpr int kHead2WtFloats = 3 * 6400 + 3 * 8400 + 5 * 6400 + 2 * 320 + 628 * 40;   // transposed weights of the dX chain

struct Head2 {
  const float* H;                       // encoder output [B*T,40]
  const float* tgt;                     // target embedding [B,20]
  // dense parameters (group-strided, layout.h) and their gradients
  const float *s_w0, *s_b0, *s_w1, *s_b1;                      // score MLP 40 -> 20 -> 1
  const float *e_w0, *e_b0, *e_w1, *e_b1;                      // 5 experts 40 -> 100 -> 64
  const float *g_w0, *g_b0, *g_w1, *g_b1;                      // 2 gates 40 -> 64 -> 5
  const float *t_w0, *t_b0, *t_w1, *t_b1, *t_wo, *t_bo;        // 3 towers 84 -> 100 -> 64 -> 1
  float *ds_w0, *ds_b0, *ds_w1, *ds_b1, *de_w0, *de_b0, *de_w1, *de_b1, *dg_w0, *dg_b0, *dg_w1, *dg_b1;
  float *dt_w0, *dt_b0, *dt_w1, *dt_b1, *dt_wo, *dt_bo;
  float* wT;                            // workspace [kHead2WtFloats]
  BnSet bn[BN_COUNT];
  double* gsums[BN_COUNT];              // [kHead2Groups][C][2] copies of the forward sums, zero between steps
  double* gbsums[BN_COUNT];             // ... of the backward sums
  // activations / gradients (workspace)
  float *z1, *z2, *aw, *new_long, *ze0, *zg0, *ze1, *zg1, *u, *zt0, *zt1, *logits;
  float *d_logits, *d_t1, *d_t0, *d_e1, *d_g1, *d_e0, *d_g0, *d_new_long, *d_tgt, *d_z2, *g_a;
  double* loss_acc;
  double* dp_scalars;
};

int head2_grid();                        // CTAs of the persistent grid (one per SM), < 0 if cooperative launch is unavailable
int launch_head2_fwd(const Head2& h, const HeadDyn& d, int grid, const char* name, cudaStream_t st);
int launch_head2_bwd(const Head2& h, const HeadDyn& d, int grid, cudaStream_t st);
void launch_head2_dw(const Head2& h, int B, cudaStream_t st);

}  // namespace pamrec
