// Program format of the persistent cooperative head kernels (kernels_headcoop.cu).  Built once per handle on the host
// (api.cu: pointers into the bound pools and the workspace never change after pamrec_bind), kept in device memory; everything
// that changes from call to call (batch size, batch pointers, training flag, barrier / mailbox epochs) travels in HeadDyn.
#pragma once
#include "kernels.h"

namespace pamrec {

constexpr int kHeadCtasPerSm = 4;          // resident CTAs per SM of the persistent grid (leaves room for side-stream kernels)
constexpr int kHeadMaxPhases = 28;
constexpr int BN_S1_ = 1, BN_E1_ = 4, BN_G1_ = 5, BN_COUNT_ = 8;   // mirror layout.h:BnId (static_assert in api.cu)

enum HeadOp { HEAD_OP_NONE = 0, HEAD_OP_DENSE_FWD, HEAD_OP_DENSE_DX, HEAD_OP_DENSE_DW, HEAD_OP_POOL_FWD, HEAD_OP_POOL_BWD,
              HEAD_OP_COMBINE_FWD, HEAD_OP_COMBINE_BWD, HEAD_OP_LOSS };

struct PoolP { const float* H; const float* Z2; float* new_long; const float* dNL; float* dA2; float* dH; };
struct CombineP { const float* ZE1; const float* ZG1; const float* tgt; float* U; const float* dU; float* dE1; float* dG1; float* dTgt; };
struct LossP { const float* logits; float* d_logits; double* loss_acc; const double* n_valid_global; };

struct alignas(16) HeadPhase {
  int op;
  int rows_n;            // 1: the phase runs over B*T rows (score MLP), 0: over B rows
  int barrier;           // grid-wide barrier after this phase
  int only;              // 0 always, 1 training only, 2 scoring only
  // leader section of the barrier
  int n_fin, fin[2], fin_rows_n;        // forward / training: batch-norm sets whose statistics are finalised (count = B*T or B rows)
  int n_sync, sync[2], sync_bwd;        // data parallel: sets whose forward sums (sync_bwd = 0) / backward sums (1) are all-reduced
  int sync_scalars;                     // ... plus the 8 data-parallel scalars (listwise groups with a non-zero label sum)
  int eval_stats;                       // scoring: (mean, invstd) of every set from the moving statistics
  union U {
    DenseP f; DenseDxP x; DenseDwP w; PoolP pl; CombineP cb; LossP ls;
    __host__ __device__ U() {}
  } u;
};

struct HeadProgram {
  int n;
  BnSet bn[BN_COUNT_];
  double* dp_scalars;
  HeadPhase ph[kHeadMaxPhases];
};

struct HeadDyn {
  int B, T, Bg, training, world, rank;
  double cntN, cntB;
  const int* mask; const float* y_sat; const float* y_play; const float* plays;
  float fuzhu_w, order_w; int sm_group;
  unsigned* bar;                                    // [0] arrivals [2] error word [32] release epoch (own 128-byte line)
  double* peer_slots[kP2PMaxWorld]; uint32_t* peer_flags[kP2PMaxWorld];
  uint32_t p2p_epoch; int p2p_slot0; uint32_t* p2p_err;
  uint4* peer_ll[kP2PMaxWorld];                     // flag-in-data mailboxes [slot][rank][kP2PMaxDoubles] (kernels.h:p2p_ll_offset)
  unsigned long long* trace;                        // [32] %globaltimer of the kernel start ([31]) and of every barrier release; may be null
  float* pred;                                      // scoring (row-stationary kernels): sigmoid(logit 0) per row, or null
  unsigned long long* trace_cta;                    // debug: [16 barriers][256 CTAs] %globaltimer of every CTA's arrival; may be null
};

int head_program_grid(int* ctas_per_sm_out);        // CTAs of the persistent grid on the current device, < 0 if unsupported
int launch_head_program(const HeadProgram* dev_prog, const HeadDyn& d, int grid, const char* name, cudaStream_t st);

}  // namespace pamrec
