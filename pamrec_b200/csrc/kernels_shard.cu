// Row-sharded embedding tables (PAMREC_TABLES_SHARDED): requester-side routing of the unique ids of a batch
// to their owners and owner-side row service.  Row r of a table lives on rank r % W at local row r / W; the
// plan sorts lookups by the composite key  owner * rows_per_shard + local_row,  so the list of unique keys is
// already grouped by owner and each group is the list of local rows to ask that owner for.
//
// No reference counterpart: the reference keeps whole tables in one tf.Session (sequential_base_model.py:572-593).
#include "kernels.h"

namespace pamrec {

// off[o] = first unique index whose key belongs to owner >= o (o = 0..W); counts[4*o + slot] = off[o+1] - off[o]
__global__ void k_owner_offsets(const int* __restrict__ ukeys, const int* __restrict__ nuniq, int world, int64_t rps,
                                int* __restrict__ off, int* __restrict__ counts, int slot) {
  __shared__ int s_off[64 + 1];
  const int o = threadIdx.x;
  const int n = *nuniq;
  if (o <= world) {
    const int64_t bound = (int64_t)o * rps;
    int lo = 0, hi = n;
    while (lo < hi) {
      int mid = (lo + hi) >> 1;
      if ((int64_t)ukeys[mid] < bound) lo = mid + 1; else hi = mid;
    }
    s_off[o] = lo;
    off[o] = lo;
  }
  __syncthreads();
  if (o < world) counts[4 * o + slot] = s_off[o + 1] - s_off[o];
}

// send_ids[u] = local row of unique key u at its owner
__global__ void k_local_rows(const int* __restrict__ ukeys, const int* __restrict__ nuniq, int64_t rps, int* __restrict__ send_ids) {
  const int u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u < *nuniq) send_ids[u] = (int)((int64_t)ukeys[u] % rps);
}

// inv[original position] = unique index of its key
__global__ void k_inverse(const int* __restrict__ sidx, const int* __restrict__ uidx, int64_t n, int* __restrict__ inv) {
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q < n) inv[sidx[q]] = uidx[q] - 1;
}

void launch_shard_route(const SparseTable& req, int64_t n, int world, int64_t rps, int* off, int* counts, int slot, int* send_ids,
                        int* inv, cudaStream_t st) {
  PAMREC_PROF("shard_route", 3, st);
  k_owner_offsets<<<1, 96, 0, st>>>(req.ukeys, req.nuniq, world, rps, off, counts, slot);
  if (n == 0) return;
  const unsigned g = (unsigned)((n + 255) / 256);
  k_local_rows<<<g, 256, 0, st>>>(req.ukeys, req.nuniq, rps, send_ids);
  k_inverse<<<g, 256, 0, st>>>(req.sidx, req.uidx, n, inv);
}

// owner service: out[j, :] = shard[ids[j], :]   (one thread per 16-byte chunk)
template <int W>
__global__ void k_serve_rows(const float* __restrict__ shard, const int* __restrict__ ids, int64_t n, float* __restrict__ out) {
  constexpr int CH = W / 4;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * CH) return;
  const int64_t j = i / CH;
  const int c = (int)(i % CH);
  st4(out + j * W + 4 * c, __ldg(reinterpret_cast<const float4*>(shard + (int64_t)__ldg(ids + j) * W) + c));
}
void launch_serve_rows(const float* shard, const int* ids, int64_t n, int width, float* out, cudaStream_t st) {
  PAMREC_PROF("shard_serve_rows", 1, st);
  if (n == 0) return;
  const int64_t total = n * (width / 4);
  const unsigned g = (unsigned)((total + 255) / 256);
  if (width == 16) k_serve_rows<16><<<g, 256, 0, st>>>(shard, ids, n, out);
  else k_serve_rows<4><<<g, 256, 0, st>>>(shard, ids, n, out);
}

// number of listwise groups with a non-zero label sum (ApproxNDCG weight, pamrec.py:76) -> out[0] (double)
__global__ void k_count_valid_groups(const float* __restrict__ plays, int G, double* __restrict__ out) {
  __shared__ int sh;
  if (threadIdx.x == 0) sh = 0;
  __syncthreads();
  int c = 0;
  for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < G; g += gridDim.x * blockDim.x) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < PAMREC_GROUP; ++i) s += plays[g * PAMREC_GROUP + i];
    c += (s > 0.f) ? 1 : 0;
  }
  if (c) atomicAdd(&sh, c);
  __syncthreads();
  if (threadIdx.x == 0 && sh) atomicAdd(out, (double)sh);
}
void launch_count_valid_groups(const float* plays, int B, double* out, cudaStream_t st) {
  PAMREC_PROF("count_valid_groups", 1, st);
  const int G = B / PAMREC_GROUP;
  if (G == 0) return;
  k_count_valid_groups<<<(G + 255) / 256, 256, 0, st>>>(plays, G, out);
}

}  // namespace pamrec
