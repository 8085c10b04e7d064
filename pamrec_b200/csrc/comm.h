// NCCL plumbing of the data-parallel step.  libnccl.so.2 is resolved at run time (dlopen) so that the
// library still loads, and every host-only entry point still works, on a machine without NCCL or a GPU.
// With world == 1 every collective degenerates to a local copy / no-op, which lets a single GPU run the
// row-sharded code path end to end.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

namespace pamrec {

enum CommType { COMM_F32 = 0, COMM_I32 = 1, COMM_F64 = 2 };

struct Comm {
  int world = 1, rank = 0;
  void* comm = nullptr;          // ncclComm_t
  std::string err;

  static int unique_id(const char* path, char out[128], std::string* err);
  int init(const char* path, const char id[128], int world, int rank);
  void destroy();
  ~Comm() { destroy(); }

  // all collectives enqueue on `st`; 0 = ok
  int group_start();
  int group_end();
  int close_group(int first_error, const char* where);   // GroupEnd even after an error inside the group, then abort
  void abort();
  int all_reduce(void* buf, int64_t count, CommType t, cudaStream_t st);
  // fixed-size all-to-all: `count` elements to / from every rank
  int all_to_all(const void* send, void* recv, int64_t count, CommType t, cudaStream_t st);
  // variable all-to-all: element offsets / counts per peer (host arrays of length world), `width` elements per unit
  int all_to_all_v(const void* send, const int64_t* soff, const int64_t* scnt, void* recv, const int64_t* roff,
                   const int64_t* rcnt, int width, CommType t, cudaStream_t st);
};

}  // namespace pamrec
