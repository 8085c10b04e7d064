// NCCL through dlopen: see comm.h.
#include "comm.h"

#include <dlfcn.h>
#include <nccl.h>

#include <cstdio>
#include <mutex>

namespace pamrec {

namespace {
struct Api {
  void* dl = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*CommAbort)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
Api g_api;
std::mutex g_mu;

bool load_api(const char* path, std::string* err) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_api.dl) return true;
  void* dl = nullptr;
  if (path && *path) dl = dlopen(path, RTLD_NOW | RTLD_GLOBAL);
  if (!dl) dl = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!dl) {
    if (err) *err = std::string("cannot load libnccl.so.2: ") + dlerror();
    return false;
  }
  Api a;
  a.dl = dl;
#define PAMREC_SYM(field, name)                                            \
  a.field = reinterpret_cast<decltype(a.field)>(dlsym(dl, name));          \
  if (!a.field) { if (err) *err = std::string("libnccl lacks ") + name; return false; }
  PAMREC_SYM(GetUniqueId, "ncclGetUniqueId")
  PAMREC_SYM(CommInitRank, "ncclCommInitRank")
  PAMREC_SYM(CommDestroy, "ncclCommDestroy")
  PAMREC_SYM(CommAbort, "ncclCommAbort")
  PAMREC_SYM(AllReduce, "ncclAllReduce")
  PAMREC_SYM(Send, "ncclSend")
  PAMREC_SYM(Recv, "ncclRecv")
  PAMREC_SYM(GroupStart, "ncclGroupStart")
  PAMREC_SYM(GroupEnd, "ncclGroupEnd")
  PAMREC_SYM(GetErrorString, "ncclGetErrorString")
#undef PAMREC_SYM
  g_api = a;
  return true;
}

ncclDataType_t nccl_type(CommType t) { return t == COMM_F32 ? ncclFloat32 : (t == COMM_I32 ? ncclInt32 : ncclFloat64); }
size_t type_bytes(CommType t) { return t == COMM_F64 ? 8 : 4; }
}  // namespace

#define PAMREC_NCCL(call)                                                                     \
  do {                                                                                        \
    ncclResult_t r_ = (call);                                                                 \
    if (r_ != ncclSuccess) { err = std::string(#call ": ") + g_api.GetErrorString(r_); return -1; } \
  } while (0)

int Comm::unique_id(const char* path, char out[128], std::string* e) {
  if (!load_api(path, e)) return -1;
  ncclUniqueId id;
  static_assert(sizeof(id) == 128, "ncclUniqueId is 128 bytes");
  ncclResult_t r = g_api.GetUniqueId(&id);
  if (r != ncclSuccess) { if (e) *e = g_api.GetErrorString(r); return -1; }
  memcpy(out, &id, 128);
  return 0;
}

int Comm::init(const char* path, const char id_bytes[128], int w, int r) {
  world = w; rank = r;
  if (w <= 1) return 0;
  if (!load_api(path, &err)) return -1;
  ncclUniqueId id;
  memcpy(&id, id_bytes, 128);
  ncclComm_t c = nullptr;
  PAMREC_NCCL(g_api.CommInitRank(&c, w, id, r));
  comm = c;
  return 0;
}

void Comm::destroy() {
  if (comm && g_api.CommDestroy) g_api.CommDestroy((ncclComm_t)comm);
  comm = nullptr;
}

// ends the NCCL group opened by the caller; `first` is the first error seen inside it (ncclSuccess = none)
int Comm::close_group(int first, const char* where) {
  ncclResult_t end = g_api.GroupEnd();
  ncclResult_t r = first != ncclSuccess ? (ncclResult_t)first : end;
  if (r == ncclSuccess) return 0;
  err = std::string(where) + ": " + g_api.GetErrorString(r);
  abort();
  return -1;
}

// a fatal communication error: peers blocked in a collective fail instead of waiting for this rank forever
void Comm::abort() {
  if (comm && g_api.CommAbort) g_api.CommAbort((ncclComm_t)comm);
  comm = nullptr;
}

int Comm::group_start() {
  if (world <= 1) return 0;
  PAMREC_NCCL(g_api.GroupStart());
  return 0;
}
int Comm::group_end() {
  if (world <= 1) return 0;
  PAMREC_NCCL(g_api.GroupEnd());
  return 0;
}

int Comm::all_reduce(void* buf, int64_t count, CommType t, cudaStream_t st) {
  if (world <= 1 || count == 0) return 0;
  if (!comm) { err = "communicator not initialised (pamrec_comm_init)"; return -1; }
  PAMREC_NCCL(g_api.AllReduce(buf, buf, (size_t)count, nccl_type(t), ncclSum, (ncclComm_t)comm, st));
  return 0;
}

int Comm::all_to_all(const void* send, void* recv, int64_t count, CommType t, cudaStream_t st) {
  const size_t eb = type_bytes(t);
  if (world <= 1) {
    if (cudaMemcpyAsync(recv, send, (size_t)count * eb, cudaMemcpyDeviceToDevice, st) != cudaSuccess) { err = "memcpy"; return -1; }
    return 0;
  }
  if (!comm) { err = "communicator not initialised (pamrec_comm_init)"; return -1; }
  // An error inside an open group must still close it (and abort the communicator): a group left open queues every
  // later collective of this thread forever while the peers wait in theirs.
  PAMREC_NCCL(g_api.GroupStart());
  ncclResult_t first = ncclSuccess;
  for (int p = 0; p < world && first == ncclSuccess; ++p) {
    first = g_api.Send(static_cast<const char*>(send) + (size_t)p * count * eb, (size_t)count, nccl_type(t), p, (ncclComm_t)comm, st);
    if (first == ncclSuccess)
      first = g_api.Recv(static_cast<char*>(recv) + (size_t)p * count * eb, (size_t)count, nccl_type(t), p, (ncclComm_t)comm, st);
  }
  return close_group(first, "all_to_all");
}

int Comm::all_to_all_v(const void* send, const int64_t* soff, const int64_t* scnt, void* recv, const int64_t* roff,
                       const int64_t* rcnt, int width, CommType t, cudaStream_t st) {
  const size_t eb = type_bytes(t) * (size_t)width;
  if (world <= 1) {
    if (scnt[0] != rcnt[0]) { err = "all_to_all_v: self counts differ"; return -1; }
    if (scnt[0] > 0 &&
        cudaMemcpyAsync(static_cast<char*>(recv) + (size_t)roff[0] * eb, static_cast<const char*>(send) + (size_t)soff[0] * eb,
                        (size_t)scnt[0] * eb, cudaMemcpyDeviceToDevice, st) != cudaSuccess) { err = "memcpy"; return -1; }
    return 0;
  }
  if (!comm) { err = "communicator not initialised (pamrec_comm_init)"; return -1; }
  PAMREC_NCCL(g_api.GroupStart());
  ncclResult_t first = ncclSuccess;
  for (int p = 0; p < world && first == ncclSuccess; ++p) {
    if (scnt[p] > 0)
      first = g_api.Send(static_cast<const char*>(send) + (size_t)soff[p] * eb, (size_t)scnt[p] * width, nccl_type(t), p,
                         (ncclComm_t)comm, st);
    if (first == ncclSuccess && rcnt[p] > 0)
      first = g_api.Recv(static_cast<char*>(recv) + (size_t)roff[p] * eb, (size_t)rcnt[p] * width, nccl_type(t), p,
                         (ncclComm_t)comm, st);
  }
  return close_group(first, "all_to_all_v");
}

}  // namespace pamrec
