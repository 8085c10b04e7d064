// Launch accounting and per-launcher CUDA-event timing.
#include "kernels.h"

namespace pamrec {

thread_local Prof* g_prof = nullptr;

int Prof::id_of(const char* name) {
  for (size_t i = 0; i < names.size(); ++i)
    if (names[i] == name) return (int)i;
  names.emplace_back(name);
  ms.push_back(0.0);
  cnt.push_back(0);
  return (int)names.size() - 1;
}
cudaEvent_t Prof::get_event() {
  if (!free_ev.empty()) { cudaEvent_t e = free_ev.back(); free_ev.pop_back(); return e; }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}
void Prof::resolve() {
  for (auto& pr : pending) {
    float t = 0.f;
    cudaEventSynchronize(pr.b);
    if (cudaEventElapsedTime(&t, pr.a, pr.b) == cudaSuccess) { ms[pr.id] += t; cnt[pr.id] += 1; }
    free_ev.push_back(pr.a);
    free_ev.push_back(pr.b);
  }
  pending.clear();
}
void Prof::reset() {
  resolve();
  for (auto& v : ms) v = 0.0;
  for (auto& c : cnt) c = 0;
}
Prof::~Prof() {
  for (auto& pr : pending) { cudaEventDestroy(pr.a); cudaEventDestroy(pr.b); }
  for (auto e : free_ev) cudaEventDestroy(e);
}

ProfScope::ProfScope(const char* name, int n_kernels, cudaStream_t s) : p(g_prof), st(s), b(nullptr), timed(false) {
  if (!p) return;
  p->launches += n_kernels;
  if (!p->on) return;
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing(s, &cs);
  if (cs != cudaStreamCaptureStatusNone) return;
  Prof::Pair pr;
  pr.a = p->get_event(); pr.b = p->get_event(); pr.id = p->id_of(name);
  cudaEventRecord(pr.a, s);
  b = pr.b;
  p->pending.push_back(pr);
  timed = true;
}
ProfScope::~ProfScope() {
  if (timed) cudaEventRecord(b, st);
}

}  // namespace pamrec
