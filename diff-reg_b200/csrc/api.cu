// Library-wide C-ABI plumbing: version, thread-local error string, launch counter.
#include <stdarg.h>

#include <mutex>
#include <vector>

#include "common.cuh"

namespace drg {
static thread_local char g_err[512] = "";
std::atomic<unsigned long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace drg

namespace drg {
std::atomic<int> g_prof_enabled{0};
namespace {
struct ProfEntry {
  int slot;
  cudaEvent_t e0, e1;
};
std::mutex g_prof_mu;
std::vector<ProfEntry> g_prof_entries;
std::vector<cudaEvent_t> g_prof_pool;
cudaEvent_t prof_get_event() {
  if (!g_prof_pool.empty()) {
    cudaEvent_t e = g_prof_pool.back();
    g_prof_pool.pop_back();
    return e;
  }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}
}  // namespace
void prof_begin(int slot, cudaStream_t st) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  ProfEntry en{slot, prof_get_event(), prof_get_event()};
  cudaEventRecord(en.e0, st);
  g_prof_entries.push_back(en);
}
void prof_end(int slot, cudaStream_t st) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (size_t k = g_prof_entries.size(); k-- > 0;) {
    if (g_prof_entries[k].slot == slot) {
      cudaEventRecord(g_prof_entries[k].e1, st);
      return;
    }
  }
}
}  // namespace drg

namespace drg {
long long* g_tuning_stamps = nullptr;
}
/* Tuning hook (tools/pose_timeline.py, tools/skh_timeline.py): a caller-owned device buffer of >= 1024 int64 in which CTA 0 of
 * the persistent Sinkhorn (clock64) and the pose kernel (globaltimer ns, slots 800+) leave stamps of their phases; NULL
 * switches the stamps off.  The library itself never allocates. */
extern "C" int drg_tuning_set_stamp_buffer(long long* device_buffer) {
  drg::g_tuning_stamps = device_buffer;
  return DRG_OK;
}
extern "C" int drg_version(void) { return 200; }
extern "C" void drg_profile_enable(int on) { drg::g_prof_enabled.store(on ? 1 : 0); }
extern "C" void drg_profile_reset(void) {
  std::lock_guard<std::mutex> lk(drg::g_prof_mu);
  for (auto& en : drg::g_prof_entries) {
    drg::g_prof_pool.push_back(en.e0);
    drg::g_prof_pool.push_back(en.e1);
  }
  drg::g_prof_entries.clear();
}
extern "C" int drg_profile_read(int slot, double* total_ms, long long* count) {
  std::lock_guard<std::mutex> lk(drg::g_prof_mu);
  double tot = 0.0;
  long long n = 0;
  for (auto& en : drg::g_prof_entries) {
    if (en.slot != slot) continue;
    if (cudaEventSynchronize(en.e1) != cudaSuccess) return DRG_ERR_CUDA;
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, en.e0, en.e1) != cudaSuccess) return DRG_ERR_CUDA;
    tot += ms;
    ++n;
  }
  if (total_ms) *total_ms = tot;
  if (count) *count = n;
  return DRG_OK;
}
extern "C" int drg_profile_slots(void) { return drg::PROF_NSLOTS; }
extern "C" size_t drg_sizeof_sinkhorn_args(void) { return sizeof(drg_sinkhorn_args); }
extern "C" size_t drg_sizeof_procrustes_args(void) { return sizeof(drg_procrustes_args); }
extern "C" const char* drg_last_error(void) { return drg::g_err; }
extern "C" unsigned long long drg_launch_count(void) { return drg::g_launches.load(); }
