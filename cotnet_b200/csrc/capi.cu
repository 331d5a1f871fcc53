// libcotb200 -- library-level C ABI: version, error reporting, launch accounting.
#include "common.cuh"
#include <map>
#include <mutex>
#include <string>
#include <vector>

namespace cotb200 {

static thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}


// ---- per-launch event timing -----------------------------------------------------------------
static std::atomic<int> g_prof_on{0};
static std::mutex g_prof_mu;
struct ProfRec { const char* name; cudaEvent_t e0, e1; double bytes; };
static std::vector<ProfRec> g_prof_recs;

ProfScope::ProfScope(const char* n, cudaStream_t s, double b) : name(n), st(s), e0(nullptr), on(false), bytes(b) {
  if (!g_prof_on.load(std::memory_order_relaxed)) return;
  if (cudaEventCreate(&e0) != cudaSuccess) return;
  cudaEventRecord(e0, st);
  on = true;
}
ProfScope::~ProfScope() {
  if (!on) return;
  cudaEvent_t e1;
  if (cudaEventCreate(&e1) != cudaSuccess) { cudaEventDestroy(e0); return; }
  cudaEventRecord(e1, st);
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_prof_recs.push_back({name, e0, e1, bytes});
}

}  // namespace cotb200

extern "C" void cotb200_prof_enable(int on) {
  std::lock_guard<std::mutex> lk(cotb200::g_prof_mu);
  for (auto& r : cotb200::g_prof_recs) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
  cotb200::g_prof_recs.clear();
  cotb200::g_prof_on.store(on ? 1 : 0);
}

// Writes one line per kernel name: "<name> <launches> <total_ms> <total_algorithmic_bytes>\n".  Synchronises the recorded events.
extern "C" int cotb200_prof_report(char* buf, int len) {
  std::lock_guard<std::mutex> lk(cotb200::g_prof_mu);
  struct Agg { long long n = 0; double ms = 0, bytes = 0; };
  std::map<std::string, Agg> agg;
  for (auto& r : cotb200::g_prof_recs) {
    if (cudaEventSynchronize(r.e1) != cudaSuccess) continue;
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.e0, r.e1) != cudaSuccess) continue;
    auto& a = agg[r.name];
    a.n += 1; a.ms += ms; a.bytes += r.bytes;
  }
  std::string out;
  char line[256];
  for (auto& kv : agg) {
    snprintf(line, sizeof(line), "%s %lld %.6f %.0f\n", kv.first.c_str(), kv.second.n, kv.second.ms, kv.second.bytes);
    out += line;
  }
  if (buf && len > 0) { snprintf(buf, len, "%s", out.c_str()); }
  return (int)out.size();
}

extern "C" int cotb200_version(void) { return COTB200_VERSION; }
extern "C" const char* cotb200_last_error(void) { return cotb200::g_err; }
extern "C" long long cotb200_launch_count(void) { return cotb200::g_launches.load(); }
