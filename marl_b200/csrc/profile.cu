// Built-in launch profiler: when enabled, every kernel launch of the library is bracketed by a
// pair of CUDA events recorded on the stream it is launched on; marl_profile_collect() reports
// count and total device time per kernel name.  Off by default (zero overhead: one branch).
#include <map>
#include <string>
#include <vector>
#include <cstdio>
#include <cstring>
#include "profile.h"
#include "../../include/marl_b200.h"

namespace marl {
namespace {
bool g_on = false;
struct Rec { const char* name; cudaEvent_t a, b; int m, n, k; };
int g_m = 0, g_n = 0, g_k = 0;
std::vector<Rec> g_recs;
std::vector<cudaEvent_t> g_pool;
cudaEvent_t get_event() {
    if (!g_pool.empty()) { cudaEvent_t e = g_pool.back(); g_pool.pop_back(); return e; }
    cudaEvent_t e; cudaEventCreate(&e); return e;
}
}  // namespace
bool prof_enabled() { return g_on; }
void prof_begin(const char* name, cudaStream_t st) {
    Rec r{name, get_event(), get_event(), g_m, g_n, g_k};
    g_m = g_n = g_k = 0;
    cudaEventRecord(r.a, st);
    g_recs.push_back(r);
}
void prof_note(int m, int n, int k) { g_m = m; g_n = n; g_k = k; }
void prof_end(cudaStream_t st) { cudaEventRecord(g_recs.back().b, st); }
}  // namespace marl

extern "C" int marl_profile_enable(int on) {
    marl::g_on = on != 0;
    return 0;
}

// Timeline of the launches recorded since the last collect: "kernel,start_us,end_us\n" relative to the first
// recorded launch (event timestamps of different streams share the device clock).  Does not clear the records.
extern "C" int marl_profile_timeline(char* buf, int buflen) {
    using namespace marl;
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) return (int)e;
    std::string out;
    char line[256];
    for (auto& r : g_recs) {
        float s = 0.f, t = 0.f;
        cudaEventElapsedTime(&s, g_recs.front().a, r.a);
        cudaEventElapsedTime(&t, g_recs.front().a, r.b);
        if (r.m) snprintf(line, sizeof line, "%s[%dx%dx%d],%.3f,%.3f\n", r.name, r.m, r.n, r.k, s * 1e3, t * 1e3);
        else snprintf(line, sizeof line, "%s,%.3f,%.3f\n", r.name, s * 1e3, t * 1e3);
        out += line;
    }
    if (buf && buflen > 0) {
        strncpy(buf, out.c_str(), buflen - 1);
        buf[buflen - 1] = 0;
    }
    return 0;
}

extern "C" int marl_profile_collect(char* buf, int buflen) {
    using namespace marl;
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) return (int)e;
    std::map<std::string, std::pair<int, double>> agg;
    for (auto& r : g_recs) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, r.a, r.b);
        auto& s = agg[r.name];
        s.first += 1; s.second += ms;
        g_pool.push_back(r.a); g_pool.push_back(r.b);
    }
    g_recs.clear();
    std::string out;
    char line[256];
    for (auto& kv : agg) {
        snprintf(line, sizeof line, "%s,%d,%.6f\n", kv.first.c_str(), kv.second.first, kv.second.second);
        out += line;
    }
    if (buf && buflen > 0) {
        strncpy(buf, out.c_str(), buflen - 1);
        buf[buflen - 1] = 0;
    }
    return 0;
}
