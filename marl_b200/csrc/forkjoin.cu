#include "forkjoin.h"
#include <cstdlib>

namespace marl {
namespace {
cudaStream_t g_side[kMaxSide];
cudaEvent_t g_fork_ev[8], g_join_ev[kMaxSide][8];
int g_ring = 0;
bool g_init = false;
bool g_enabled = true;

void init_once() {
    if (g_init) return;
    const char* e = getenv("MARL_B200_NO_FORK");
    g_enabled = !(e && e[0] == '1');
    for (int i = 0; i < kMaxSide; ++i) cudaStreamCreateWithFlags(&g_side[i], cudaStreamNonBlocking);
    for (int k = 0; k < 8; ++k) {
        cudaEventCreateWithFlags(&g_fork_ev[k], cudaEventDisableTiming);
        for (int i = 0; i < kMaxSide; ++i) cudaEventCreateWithFlags(&g_join_ev[i][k], cudaEventDisableTiming);
    }
    g_init = true;
}
}  // namespace

ForkJoin::ForkJoin(cudaStream_t main_stream, int n) : main_(main_stream), n_(n) {
    init_once();
    if (!g_enabled || n_ > kMaxSide + 1) n_ = 1;
    if (n_ < 1) n_ = 1;
    for (int i = 0; i < kMaxSide; ++i) side_[i] = (i < n_ - 1) ? g_side[i] : main_;
    if (n_ > 1) {
        g_ring = (g_ring + 1) & 7;
        cudaEventRecord(g_fork_ev[g_ring], main_);
        for (int i = 0; i < n_ - 1; ++i) cudaStreamWaitEvent(side_[i], g_fork_ev[g_ring], 0);
    }
}

void ForkJoin::join() {
    if (n_ <= 1) return;
    const int k = g_ring;
    for (int i = 0; i < n_ - 1; ++i) {
        cudaEventRecord(g_join_ev[i][k], side_[i]);
        cudaStreamWaitEvent(main_, g_join_ev[i][k], 0);
    }
}

}  // namespace marl
