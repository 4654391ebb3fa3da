// Optional per-kernel timing with CUDA events on the launching stream (bench.py's roofline leg).
#pragma once
#include <cuda_runtime.h>

namespace marl {
bool prof_enabled();
void prof_begin(const char* name, cudaStream_t st);
void prof_end(cudaStream_t st);
void prof_note(int m, int n, int k);   // problem size shown beside the next recorded launch in the timeline

struct ProfScope {
    cudaStream_t st; bool on;
    ProfScope(const char* name, cudaStream_t s) : st(s), on(prof_enabled()) { if (on) prof_begin(name, st); }
    ~ProfScope() { if (on) prof_end(st); }
};
}  // namespace marl
