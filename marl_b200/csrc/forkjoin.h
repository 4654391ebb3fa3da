// Fork/join of independent launches onto side streams (capturable in CUDA graphs).
// The dense layers of one operator are small (one CTA per SM); running the independent ones side by
// side fills the machine without any change to the kernels.
#pragma once
#include <cuda_runtime.h>

namespace marl {

constexpr int kMaxSide = 4;

class ForkJoin {
public:
    // After fork(), work may be enqueued on lane(i) for i < n (lane(0) is the caller's stream).
    explicit ForkJoin(cudaStream_t main_stream, int n);
    cudaStream_t lane(int i) const { return i == 0 ? main_ : side_[i - 1]; }
    // Makes the caller's stream wait for every side lane.
    void join();

private:
    cudaStream_t main_;
    cudaStream_t side_[kMaxSide];
    int n_;
};

}  // namespace marl
