// Library identification / device probe.
#include "common.cuh"
#include "../../include/marl_b200.h"

extern "C" int marl_version(void) { return MARL_B200_VERSION; }

extern "C" int marl_device_info(int* sm_count, int* cc_major, int* cc_minor) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    cudaDeviceProp p;
    e = cudaGetDeviceProperties(&p, dev);
    if (e != cudaSuccess) return (int)e;
    if (sm_count) *sm_count = p.multiProcessorCount;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    return MARL_OK;
}

// FP32 FMA throughput probe: the denominator of the learner's compute roofline (bench.py).
__global__ void __launch_bounds__(256) fma_probe_kernel(float* out, int iters) {
    float a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = 1.0f + 1e-3f * (threadIdx.x + i);
    const float x = 1.0000001f, y = 1e-7f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = fmaf(a[i], x, y);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i];
    if (s == 12345.678f) out[0] = s;   // keep the loop alive
}

extern "C" int marl_fma_probe(float* scratch_device, int iters, int blocks, double* flops_out, void* stream) {
    if (!scratch_device || iters <= 0 || blocks <= 0) return MARL_EINVAL;
    fma_probe_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(scratch_device, iters);
    MARL_LAUNCH_CHECK();
    if (flops_out) *flops_out = 2.0 * 8.0 * (double)iters * 256.0 * (double)blocks;
    return MARL_OK;
}

// Holds the stream busy for ~`us` microseconds (one thread polling %globaltimer).  bench.py queues it in
// front of a profiled step so that every launch of the step is already enqueued when the GPU gets to it:
// the event pairs around the kernels then measure device time, not host launch gaps.
__global__ void spin_kernel(unsigned long long ns) {
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    do { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1)); } while (t1 - t0 < ns);
}

extern "C" int marl_spin_us(int us, void* stream) {
    if (us < 0 || us > 100000) return MARL_EINVAL;
    spin_kernel<<<1, 1, 0, (cudaStream_t)stream>>>((unsigned long long)us * 1000ull);
    MARL_LAUNCH_CHECK();
    return MARL_OK;
}

// 1 when [p, p + bytes) is page-locked host memory that kernels of the current device can read in place (cudaHostAlloc /
// cudaHostRegister; unified addressing makes the host pointer the device pointer), else 0.  The replay buffer uses it to ingest
// an episode straight from the caller's pinned arrays -- the float64 -> fp32 cast kernel pulls the bytes over PCIe itself --
// instead of packing them into a staging buffer first.
extern "C" int marl_host_registered(const void* p, size_t bytes) {
    if (!p) return 0;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return 0; }
    if (at.type != cudaMemoryTypeHost || at.devicePointer != p) return 0;
    if (bytes > 1) {
        cudaPointerAttributes last;
        if (cudaPointerGetAttributes(&last, static_cast<const char*>(p) + bytes - 1) != cudaSuccess) { cudaGetLastError(); return 0; }
        if (last.type != cudaMemoryTypeHost) return 0;
    }
    return 1;
}

// The same test for n ranges in one call (the 11 arrays of an episode): 1 when ALL of them qualify.
extern "C" int marl_host_registered_all(const void* const* p, const size_t* bytes, int n) {
    if (!p || !bytes || n <= 0) return 0;
    for (int i = 0; i < n; ++i)
        if (marl_host_registered(p[i], bytes[i]) != 1) return 0;
    return 1;
}
