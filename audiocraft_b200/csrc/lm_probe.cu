// Measurement aids for the persistent decode step (no reference counterpart): what one grid-wide barrier costs on this
// GPU when every SM holds one co-resident CTA.  The number bounds how finely a decode step can be cut into dependent
// phases inside ONE kernel: the step kernel in lm_step.cu uses exactly this barrier.
#include "common.cuh"
#include "gridbar.cuh"

// variant 0: one arrival (red.release.gpu) + one polling thread (ld.acquire.gpu) per CTA on a single monotonic counter
// variant 1: the same, polling with ld.relaxed + a trailing fence.acquire (fewer L1 invalidations while spinning)
// variant 2: classic cooperative-groups style: __threadfence + atomicAdd (returning) + volatile spin
__global__ void __launch_bounds__(1024) acb_gridbar_probe_kernel(unsigned* ctr, int n, int variant, int work) {
    const unsigned G = gridDim.x;
    float acc = 0.f;
    for (int i = 0; i < n; ++i) {
        for (int w = 0; w < work; ++w) acc = fmaf(acc, 1.0001f, 0.5f);   // optional dependent work between barriers
        __syncthreads();
        if (threadIdx.x == 0) {
            const unsigned target = (unsigned)(i + 1) * G;
            if (variant == 0) {
                gridbar_arrive(ctr);
                gridbar_wait(ctr, target);
            } else if (variant == 1) {
                gridbar_arrive(ctr);
                gridbar_wait_relaxed(ctr, target);
            } else {
                __threadfence();
                atomicAdd(ctr, 1u);
                while (*((volatile unsigned*)ctr) < target) { }
                __threadfence();
            }
        }
        __syncthreads();
    }
    if (acc == 12345.678f) ctr[1] = 1;   // keep the dependent work alive
}

extern "C" int acb_debug_grid_barrier(int ctas, int threads, int n_barriers, int variant, int work, int reps, float* us_per_barrier) {
    ACB_REQUIRE(ctas >= 1 && threads >= 32 && threads <= 1024 && n_barriers >= 1 && reps >= 1 && us_per_barrier,
                "acb_debug_grid_barrier: bad argument");
    int dev = 0, sms = 0, per_sm = 0;
    ACB_CHECK_CUDA(cudaGetDevice(&dev));
    ACB_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    ACB_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, acb_gridbar_probe_kernel, threads, 0));
    ACB_REQUIRE(ctas <= sms * per_sm, "acb_debug_grid_barrier: %d CTAs cannot be co-resident (%d SMs x %d)", ctas, sms, per_sm);
    unsigned* ctr = nullptr;
    ACB_CHECK_CUDA(cudaMalloc(&ctr, 256));
    cudaStream_t s;
    ACB_CHECK_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaError_t e = cudaSuccess;
    float total_ms = 0.f;
    for (int r = 0; r < reps + 2 && e == cudaSuccess; ++r) {
        cudaMemsetAsync(ctr, 0, 256, s);
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(ctas); cfg.blockDim = dim3(threads); cfg.stream = s;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeCooperative;
        attr[0].val.cooperative = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        cudaEventRecord(e0, s);
        e = cudaLaunchKernelEx(&cfg, acb_gridbar_probe_kernel, ctr, n_barriers, variant, work);
        cudaEventRecord(e1, s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        float ms = 0.f;
        if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, e0, e1);
        if (r >= 2) total_ms += ms;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaStreamDestroy(s);
    cudaFree(ctr);
    if (e != cudaSuccess) { acb_set_error("acb_debug_grid_barrier: %s", cudaGetErrorString(e)); cudaGetLastError(); return ACB_ERR_CUDA; }
    *us_per_barrier = total_ms * 1e3f / (float)reps / (float)n_barriers;
    return ACB_OK;
}
