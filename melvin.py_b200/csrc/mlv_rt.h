// melvin-b200: thin runtime layer (allocation, copies, kernel launch).
// The CUDA build maps 1:1 onto the CUDA runtime.  The MLV_EMU build (tests/emu,
// development harness only, see mlv_common.cuh) runs CTAs as fibers on the host.
#pragma once

#include "mlv_common.cuh"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

namespace mlv {

void set_error(const char* fmt, ...);
extern unsigned long long g_launches;

#ifdef MLV_EMU
typedef void* stream_t;
void emu_launch(unsigned grid, unsigned block, size_t smem, const std::function<void()>& body);
inline int rt_malloc(void** p, size_t n) { *p = malloc(n ? n : 1); return *p ? 0 : MLV_ERR_NOMEM; }
inline void rt_free(void* p) { free(p); }
inline int rt_h2d(void* d, const void* h, size_t n, stream_t) { memcpy(d, h, n); return 0; }
inline size_t rt_max_smem() { return 232448; }
#define MLV_LAUNCH(kfn, grid, block, smem, stream, ...)                                   \
    do {                                                                                  \
        if ((size_t)(smem) > mlv::rt_max_smem()) {                                        \
            mlv::set_error("shared memory request %zu too large", (size_t)(smem));        \
            return MLV_ERR_UNSUPPORTED;                                                   \
        }                                                                                 \
        mlv::emu_launch((grid), (block), (smem), [=]() { kfn(__VA_ARGS__); });            \
        ++mlv::g_launches;                                                                \
    } while (0)
#else
typedef cudaStream_t stream_t;
inline int rt_check(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return 0;
    set_error("%s: %s", what, cudaGetErrorString(e));
    return MLV_ERR_CUDA;
}
inline int rt_malloc(void** p, size_t n) { return rt_check(cudaMalloc(p, n ? n : 1), "cudaMalloc"); }
inline void rt_free(void* p) { cudaFree(p); }
inline int rt_h2d(void* d, const void* h, size_t n, stream_t s) {
    // plan tables are uploaded once at context creation: synchronous copy
    (void)s;
    return rt_check(cudaMemcpy(d, h, n, cudaMemcpyHostToDevice), "cudaMemcpy");
}
inline size_t rt_max_smem() { return 232448; }   // 227 KB opt-in per CTA on sm_100
#define MLV_LAUNCH(kfn, grid, block, smem, stream, ...)                                   \
    do {                                                                                  \
        if ((size_t)(smem) > mlv::rt_max_smem()) {                                        \
            mlv::set_error("shared memory request %zu too large", (size_t)(smem));        \
            return MLV_ERR_UNSUPPORTED;                                                   \
        }                                                                                 \
        static size_t configured_ = 48 * 1024; /* per call site = per instantiation */   \
        if ((size_t)(smem) > configured_) {                                               \
            cudaError_t e_ = cudaFuncSetAttribute(                                        \
                kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem));           \
            if (e_ != cudaSuccess) return mlv::rt_check(e_, "cudaFuncSetAttribute");      \
            configured_ = (size_t)(smem);                                                 \
        }                                                                                 \
        kfn<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);                          \
        cudaError_t e2_ = cudaGetLastError();                                             \
        if (e2_ != cudaSuccess) return mlv::rt_check(e2_, #kfn);                          \
        ++mlv::g_launches;                                                                \
    } while (0)
#endif

}  // namespace mlv
