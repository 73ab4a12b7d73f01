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

// development switches (kernel variants kept for A/B measurements): set = non-empty and not "0"
inline bool rt_env_flag(const char* name) {
    const char* v = getenv(name);
    return v && *v && !(v[0] == '0' && v[1] == 0);
}

#ifdef MLV_EMU
typedef void* stream_t;
void emu_launch(unsigned grid, unsigned block, size_t smem, const std::function<void()>& body);
inline int rt_malloc(void** p, size_t n) { *p = malloc(n ? n : 1); return *p ? 0 : MLV_ERR_NOMEM; }
inline void rt_free(void* p) { free(p); }
inline int rt_h2d(void* d, const void* h, size_t n, stream_t) { memcpy(d, h, n); return 0; }
inline size_t rt_max_smem() { return 232448; }
// tensor maps are emulated (mlv_common.cuh: EmuTmap): the TMA paths of the kernels run on the host too
inline bool rt_tma_enabled() {
    static int on = -1;
    if (on < 0) on = getenv("MLV_NO_TMA") ? 0 : 1;
    return on == 1;
}
inline bool rt_make_tmap(CUtensorMap* m, void* base, unsigned long long inner, unsigned long long rows,
                         unsigned long long row_bytes, unsigned box_inner, unsigned box_rows) {
    if (((uintptr_t)base & 15) || (row_bytes & 15) || box_inner > 256 || box_rows > 256) return false;
    EmuTmap* e = reinterpret_cast<EmuTmap*>(m);
    e->base = base; e->inner = inner; e->rows = rows; e->row_bytes = row_bytes;
    e->box_inner = box_inner; e->box_rows = box_rows;
    return true;
}
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

// TMA descriptor of a row-major float64 matrix (rows x inner doubles, row pitch in bytes) cut
// into boxes of box_rows x box_inner.  cuTensorMapEncodeTiled is a driver entry point: it is
// looked up at run time so that the library loads on machines without libcuda (build check).
inline bool rt_tma_enabled() {
    static int on = -1;
    if (on < 0) on = getenv("MLV_NO_TMA") ? 0 : 1;
    return on == 1;
}
inline bool rt_make_tmap(CUtensorMap* m, void* base, unsigned long long inner, unsigned long long rows,
                         unsigned long long row_bytes, unsigned box_inner, unsigned box_rows) {
    typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);
    static encode_fn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (encode_fn)p;
    }
    if (!fn || ((uintptr_t)base & 15) || (row_bytes & 15) || box_inner > 256 || box_rows > 256) return false;
    const cuuint64_t gdim[2] = {inner, rows};
    const cuuint64_t gstride[1] = {row_bytes};
    const cuuint32_t box[2] = {box_inner, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    return fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, base, gdim, gstride, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
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
