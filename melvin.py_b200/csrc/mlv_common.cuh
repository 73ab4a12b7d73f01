// melvin-b200: common device helpers.
//
// The kernels in this directory are written for sm_100a (B200).  For
// development without a GPU the same sources can be compiled by a host C++
// compiler with -DMLV_EMU: every CUDA thread of a CTA becomes a fiber and
// __syncthreads() a yield to the fiber scheduler (tests/emu/).  MLV_EMU is a TEST HARNESS for
// index logic only -- it is never built into libmelvin_b200.so and there is no
// CPU fallback in the product.
#pragma once

#include <stdint.h>
#include <math.h>

#include "../../include/melvin_b200.h"   // MLV_OK / MLV_ERR_* codes

#ifdef MLV_EMU
// ------------------------------------------------------------------ emulation
#include <cstring>
#define __grid_constant__
struct alignas(64) CUtensorMap { unsigned long long opaque[16]; };
// emulated 2-D tensor map of doubles (filled by rt_make_tmap, mlv_rt.h): copies happen at once
struct EmuTmap { void* base; unsigned long long inner, rows, row_bytes; unsigned box_inner, box_rows; };
#include <functional>
#include <vector>
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __restrict__
#define MLV_UNROLL
struct double2 { double x, y; };
static inline double2 make_double2(double x, double y) { return double2{x, y}; }
struct emu_dim3 { unsigned x = 1, y = 1, z = 1; };
// CTAs run one at a time, their threads as fibers of one OS thread (tests/emu/emu_rt.cpp)
extern emu_dim3 threadIdx;
extern emu_dim3 blockIdx;
extern emu_dim3 blockDim;
extern emu_dim3 gridDim;
extern unsigned char* emu_smem_base;
void __syncthreads();
void emu_bar_sync(int id, int count);
void emu_bar_arrive(int id, int count);
void emu_yield();
#define MLV_SMEM_BASE() (emu_smem_base)
static inline double __ldg(const double* p) { return *p; }
static inline double2 __ldg(const double2* p) { return *p; }
#else
// ----------------------------------------------------------------------- CUDA
#include <cuda.h>            // CUtensorMap (types only; the driver entry point is looked up at run time)
#include <cuda_runtime.h>
#define MLV_UNROLL _Pragma("unroll")
extern __shared__ __align__(128) unsigned char mlv_dyn_smem[];
#define MLV_SMEM_BASE() (mlv_dyn_smem)
#endif

// compiler-only fence: keeps memory operations (and the registers they need) of one
// kernel phase from being hoisted into the previous one
#ifdef MLV_EMU
#define MLV_SCHED_FENCE() do {} while (0)
#else
#define MLV_SCHED_FENCE() asm volatile("" ::: "memory")
#endif

#define MLV_DEV __device__ __forceinline__

// Returns x, but hides the value from common-subexpression elimination: index and
// address arithmetic derived from it is recomputed per kernel phase instead of being
// kept live (and spilled) across whole transforms.
__device__ __forceinline__ int opaque_int(int x) {
#ifndef MLV_EMU
    asm volatile("" : "+r"(x));
#endif
    return x;
}
#define MLV_HD __host__ __device__ __forceinline__

namespace mlv {

typedef double2 cplx;

MLV_HD cplx mk(double x, double y) { cplx r; r.x = x; r.y = y; return r; }
MLV_HD cplx cadd(cplx a, cplx b) { return mk(a.x + b.x, a.y + b.y); }
MLV_HD cplx csub(cplx a, cplx b) { return mk(a.x - b.x, a.y - b.y); }
MLV_HD cplx cneg(cplx a) { return mk(-a.x, -a.y); }
MLV_HD cplx cconj(cplx a) { return mk(a.x, -a.y); }
MLV_HD cplx cscale(cplx a, double s) { return mk(a.x * s, a.y * s); }
MLV_HD cplx cmul(cplx a, cplx b) {
    return mk(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
// a * conj(b)
MLV_HD cplx cmulc(cplx a, cplx b) {
    return mk(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}
// multiply by +i
MLV_HD cplx cmuli(cplx a) { return mk(-a.y, a.x); }
// multiply by -i
MLV_HD cplx cmulni(cplx a) { return mk(a.y, -a.x); }

// Named barriers (PTX bar.sync / bar.arrive): sub-CTA synchronisation and
// producer/consumer hand-off between the two transform lines of a ping-pong CTA.
MLV_DEV void bar_sync(int id, int count) {
#ifdef MLV_EMU
    emu_bar_sync(id, count);
#else
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
#endif
}
MLV_DEV void bar_arrive(int id, int count) {
#ifdef MLV_EMU
    emu_bar_arrive(id, count);
#else
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
#endif
}

// L2 prefetch hints (no registers, no shared memory): the kernels keep only two
// transform lines per SM in flight, so loads that would otherwise expose a DRAM round
// trip are announced one phase ahead and later hit the L2.
MLV_DEV void l2_prefetch_bulk(const void* p, unsigned bytes) {   // 16-byte aligned, multiple of 16
#ifndef MLV_EMU
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
#else
    (void)p; (void)bytes;
#endif
}
MLV_DEV void l2_prefetch_line(const void* p) {
#ifndef MLV_EMU
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#else
    (void)p;
#endif
}

// log2 of a power of two
MLV_DEV int log2_pow2(int x) {
#ifdef MLV_EMU
    return __builtin_ctz((unsigned)x);
#else
    return __ffs(x) - 1;
#endif
}

// Predicated 16-byte global load: returns 0 where `pred` is false, without a branch.  A load
// inside a divergent `if` cannot be hoisted by the compiler, so a run of conditional loads
// becomes a chain of dependent memory round trips; predicated loads all issue back to back.
MLV_DEV cplx ldg_pred(const cplx* p, bool pred) {
#ifdef MLV_EMU
    return pred ? *p : mk(0.0, 0.0);
#else
    double x = 0.0, y = 0.0;
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %3, 0;\n\t@q ld.global.v2.f64 {%0, %1}, [%2];\n\t}"
                 : "+d"(x), "+d"(y) : "l"(p), "r"((int)pred));
    return mk(x, y);
#endif
}

// ---- TMA tensor stores (shared -> global, asynchronous, off the LSU pipe).  A column tile
// of an x pass is C*16 bytes wide and thousands of rows tall: as ordinary stores every warp
// instruction touches 16-32 different 128-byte lines; as one tensor store per 256 rows the
// copy engine walks the rows while the SM starts on the next transform.
MLV_DEV void tma_fence_smem() {        // generic-proxy smem writes -> visible to the async proxy
#ifndef MLV_EMU
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#endif
}
MLV_DEV void tma_store_2d(const CUtensorMap* map, const void* smem, int c0, int c1) {
#ifndef MLV_EMU
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"((unsigned long long)map), "r"(s), "r"(c0), "r"(c1) : "memory");
#else
    const EmuTmap* m = reinterpret_cast<const EmuTmap*>(map);     // elements outside the tensor are clipped
    const size_t w = (unsigned long long)c0 >= m->inner ? 0 :
                     ((unsigned long long)c0 + m->box_inner <= m->inner ? m->box_inner : (size_t)(m->inner - c0));
    for (unsigned r = 0; r < m->box_rows; ++r)
        if ((unsigned long long)c1 + r < m->rows)
            memcpy((char*)m->base + ((unsigned long long)c1 + r) * m->row_bytes + (size_t)c0 * 8,
                   (const char*)smem + (size_t)r * m->box_inner * 8, w * 8);
#endif
}
MLV_DEV void tma_commit() {
#ifndef MLV_EMU
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
#endif
}
MLV_DEV void tma_wait_read() {         // the shared-memory source of every committed store is free again
#ifndef MLV_EMU
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
#endif
}

// ---- asynchronous global -> shared copies completing on an mbarrier (TMA loads): the copy
// engine fetches a whole tile while the threads do something else; consumers wait on the
// barrier's phase parity.
MLV_DEV void mbar_init(unsigned long long* bar, int count) {
#ifndef MLV_EMU
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#else
    (void)count;
    *bar = 0;                  // emulation: low 32 bits = bytes still expected, bit 32 = phase parity
#endif
}
MLV_DEV void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
#ifndef MLV_EMU
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(bytes) : "memory");
#else
    *bar += bytes;
#endif
}
MLV_DEV void mbar_wait(unsigned long long* bar, unsigned parity) {
#ifndef MLV_EMU
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "MLV_WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, 0x989680;\n\t"
        "@p bra MLV_DONE_%=;\n\t"
        "bra MLV_WAIT_%=;\n\t"
        "MLV_DONE_%=:\n\t}" ::"r"(a), "r"(parity) : "memory");
#else
    while (((*(volatile unsigned long long*)bar >> 32) & 1ull) == parity) emu_yield();
#endif
}
#ifdef MLV_EMU
// emulated copies are complete on return: take their bytes off the barrier, flip the phase at 0
inline void emu_mbar_complete(unsigned long long* bar, unsigned bytes) {
    *bar -= bytes;
    if ((*bar & 0xffffffffull) == 0) *bar ^= (1ull << 32);
}
#endif
// contiguous global -> shared (multiple of 16 bytes, 16-byte aligned on both sides)
MLV_DEV void bulk_load(void* smem, const void* gmem, unsigned bytes, unsigned long long* bar) {
#ifndef MLV_EMU
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(s), "l"(gmem), "r"(bytes), "r"(b) : "memory");
#else
    memcpy(smem, gmem, bytes);
    emu_mbar_complete(bar, bytes);
#endif
}
// one box of a 2-D tensor map, global -> shared
MLV_DEV void tma_load_2d(void* smem, const CUtensorMap* map, int c0, int c1, unsigned long long* bar) {
#ifndef MLV_EMU
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(s), "l"((unsigned long long)map), "r"(b), "r"(c0), "r"(c1) : "memory");
#else
    const EmuTmap* m = reinterpret_cast<const EmuTmap*>(map);
    const size_t w = (unsigned long long)c0 >= m->inner ? 0 :
                     ((unsigned long long)c0 + m->box_inner <= m->inner ? m->box_inner : (size_t)(m->inner - c0));
    for (unsigned r = 0; r < m->box_rows; ++r) {
        char* d = (char*)smem + (size_t)r * m->box_inner * 8;
        memset(d, 0, (size_t)m->box_inner * 8);
        if ((unsigned long long)c1 + r < m->rows)
            memcpy(d, (const char*)m->base + ((unsigned long long)c1 + r) * m->row_bytes + (size_t)c0 * 8, w * 8);
    }
    emu_mbar_complete(bar, (unsigned)((size_t)m->box_inner * m->box_rows * 8));
#endif
}

// L2 prefetch of one box of a 2-D tensor map: one instruction for up to 256 rows of a column tile
MLV_DEV void tma_prefetch_2d(const CUtensorMap* map, int c0, int c1) {
#ifndef MLV_EMU
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];"
                 ::"l"((unsigned long long)map), "r"(c0), "r"(c1) : "memory");
#else
    (void)map; (void)c0; (void)c1;
#endif
}

// ---- cross-GPU ordering of producer and consumer kernels (peer-memory exchange): after a
// producer kernel a one-warp kernel bumps a 64-bit arrival counter in each consumer rank's
// memory (release at system scope); consumer CTAs spin on their own rank's counter (acquire at
// system scope) until the arrivals of all ranks are in.  Counters only grow; the host passes
// the cumulative count a launch has to wait for.
MLV_DEV void flag_signal(unsigned long long* counter) {
#ifndef MLV_EMU
    __threadfence_system();
    asm volatile("red.release.sys.global.add.u64 [%0], 1;" ::"l"(counter) : "memory");
#else
    (void)counter;
#endif
}
MLV_DEV void flag_wait(const unsigned long long* counter, unsigned long long value) {
#ifndef MLV_EMU
    unsigned long long cur;
    do {
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(cur) : "l"(counter) : "memory");
    } while (cur < value);
#else
    (void)counter; (void)value;
#endif
}

// Reciprocal to ~1 ulp without the branchy IEEE division sequence: 20-bit hardware
// seed + two Newton steps (operands here are O(1)..O(1e9), never denormal or zero).
MLV_DEV double fast_rcp(double x) {
#if defined(MLV_EMU)
    return 1.0 / x;
#else
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    return r;
#endif
}

}  // namespace mlv
