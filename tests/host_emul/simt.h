// TEST INFRASTRUCTURE ONLY -- never linked into, loaded by or reachable from the package.
//
// A minimal SIMT-on-CPU shim: lets g++ compile the DEVICE code of fdtd_em_solver_b200/csrc/kernels.cuh unchanged
// (built with -DFDTD_HOST_EMUL, which only swaps the three PTX helpers ldvec / stvec / cp.async for plain C++) and run
// it one warp at a time, every lane a ucontext fiber, with the warp shuffles exchanged through a slot array.  The
// streaming kernels use no __syncthreads() and no cross-warp shared memory ("warps are independent"), so a warp at a
// time is exact.  What this pins without a GPU: the arithmetic and its order, the lane/halo/latch/ring bookkeeping,
// tile geometry and predicates -- bitwise against the oracle.  What it cannot see: memory-ordering around cp.async,
// register pressure, speed.
//
// Include BEFORE kernels.cuh.
#pragma once
#include <cuda_runtime.h>
#include <ucontext.h>

#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

namespace simt {

struct Idx { unsigned x = 0, y = 0, z = 0; };
struct Lane { Idx thread; ucontext_t ctx; std::vector<unsigned char> stack; bool done = false; };

inline Idx g_block_idx, g_block_dim, g_grid_dim;
inline Lane* g_cur = nullptr;                  // the lane that is running
inline ucontext_t g_sched;
inline uint64_t g_slot[2][32];                  // shuffle exchange, double-buffered by shuffle parity
inline unsigned g_round = 0;                    // shuffles completed by the running lane = index of its next one
inline unsigned g_lane_round[32];
inline std::function<void()> g_body;

inline void trampoline() { g_body(); g_cur->done = true; swapcontext(&g_cur->ctx, &g_sched); }

// Run `body` for the 32 lanes of one warp; threadIdx.x = warp * 32 + lane.
inline void run_warp(unsigned warp, const std::function<void()>& body) {
    static std::vector<Lane> lanes(32);
    g_body = body;
    for (unsigned l = 0; l < 32; ++l) {
        Lane& L = lanes[l];
        L.thread = Idx{warp * 32 + l, 0, 0};
        L.done = false;
        if (L.stack.empty()) L.stack.resize(256 * 1024);
        getcontext(&L.ctx);
        L.ctx.uc_stack.ss_sp = L.stack.data();
        L.ctx.uc_stack.ss_size = L.stack.size();
        L.ctx.uc_link = nullptr;
        makecontext(&L.ctx, trampoline, 0);
        g_lane_round[l] = 0;
    }
    for (;;) {                                   // rounds: every live lane runs up to its next shuffle
        bool any = false;
        for (unsigned l = 0; l < 32; ++l) {
            if (lanes[l].done) continue;
            any = true;
            g_cur = &lanes[l];
            swapcontext(&g_sched, &lanes[l].ctx);
        }
        if (!any) break;
    }
}

inline unsigned lane_id() { return g_cur->thread.x & 31; }

template <typename T>
inline T exchange(T v, int src_lane, bool valid) {           // all 32 lanes call this the same number of times
    static_assert(sizeof(T) <= 8, "shuffle of at most 64 bits");
    const unsigned me = lane_id();
    const unsigned r = g_lane_round[me]++;
    uint64_t raw = 0;
    std::memcpy(&raw, &v, sizeof(T));
    g_slot[r & 1][me] = raw;
    swapcontext(&g_cur->ctx, &g_sched);                        // everyone has written its slot when we come back
    T out = v;
    if (valid) std::memcpy(&out, &g_slot[r & 1][src_lane], sizeof(T));
    return out;
}

}  // namespace simt

// ---- the CUDA built-ins kernels.cuh uses --------------------------------------------------------------------------
#define threadIdx (simt::g_cur->thread)
#define blockIdx simt::g_block_idx
#define blockDim simt::g_block_dim
#define gridDim simt::g_grid_dim

template <typename T> inline T __shfl_sync(unsigned, T v, int src) { return simt::exchange(v, src & 31, true); }
template <typename T> inline T __shfl_down_sync(unsigned, T v, unsigned d) {
    const int s = (int)simt::lane_id() + (int)d;
    return simt::exchange(v, s, s < 32);
}
template <typename T> inline T __shfl_up_sync(unsigned, T v, unsigned d) {
    const int s = (int)simt::lane_id() - (int)d;
    return simt::exchange(v, s, s >= 0);
}
template <typename T> inline T __ldg(const T* p) { return *p; }
inline double __dadd_rn(double a, double b) { return a + b; }      // built with -ffp-contract=off: one rounding each
inline double __dsub_rn(double a, double b) { return a - b; }
inline double __dmul_rn(double a, double b) { return a * b; }
inline double __ddiv_rn(double a, double b) { return a / b; }
inline float __fadd_rn(float a, float b) { return a + b; }
inline float __fsub_rn(float a, float b) { return a - b; }
inline float __fmul_rn(float a, float b) { return a * b; }
using std::max;
using std::min;
#ifndef __grid_constant__
#define __grid_constant__
#endif
#ifndef __noinline__
#define __noinline__ __attribute__((noinline))
#endif
#ifndef __launch_bounds__
#define __launch_bounds__(...)
#endif
