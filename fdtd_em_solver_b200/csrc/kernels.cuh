// Device code of libfdtd2d.so: the fused 2D Yee leapfrog step for sm_100a.
//
// Three kernels compute the SAME function (calls of the reference's Solver.update,
// /root/reference/solver.py:169-257), from generation A of the fields into generation B (ping-pong):
//
//   step_exact    one thread per cell; every output value is the closed form of the reference's statement
//                 sequence, written as the functions of struct Exact.  Slow (it re-reads its
//                 stencil from L1/L2) but obviously right: the in-repo GPU reference, and the fallback for the few
//                 corner cells of step_stream.
//   step_stream   ONE step.  Each warp owns a (tile_rows x 32*V) tile and marches down its rows with ex[i,:] and
//                 hn[i-1,:] carried in registers, hn[i,j-1] by warp shuffle and a one-column halo recomputed by
//                 lane 0.  Each field/coefficient word is read once with 128/256-bit loads; H and E updates are
//                 fused.  Tiles that touch a wall, a source line, a reflector or a probe ("mixed" tiles) run the
//                 same loop with guarded loads and exact per-cell helpers on register operands -- walls, sources
//                 and reflectors are tile predicates of the one launch, not extra launches.
//   stepK_stream  K = 2..8 steps per pass over HBM (temporal blocking): the same march as a K-stage register
//                 pipeline with sacrificial halo lanes and a cp.async shared-memory prefetch ring; plain interior
//                 tiles only, the host advances the rest ("frame") with masked step_stream launches.
//
// Arithmetic: explicit round-to-nearest intrinsics, never contracted to FMA (numpy rounds after every ufunc);
// the association order is the reference's, so fp64 results are bit-identical to the reference.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <string.h>
#include <type_traits>

#define FDTD_MAX_SRC 16
#define FDTD_MAX_RECT 24
#define FDTD_MAX_PROBE 16
#ifndef FDTD_LAUNCH_BOUNDS
#define FDTD_LAUNCH_BOUNDS 256
#endif
#ifndef FDTD_CHUNK
#define FDTD_CHUNK 4
#endif
#ifndef FDTD_LEAN_UNROLL
#define FDTD_LEAN_UNROLL 2
#endif
#ifndef FDTD_LEAN_UNROLL_BIG             // row-loop unroll of the K >= 7 instantiations of the lean / skew pipelines (B200, K = 8,
#define FDTD_LEAN_UNROLL_BIG 4           // unroll 2 -> 4: general 684 -> 699, uniform mu 754 -> 769, skewed 773 -> 788 Gcell/s)
#endif
#ifndef FDTD_LEAN_MIN_BLOCKS_K6
#define FDTD_LEAN_MIN_BLOCKS_K6 3
#endif

namespace fdtd {

__device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }

template <typename T>
struct SrcStep {          // one ACTIVE pulse of this call, in list order
    int kind, row, col, pad;
    T mag;                // pulse.magnitude(step)
    T magz;               // pulse.magnitude(step) * sqrt(MU/EPSILON)
};

struct RectI { int r0, r1, c0, c1; };   // already clipped to the array it applies to

template <typename T>
struct StepParams {
    const T* h; const T* ex; const T* ey; const T* muc; const T* epc;   // generation A + coefficients
    T* hn; T* exn; T* eyn;                                              // generation B
    double* probe_out;        // this call's record: n_probe x (h,ex,ey)
    long long pitch;
    int NR, NC;               // global h shape
    int row_off;              // global row of local row 0
    int lrows;                // local rows allocated
    int row_lo, row_hi;       // rows of the (NR+1)-row index space this launch writes
    int tile_rows;
    int absorb_u, absorb_d, absorb_l, absorb_r;
    int n_src, n_rect, n_probe;
    T S, oms;                 // C*dt/ds and 1 - C*dt/ds
    T last_mag;               // magnitude of the last active pulse in list order
    SrcStep<T> src[FDTD_MAX_SRC];
    RectI rh[FDTD_MAX_RECT], rex[FDTD_MAX_RECT], rey[FDTD_MAX_RECT];
    int probe_i[FDTD_MAX_PROBE], probe_j[FDTD_MAX_PROBE];
    const unsigned char* tile_mask;   // optional: [row tile][warp tile] != 0 -> compute, else skip (temporal-blocking frame)
    const int2* tile_list;            // optional compact launch: blockIdx.x -> (block column, row tile)
    int mask_stride;
    // split storage (dtype FDTD_F32X, T = double): h/ex/ey/hn/exn/eyn point at the fp32 "hi" plane of their buffer, the
    // bf16 "lo" plane of the same buffer starts split_lo_off BYTES behind it; 0 = plain storage.  Last on purpose: the
    // offsets of every other field, and with them the code of the kernels that do not read it, stay what they were.
    long long split_lo_off;
    // uniform mu_coef (every stored cell of the plane holds this value -- mu_r = 1 everywhere, as in all of the reference's
    // scripts and the C5 recipe; detected on the device when the coefficients arrive): stepK_lean_umu reads it from here
    // instead of streaming the plane.  Behind split_lo_off for the same reason.
    T mu_uniform;
};

// ---- split field storage: value ~ hi + lo, hi = fp32(value), lo = bf16(fp32(value - hi)), both round-to-nearest-even.
// 32 bits of significand in 6 bytes.  Bit operations only for the bf16 half, so that the device, the CPU emulator and
// the numpy oracle (oracle/split_storage.py) produce the same bits.
#if defined(FDTD_HOST_EMUL)
#define FDTD_HD inline
#else
#define FDTD_HD __host__ __device__ __forceinline__
#endif
FDTD_HD unsigned short split_bf16_bits(float f) {
#if defined(__CUDA_ARCH__)
    // one F2FP instead of eight integer instructions; the same bits for every finite value and for +-inf (both are
    // round-to-nearest-even on the bit pattern), a NaN becomes the canonical one -- no stored field is ever NaN
    return __bfloat16_as_ushort(__float2bfloat16_rn(f));
#else
    unsigned int u;
    memcpy(&u, &f, 4);
    if ((u & 0x7f800000u) == 0x7f800000u) return (unsigned short)(u >> 16);       // inf / nan: truncate
    u += 0x7fffu + ((u >> 16) & 1u);
    return (unsigned short)(u >> 16);
#endif
}
FDTD_HD float split_bf16_value(unsigned short b) {
    const unsigned int u = (unsigned int)b << 16;
    float f;
    memcpy(&f, &u, 4);
    return f;
}
FDTD_HD void split_encode(double x, float& hi, unsigned short& lo) {
    hi = (float)x;
    lo = split_bf16_bits((float)(x - (double)hi));
}
FDTD_HD double split_decode(float hi, unsigned short lo) { return (double)hi + (double)split_bf16_value(lo); }
// (double)(float)x without a conversion instruction: x rounded (to nearest, ties to even) to 53 - shift significant bits
// (24 for shift = 29) on the exponent range of float32, by adding and subtracting M = 1.5 * 2^(e + shift), e = the exponent of x clamped to float32's
// smallest normal exponent -- the sum lies in M's binade, whose spacing 2^(e - 23) is float32's for x (2^-149 under the
// clamp, the subnormal spacing), M is an even multiple of that spacing, so the rounding of the sum IS the rounding of x;
// the sign is copied back so that a value that rounds to zero keeps its sign as the conversion does.  Equal to
// (double)(float)x for every |x| <= FLT_MAX (beyond it the conversion gives inf; no stored field is).  Why: on sm_100 the
// conversions with a 64-bit side run on the XU pipe at 16 lanes per clock and SM, a quarter of the FP64 pipe's DADD rate;
// ncu showed the first split K-step kernel 87 % XU-bound (profiles/r2j_ncu_full_stepK_lean_split_raw.csv).
FDTD_HD double split_round_sig(double x, unsigned int shift) {  // 29: to float32's 24 bits; 45: to bfloat16's 8 bits
    unsigned long long u;
    memcpy(&u, &x, 8);
    const unsigned int top = (unsigned int)(u >> 32);
    unsigned int e = top & 0x7ff00000u;
    e = e < (897u << 20) ? (897u << 20) : e;
    const unsigned long long mb = (unsigned long long)(e + (shift << 20) + 0x00080000u) << 32;
    double M;
    memcpy(&M, &mb, 8);
#if defined(__CUDA_ARCH__)
    const double r = __dsub_rn(__dadd_rn(x, M), M);
#else
    volatile double t = x + M;                                   // one rounding each, whatever the host compiler's flags
    const double r = t - M;
#endif
    unsigned long long v;
    memcpy(&v, &r, 8);
    v = (v & 0x7fffffffffffffffull) | (u & 0x8000000000000000ull);
    double out;
    memcpy(&out, &v, 8);
    return out;
}
FDTD_HD double split_round_hi(double x) { return split_round_sig(x, 29u); }
// The same value with NO conversion instruction at all: the bf16 half is float32(x - hi) rounded once more to 8 bits, both
// as magic-number additions (bfloat16 has float32's exponent range, so the same clamp applies).  Four DADDs instead of two
// XU conversions: the K-step kernel gives some values of a stage to each pipe (FDTD_SPLIT_XU_PER_STAGE).
FDTD_HD double split_round_fp64(double x) {
    const double hi = split_round_sig(x, 29u);
    const double lo = split_round_sig(split_round_sig(x - hi, 29u), 45u);
    return hi + lo;
}
FDTD_HD double split_round(double x) {                         // what a value becomes on its way through storage
    const double hi = split_round_hi(x);                        // == (double)(float)x
    const float lo = split_bf16_value(split_bf16_bits((float)(x - hi)));
    return hi + (double)lo;
}

__device__ __forceinline__ bool inside(const RectI& r, int i, int j) {
    return i >= r.r0 && i < r.r1 && j >= r.c0 && j < r.c1;
}

// ---------------------------------------------------------------------------
// Closed-form semantics of one update() call.
// ---------------------------------------------------------------------------
template <typename T>
struct Exact {
    const StepParams<T>& P;
    __device__ explicit Exact(const StepParams<T>& p) : P(p) {}

    __device__ __forceinline__ T ld(const T* a, int i, int j) const {
        return __ldg(a + (long long)(i - P.row_off) * P.pitch + j);
    }

    // h after the hard point-source writes of solver.py:176-180
    __device__ __noinline__ T h_old(int i, int j) const {
        T v = ld(P.h, i, j);
        for (int s = 0; s < P.n_src; ++s)
            if (P.src[s].kind == 0 && P.src[s].row == i && P.src[s].col == j) v = P.src[s].mag;
        return v;
    }

    // new h before the wall overwrite: solver.py:182, :184-200, :203-204
    __device__ __noinline__ T hn_pre(int i, int j) const {
        const T ho = h_old(i, j);
        const T mu = ld(P.muc, i, j);
        const T dex = sub(ld(P.ex, i + 1, j), ld(P.ex, i, j));
        const T eyc = ld(P.ey, i, j), eyr = ld(P.ey, i, j + 1);
        T v = add(ho, mul(mu, sub(dex, sub(eyr, eyc))));
        for (int s = 0; s < P.n_src; ++s) {
            const SrcStep<T>& q = P.src[s];
            if (q.kind == 1) { if (j == q.col) v = add(ho, mul(mu, sub(dex, sub(eyr, add(eyc, q.magz))))); }
            else if (q.kind == 2) { if (j == q.col) v = add(v, mul(mu, P.last_mag)); }
            else if (q.kind == 3) { if (i == q.row + 1) v = sub(v, mul(mu, P.last_mag)); }
            else if (q.kind == 4) { if (i == q.row) v = sub(v, mul(mu, P.last_mag)); }
        }
        for (int r = 0; r < P.n_rect; ++r) if (inside(P.rh[r], i, j)) v = T(0);
        return v;
    }

    // ex after solver.py:207 (rows 0 and NR untouched)
    __device__ __noinline__ T ex_pre(int i, int j) const {
        const T e = ld(P.ex, i, j);
        if (i >= 1 && i <= P.NR - 1) return add(e, mul(ld(P.epc, i, j), sub(hn_pre(i, j), hn_pre(i - 1, j))));
        return e;
    }

    // ey after solver.py:208 and the TF/SF re-application of :211-217
    __device__ __noinline__ T ey_pre(int i, int j) const {
        T v = ld(P.ey, i, j);
        if (j >= 1 && j <= P.NC - 1) v = sub(v, mul(ld(P.epc, i, j), sub(hn_pre(i, j), hn_pre(i, j - 1))));
        for (int s = 0; s < P.n_src; ++s) {
            const SrcStep<T>& q = P.src[s];
            if (q.kind == 1 && j == q.col) {
                const int jm = (j == 0) ? P.NC - 1 : j - 1;     // python's h[:, col-1] wraps at col 0
                v = sub(v, mul(ld(P.epc, i, j), sub(sub(hn_pre(i, j), q.mag), hn_pre(i, jm))));
            }
        }
        return v;
    }

    // ex / ey after the up and down walls (solver.py:220-227)
    __device__ __noinline__ T ex_ud(int i, int j) const {
        if (P.absorb_u && i == 0) return add(mul(ld(P.ex, 0, j), P.oms), mul(ex_pre(1, j), P.S));
        if (P.absorb_d && i == P.NR) return add(mul(ld(P.ex, P.NR, j), P.oms), mul(ex_pre(P.NR - 1, j), P.S));
        return ex_pre(i, j);
    }
    __device__ __noinline__ T ey_ud(int i, int j) const {
        if (P.absorb_u && i == 0) return add(mul(ey_pre(0, j), P.oms), mul(ey_pre(1, j), P.S));
        if (P.absorb_d && i == P.NR - 1) return add(mul(ey_pre(P.NR - 1, j), P.oms), mul(ey_pre(P.NR - 2, j), P.S));
        return ey_pre(i, j);
    }

    // final values: left/right walls (solver.py:228-235), reflector zeroing (:238-242)
    __device__ __noinline__ T h_final(int i, int j) const {
        T v = hn_pre(i, j);
        if (P.absorb_u && i == 0) v = add(mul(h_old(0, j), P.oms), mul(h_old(1, j), P.S));
        if (P.absorb_d && i == P.NR - 1) v = add(mul(h_old(P.NR - 1, j), P.oms), mul(h_old(P.NR - 2, j), P.S));
        if (P.absorb_l && j == 0) v = add(mul(h_old(i, 0), P.oms), mul(h_old(i, 1), P.S));
        if (P.absorb_r && j == P.NC - 1) v = add(mul(h_old(i, P.NC - 1), P.oms), mul(h_old(i, P.NC - 2), P.S));
        return v;
    }
    __device__ __noinline__ T ex_final(int i, int j) const {
        T v;
        if (P.absorb_l && j == 0) v = add(mul(ex_ud(i, 0), P.oms), mul(ex_ud(i, 1), P.S));
        else if (P.absorb_r && j == P.NC - 1) v = add(mul(ex_ud(i, P.NC - 1), P.oms), mul(ex_ud(i, P.NC - 2), P.S));
        else v = ex_ud(i, j);
        for (int r = 0; r < P.n_rect; ++r) if (inside(P.rex[r], i, j)) v = T(0);
        return v;
    }
    __device__ __noinline__ T ey_final(int i, int j) const {
        T v;
        if (P.absorb_l && j == 0) v = add(mul(ey_ud(i, 0), P.oms), mul(ey_ud(i, 1), P.S));
        else if (P.absorb_r && j == P.NC) v = add(mul(ey_ud(i, P.NC), P.oms), mul(ey_ud(i, P.NC - 1), P.S));
        else v = ey_ud(i, j);
        for (int r = 0; r < P.n_rect; ++r) if (inside(P.rey[r], i, j)) v = T(0);
        return v;
    }

    __device__ __noinline__ void probe(int i, int j, T hv, T exv, T eyv) const {
        for (int k = 0; k < P.n_probe; ++k)
            if (P.probe_i[k] == i && P.probe_j[k] == j) {
                P.probe_out[3 * k + 0] = (double)hv;
                P.probe_out[3 * k + 1] = (double)exv;
                P.probe_out[3 * k + 2] = (double)eyv;
            }
    }
};

template <typename T>
__global__ void __launch_bounds__(256) step_exact(const __grid_constant__ StepParams<T> P) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j > P.NC) return;
    const Exact<T> X(P);
    for (int i = P.row_lo + blockIdx.y; i < P.row_hi; i += gridDim.y) {      // grid.y is capped at 65535 rows
        const long long o = (long long)(i - P.row_off) * P.pitch + j;
        T hv = T(0), exv = T(0), eyv = T(0);
        if (i < P.NR && j < P.NC) { hv = X.h_final(i, j); P.hn[o] = hv; }
        if (j < P.NC) { exv = X.ex_final(i, j); P.exn[o] = exv; }
        if (i < P.NR) { eyv = X.ey_final(i, j); P.eyn[o] = eyv; }
        if (P.n_probe) X.probe(i, j, hv, exv, eyv);
    }
}

// ---------------------------------------------------------------------------
// Streaming kernel.
// ---------------------------------------------------------------------------
#ifndef FDTD_LD_MOD
#define FDTD_LD_MOD ".nc"            // inputs are never written by the running kernel (ping-pong)
#endif
#ifndef FDTD_ST_MOD
#define FDTD_ST_MOD ""
#endif

// V consecutive elements, 128-bit (V*sizeof(T)==16) or 256-bit (==32) per lane
// (FDTD_HOST_EMUL: tests/host_emul/simt.h runs this file's device code on the CPU, lane by lane, to pin kernel edits
// bitwise against the oracle without a GPU; the product build never defines it, so the PTX paths below are what ships)
template <typename T, int V>
__device__ __forceinline__ void ldvec(T (&r)[V], const T* p) {
#if defined(FDTD_HOST_EMUL)
    for (int v = 0; v < V; ++v) r[v] = p[v];
#else
    if constexpr (sizeof(T) == 8 && V == 2)
        asm("ld.global" FDTD_LD_MOD ".v2.f64 {%0,%1}, [%2];" : "=d"(r[0]), "=d"(r[1]) : "l"(p));
    else if constexpr (sizeof(T) == 8 && V == 4)
        asm("ld.global" FDTD_LD_MOD ".v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r[0]), "=d"(r[1]), "=d"(r[2]), "=d"(r[3]) : "l"(p));
    else if constexpr (sizeof(T) == 4 && V == 4)
        asm("ld.global" FDTD_LD_MOD ".v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]) : "l"(p));
    else
        asm("ld.global" FDTD_LD_MOD ".v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
            : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7]) : "l"(p));
#endif
}

template <typename T, int V>
__device__ __forceinline__ void stvec(T* p, const T (&r)[V]) {
#if defined(FDTD_HOST_EMUL)
    for (int v = 0; v < V; ++v) p[v] = r[v];
#else
    if constexpr (sizeof(T) == 8 && V == 2)
        asm volatile("st.global" FDTD_ST_MOD ".v2.f64 [%0], {%1,%2};" ::"l"(p), "d"(r[0]), "d"(r[1]) : "memory");
    else if constexpr (sizeof(T) == 8 && V == 4)
        asm volatile("st.global" FDTD_ST_MOD ".v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(r[0]), "d"(r[1]), "d"(r[2]), "d"(r[3]) : "memory");
    else if constexpr (sizeof(T) == 4 && V == 4)
        asm volatile("st.global" FDTD_ST_MOD ".v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(r[0]), "f"(r[1]), "f"(r[2]), "f"(r[3]) : "memory");
    else
        asm volatile("st.global" FDTD_ST_MOD ".v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(r[0]), "f"(r[1]),
                     "f"(r[2]), "f"(r[3]), "f"(r[4]), "f"(r[5]), "f"(r[6]), "f"(r[7]) : "memory");
#endif
}

template <typename T, int V>
struct Row {              // everything one lane loads for row i
    T h[V], mu[V], ep[V], exn[V], ey[V];
    T halo;               // lanes 0..4: h, mu, ex[i+1], ey at column jw-1; ey at column jw+W
};

template <typename T, int V, bool MIXED>
struct Tile {
    const StepParams<T>& P;
    int jw, lane, j;
    const T* halo_base; int halo_col; bool halo_on;

    __device__ Tile(const StepParams<T>& p, int jw_, int lane_) : P(p), jw(jw_), lane(lane_), j(jw_ + lane_ * V) {
        constexpr int W = 32 * V;
        halo_base = lane == 0 ? P.h : lane == 1 ? P.muc : lane == 2 ? P.ex : P.ey;
        halo_col = (lane == 4) ? jw + W : jw - 1;
        halo_on = lane < 5 && halo_col >= 0 && halo_col < P.pitch;
    }
    __device__ __forceinline__ bool row_ok(int i) const { return !MIXED || (i >= P.row_off && i < P.row_off + P.lrows); }
    __device__ __forceinline__ bool col_ok() const { return !MIXED || (j < P.pitch); }
    __device__ __forceinline__ long long off(int i) const { return (long long)(i - P.row_off) * P.pitch; }

    __device__ __forceinline__ void ld(T (&r)[V], const T* a, int i) const {
        if (row_ok(i) && col_ok()) ldvec<T, V>(r, a + off(i) + j);
        else {
#pragma unroll
            for (int v = 0; v < V; ++v) r[v] = T(0);
        }
    }
    // lanes 0,1,3,4 read row i; lane 2 reads ex row i+1
    __device__ __forceinline__ T ld_halo(int i) const {
        const int ii = i + (lane == 2 ? 1 : 0);
        if (halo_on && row_ok(ii)) return __ldg(halo_base + off(ii) + halo_col);
        return T(0);
    }
    __device__ __forceinline__ void load(Row<T, V>& r, int i) const {
        ld(r.h, P.h, i); ld(r.mu, P.muc, i); ld(r.ep, P.epc, i); ld(r.exn, P.ex, i + 1); ld(r.ey, P.ey, i);
        r.halo = ld_halo(i);
    }
};

// h after the point-source writes (solver.py:176-180), operand already in a register
template <typename T>
__device__ __forceinline__ T h_old_reg(const StepParams<T>& P, int i, int jj, T h) {
#pragma unroll 1
    for (int s = 0; s < P.n_src; ++s)
        if (P.src[s].kind == 0 && P.src[s].row == i && P.src[s].col == jj) h = P.src[s].mag;
    return h;
}

// hn_pre(i,jj) from register operands: plain curl, then the TF/SF fix-ups in list order, then reflector zeroing
template <typename T>
__device__ __forceinline__ T hn_exact_reg(const StepParams<T>& P, int i, int jj, T ho, T mu, T dex, T eyc, T eyr) {
    T v = add(ho, mul(mu, sub(dex, sub(eyr, eyc))));
#pragma unroll 1
    for (int s = 0; s < P.n_src; ++s) {
        const SrcStep<T>& q = P.src[s];
        if (q.kind == 1) { if (jj == q.col) v = add(ho, mul(mu, sub(dex, sub(eyr, add(eyc, q.magz))))); }
        else if (q.kind == 2) { if (jj == q.col) v = add(v, mul(mu, P.last_mag)); }
        else if (q.kind == 3) { if (i == q.row + 1) v = sub(v, mul(mu, P.last_mag)); }
        else if (q.kind == 4) { if (i == q.row) v = sub(v, mul(mu, P.last_mag)); }
    }
#pragma unroll 1
    for (int r = 0; r < P.n_rect; ++r) if (inside(P.rh[r], i, jj)) v = T(0);
    return v;
}

// ey after solver.py:208 and :211-217 for one cell, operands in registers (hl = hn_pre(i, jj-1))
template <typename T>
__device__ __forceinline__ T ey_exact_reg(const StepParams<T>& P, int jj, T ey, T ep, T hn, T hl) {
    T e = ey;
    if (jj >= 1 && jj <= P.NC - 1) e = sub(e, mul(ep, sub(hn, hl)));
#pragma unroll 1
    for (int s = 0; s < P.n_src; ++s)
        if (P.src[s].kind == 1 && P.src[s].col == jj && jj > 0) e = sub(e, mul(ep, sub(sub(hn, P.src[s].mag), hl)));
    return e;
}

// Rows 0, NR-1 and NR (up/down walls, solver.py:220-227, and the untouched ex edge rows) for an interior column
// 1 <= jj <= NC-2: the whole 2x2(+1) stencil is fetched in ONE batch of independent loads and evaluated in registers
// (the generic Exact recursion costs ~15 dependent memory round trips per cell).  Same operations, same order.
template <typename T>
__device__ __noinline__ void wall_row_cell(const StepParams<T>& P, int i, int jj, T& hw, T& exw, T& eyw) {
    const int r0 = (i == 0) ? 0 : P.NR - 2;
    const long long o0 = (long long)(r0 - P.row_off) * P.pitch + jj;
    T h[2][2], mu[2][2], ep[2], ey[2][3], ex[3][2];
#pragma unroll
    for (int a = 0; a < 2; ++a) {
        const long long o = o0 + a * P.pitch;
        h[a][0] = __ldg(P.h + o - 1); h[a][1] = __ldg(P.h + o);
        mu[a][0] = __ldg(P.muc + o - 1); mu[a][1] = __ldg(P.muc + o);
        ep[a] = __ldg(P.epc + o);
        ey[a][0] = __ldg(P.ey + o - 1); ey[a][1] = __ldg(P.ey + o); ey[a][2] = __ldg(P.ey + o + 1);
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) { ex[a][0] = __ldg(P.ex + o0 + a * P.pitch - 1); ex[a][1] = __ldg(P.ex + o0 + a * P.pitch); }
    T ho[2][2], hn[2][2], eyp[2];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            ho[a][b] = h_old_reg(P, r0 + a, jj - 1 + b, h[a][b]);
            hn[a][b] = hn_exact_reg(P, r0 + a, jj - 1 + b, ho[a][b], mu[a][b], sub(ex[a + 1][b], ex[a][b]), ey[a][b], ey[a][b + 1]);
        }
    const T exp1 = add(ex[1][1], mul(ep[1], sub(hn[1][1], hn[0][1])));          // ex_pre(r0+1, jj)
#pragma unroll
    for (int a = 0; a < 2; ++a) eyp[a] = ey_exact_reg(P, jj, ey[a][1], ep[a], hn[a][1], hn[a][0]);
    if (i == 0) {
        hw = P.absorb_u ? add(mul(ho[0][1], P.oms), mul(ho[1][1], P.S)) : hn[0][1];
        exw = P.absorb_u ? add(mul(ex[0][1], P.oms), mul(exp1, P.S)) : ex[0][1];
        eyw = P.absorb_u ? add(mul(eyp[0], P.oms), mul(eyp[1], P.S)) : eyp[0];
    } else if (i == P.NR - 1) {
        hw = P.absorb_d ? add(mul(ho[1][1], P.oms), mul(ho[0][1], P.S)) : hn[1][1];
        exw = exp1;
        eyw = P.absorb_d ? add(mul(eyp[1], P.oms), mul(eyp[0], P.S)) : eyp[1];
    } else {
        exw = P.absorb_d ? add(mul(ex[2][1], P.oms), mul(exp1, P.S)) : ex[2][1];
    }
#pragma unroll 1
    for (int r = 0; r < P.n_rect; ++r) {
        if (inside(P.rex[r], i, jj)) exw = T(0);
        if (inside(P.rey[r], i, jj)) eyw = T(0);
    }
}

// Walls (solver.py:220-235), reflector E zeroing (:238-242), closed-form fallbacks, store and probe for ONE cell
// of a mixed tile.  own = values at (i,jj); lft/rgt = values one column to the left/right (h_old, ex_pre, ey_pre).
template <typename T>
__device__ __noinline__ void finish_cell(const StepParams<T>& P, int i, int jj, bool no_left_neighbour, long long o,
                                         T hn, T ho, T exo, T eyo, T ho_l, T ex_l, T ey_l, T ho_r, T ex_r, T ey_r) {
    if (jj > P.NC) return;
    T hw = hn, exw = exo, eyw = eyo;
    bool closed = (i <= 0 || i >= P.NR - 1);                // rows 0, NR-1, NR: up/down walls and untouched edges
    if (jj == 0) {
#pragma unroll 1
        for (int s = 0; s < P.n_src; ++s) if (P.src[s].kind == 1 && P.src[s].col == 0) closed = true;   // wraps to NC-1
        if (P.absorb_l) {
            hw = add(mul(ho, P.oms), mul(ho_r, P.S));
            exw = add(mul(exo, P.oms), mul(ex_r, P.S));
            eyw = add(mul(eyo, P.oms), mul(ey_r, P.S));
        }
    }
    if (P.absorb_r && (jj == P.NC - 1 || jj == P.NC)) {
        if (no_left_neighbour) closed = true;               // the neighbour column belongs to another warp
        if (jj == P.NC - 1) { hw = add(mul(ho, P.oms), mul(ho_l, P.S)); exw = add(mul(exo, P.oms), mul(ex_l, P.S)); }
        else eyw = add(mul(eyo, P.oms), mul(ey_l, P.S));
    }
#pragma unroll 1
    for (int r = 0; r < P.n_rect; ++r) {
        if (inside(P.rex[r], i, jj)) exw = T(0);
        if (inside(P.rey[r], i, jj)) eyw = T(0);
    }
    const Exact<T> X(P);
    if (closed) {
        if ((i <= 0 || i >= P.NR - 1) && jj >= 1 && jj <= P.NC - 2) {
            wall_row_cell<T>(P, i, jj, hw, exw, eyw);
        } else {                                            // corner columns, TF/SF wrap at column 0, lone last column
            if (i < P.NR && jj < P.NC) hw = X.h_final(i, jj);
            if (jj < P.NC) exw = X.ex_final(i, jj);
            if (i < P.NR) eyw = X.ey_final(i, jj);
        }
    }
    if (i < P.NR && jj < P.NC) P.hn[o] = hw;
    if (jj < P.NC) P.exn[o] = exw;
    if (i < P.NR) P.eyn[o] = eyw;
    if (P.n_probe) X.probe(i, jj, hw, exw, eyw);
}

// New h of row i for the lane's V columns and for the halo column jw-1 (meaningful on lane 0).
// MIXED: exact hn_pre (sources, reflectors) and h_old; otherwise the plain interior curl.
template <typename T, int V, bool MIXED>
__device__ __forceinline__ void row_hn(const StepParams<T>& P, int i, int j, int jw, T (&hn)[V], T& hn_halo, T (&ho)[V],
                                       const T (&h)[V], const T (&mu)[V], const T (&exn)[V], const T (&exc)[V],
                                       const T (&ey)[V], T halo, T exc_halo, T& exn_halo) {
    const unsigned full = 0xffffffffu;
    T h_l = __shfl_sync(full, halo, 0);
    const T mu_l = __shfl_sync(full, halo, 1);
    exn_halo = __shfl_sync(full, halo, 2);
    const T ey_l = __shfl_sync(full, halo, 3), ey_far = __shfl_sync(full, halo, 4);
    T ey_next = __shfl_down_sync(full, ey[0], 1);
    if ((threadIdx.x & 31) == 31) ey_next = ey_far;
#pragma unroll
    for (int v = 0; v < V; ++v) {
        const T eyr = (v + 1 < V) ? ey[v + 1 < V ? v + 1 : 0] : ey_next;
        const T dex = sub(exn[v], exc[v]);
        if constexpr (MIXED) {
            ho[v] = h_old_reg(P, i, j + v, h[v]);
            hn[v] = hn_exact_reg(P, i, j + v, ho[v], mu[v], dex, ey[v], eyr);
        } else {
            ho[v] = h[v];
            hn[v] = add(h[v], mul(mu[v], sub(dex, sub(eyr, ey[v]))));
        }
    }
    const T dex_l = sub(exn_halo, exc_halo);
    if constexpr (MIXED) {
        h_l = h_old_reg(P, i, jw - 1, h_l);
        hn_halo = hn_exact_reg(P, i, jw - 1, h_l, mu_l, dex_l, ey_l, ey[0]);
    } else {
        hn_halo = add(h_l, mul(mu_l, sub(dex_l, sub(ey[0], ey_l))));
    }
}

template <typename T, int V, bool MIXED>
__device__ __forceinline__ void run_tile(const StepParams<T>& P, int ta, int tb, int jw, int lane) {
    const unsigned full = 0xffffffffu;
    const Tile<T, V, MIXED> t(P, jw, lane);
    const int j = t.j;

    // prologue: exc = ex[ta], hnp = hn_pre(ta-1)
    T exc[V], hnp[V], exc_halo;
    {
        T h[V], mu[V], exm[V], ey[V], ho[V], hh;
        t.ld(h, P.h, ta - 1); t.ld(mu, P.muc, ta - 1); t.ld(exm, P.ex, ta - 1); t.ld(exc, P.ex, ta); t.ld(ey, P.ey, ta - 1);
        const T halo = t.ld_halo(ta - 1);
        row_hn<T, V, MIXED>(P, ta - 1, j, jw, hnp, hh, ho, h, mu, exc, exm, ey, halo, T(0), exc_halo);   // lane 2 loaded ex[ta, jw-1]
    }

    Row<T, V> cur;
    t.load(cur, ta);
    for (int i = ta; i < tb; ++i) {
        Row<T, V> nxt;
        if (i + 1 < tb) t.load(nxt, i + 1);

        T hn[V], ho[V], hn_halo, exn_halo;
        row_hn<T, V, MIXED>(P, i, j, jw, hn, hn_halo, ho, cur.h, cur.mu, cur.exn, exc, cur.ey, cur.halo, exc_halo, exn_halo);
        T hn_left = __shfl_up_sync(full, hn[V - 1], 1);
        if (lane == 0) hn_left = hn_halo;

        T hout[V], exo[V], eyo[V];
#pragma unroll
        for (int v = 0; v < V; ++v) {
            const T hl = (v == 0) ? hn_left : hn[v > 0 ? v - 1 : 0];
            hout[v] = hn[v];
            exo[v] = add(exc[v], mul(cur.ep[v], sub(hn[v], hnp[v])));                     // solver.py:207
            if constexpr (!MIXED) eyo[v] = sub(cur.ey[v], mul(cur.ep[v], sub(hn[v], hl)));        // solver.py:208
            else eyo[v] = ey_exact_reg(P, j + v, cur.ey[v], cur.ep[v], hn[v], hl);
        }

        const long long o = t.off(i) + j;
        if constexpr (!MIXED) {
            stvec<T, V>(P.hn + o, hout); stvec<T, V>(P.exn + o, exo); stvec<T, V>(P.eyn + o, eyo);
        } else {
            // left neighbours of this lane's first column, for the right wall (solver.py:232-235)
            const T ho_lft = __shfl_up_sync(full, ho[V - 1], 1);
            const T ex_lft = __shfl_up_sync(full, exo[V - 1], 1);
            const T ey_lft = __shfl_up_sync(full, eyo[V - 1], 1);
#pragma unroll
            for (int v = 0; v < V; ++v) {
                const int l = v > 0 ? v - 1 : 0, r = v + 1 < V ? v + 1 : V - 1;
                finish_cell<T>(P, i, j + v, v == 0 && lane == 0, o + v, hn[v], ho[v], exo[v], eyo[v],
                               v == 0 ? ho_lft : ho[l], v == 0 ? ex_lft : exo[l], v == 0 ? ey_lft : eyo[l],
                               ho[r], exo[r], eyo[r]);
            }
        }
#pragma unroll
        for (int v = 0; v < V; ++v) { exc[v] = cur.exn[v]; hnp[v] = hn[v]; }
        exc_halo = exn_halo;
        cur = nxt;
    }
}

// Separate (non-inlined) bodies: each gets its own register allocation, so the slow path's calls into
// Exact cannot push spills into the streaming loop of the fast path.
template <typename T, int V>
__device__ __noinline__ void run_tile_mixed(const StepParams<T>& P, int ta, int tb, int jw, int lane) {
    run_tile<T, V, true>(P, ta, tb, jw, lane);
}
template <typename T, int V>
__device__ __noinline__ void run_tile_fast(const StepParams<T>& P, int ta, int tb, int jw, int lane) {
    run_tile<T, V, false>(P, ta, tb, jw, lane);
}

// Does anything non-plain touch index rows [ta,tb) x cols [ja,jb)?  (hn of one row up / one column left included)
template <typename T>
__device__ bool tile_is_mixed(const StepParams<T>& P, int ta, int tb, int ja, int jb) {
    if (ta <= 0 || tb > P.NR - 1 || ja <= 0 || jb > P.NC - 1) return true;
    const int ra = ta - 1, ca = ja - 1;     // hn region [ra,tb) x [ca,jb)
    for (int s = 0; s < P.n_src; ++s) {
        const SrcStep<T>& q = P.src[s];
        if (q.kind == 0) { if (q.row >= ra && q.row < tb && q.col >= ca && q.col < jb) return true; }
        else if (q.kind <= 2) { if (q.col >= ca && q.col < jb) return true; }
        else if (q.kind == 3) { if (q.row + 1 >= ra && q.row + 1 < tb) return true; }
        else { if (q.row >= ra && q.row < tb) return true; }
    }
    for (int r = 0; r < P.n_rect; ++r) {
        const RectI& a = P.rh[r];
        if (a.r0 < tb && a.r1 > ra && a.c0 < jb && a.c1 > ca) return true;
        const RectI& b = P.rex[r];
        if (b.r0 < tb && b.r1 > ta && b.c0 < jb && b.c1 > ja) return true;
        const RectI& c = P.rey[r];
        if (c.r0 < tb && c.r1 > ta && c.c0 < jb && c.c1 > ja) return true;
    }
    for (int k = 0; k < P.n_probe; ++k)
        if (P.probe_i[k] >= ta && P.probe_i[k] < tb && P.probe_j[k] >= ja && P.probe_j[k] < jb) return true;
    return false;
}

template <typename T, int V>
__global__ void __launch_bounds__(FDTD_LAUNCH_BOUNDS) step_stream(const __grid_constant__ StepParams<T> P) {
    constexpr int W = 32 * V;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int bx = blockIdx.x, by = blockIdx.y;
    if (P.tile_list) { const int2 t = P.tile_list[blockIdx.x]; bx = t.x; by = t.y; }
    const int jw = (bx * (blockDim.x >> 5) + warp) * W;
    if (jw > P.NC) return;                                   // index space has NC+1 columns
    const int ta = P.row_lo + by * P.tile_rows;
    const int tb = min(ta + P.tile_rows, P.row_hi);
    if (ta >= tb) return;
    if (P.tile_mask && !P.tile_mask[(long long)by * P.mask_stride + jw / W]) return;
    if (!tile_is_mixed(P, ta, tb, jw, jw + W)) {             // the common case, inlined: no call in its loop
        run_tile<T, V, false>(P, ta, tb, jw, lane);
        return;
    }
    // Classify in chunks of FDTD_CHUNK rows so that a wall row or a source row makes only its own chunk
    // "mixed"; consecutive plain chunks are merged into one streaming pass.
    int r = ta;
    while (r < tb) {
        int e = min(r + FDTD_CHUNK, tb);
        if (tile_is_mixed(P, r, e, jw, jw + W)) {
            while (e < tb) {                                 // merge consecutive mixed chunks too: one prologue per run
                const int e2 = min(e + FDTD_CHUNK, tb);
                if (!tile_is_mixed(P, e, e2, jw, jw + W)) break;
                e = e2;
            }
            run_tile_mixed<T, V>(P, r, e, jw, lane);
        } else {
            while (e < tb) {
                const int e2 = min(e + FDTD_CHUNK, tb);
                if (tile_is_mixed(P, e, e2, jw, jw + W)) break;
                e = e2;
            }
            run_tile_fast<T, V>(P, r, e, jw, lane);
        }
        r = e;
    }
}

// ---------------------------------------------------------------------------
// Temporal blocking: K leapfrog steps per HBM pass (generation A -> generation B = A + K steps), K = 2 .. 8.
//
// Only for tiles whose K-step dependency cone touches no wall, source, reflector or probe (the PLAIN tiles); the host
// marks all other tiles in `skip` and gives them to stepK_edge (edge.cuh; fdtd2d.cu: group_edge) -- or, in the fallback
// frame mode 0, advances them with K masked single-step launches through scratch generations (group_impl).
// Per warp: 32 lanes x V columns are LOADED; stage 1 (step n) is computed for
// the row just loaded, stage s (step n+s-1) trails s-1 rows behind, fed from register latches that hold the
// previous stage's last output row and from a register delay line of the coefficient rows; only stage K is stored.
// The first and last ceil(K/V) lanes are sacrificial halo lanes (their neighbour values are garbage, which moves
// one column per stage and so cannot reach the storing lanes), so all loads are aligned vectors and no scalar halo
// loads exist.  All stages run unconditionally: values computed from not-yet-valid latches in the first iterations
// are never consumed by a stored value (each stage's first needed row is fed by valid rows only).
// Same rounding order as the single-step kernel => bit-identical to K single steps.
// Algorithmic traffic: 8 words per cell per K steps.
// ---------------------------------------------------------------------------
// rows of the per-warp shared-memory prefetch ring of stepK_stream (0: one-row register prefetch; measured best for
// K = 2, while K >= 3 needs the deeper ring: K = 4 went from 243 to 315 Gcell/s, profiles/sweep_r1l_ring.jsonl)
template <int K> struct KRing { static constexpr int ROWS = (K <= 2) ? 0 : 4; };

#if defined(FDTD_HOST_EMUL)          // the copy lands at once: ordering bugs around wait_group are not modelled
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) { __builtin_memcpy(smem, gmem, 16); }
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) { __builtin_memcpy(smem, gmem, 8); }
__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) { __builtin_memcpy(smem, gmem, 4); }
__device__ __forceinline__ void cp_async_commit() {}
template <int N> __device__ __forceinline__ void cp_async_wait() {}
#else
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
#endif

template <int K, int V> struct KGeom {
    static constexpr int MARGIN = (K + V - 1) / V;           // sacrificial lanes on each side
    static constexpr int OUT = (32 - 2 * MARGIN) * V;        // columns stored per warp
};

template <typename T, int V, int K, int RING>
__device__ __forceinline__ void run_tileK(const StepParams<T>& P, int ta, int tb, int jw, int lane) {
    const unsigned full = 0xffffffffu;
    constexpr int MARGIN = KGeom<K, V>::MARGIN;
    const int j = jw + lane * V;
    const long long pitch = P.pitch;
    auto at = [&](const T* a, int i) { return a + (long long)(i - P.row_off) * pitch + j; };

    T ex0c[V];                       // ex of generation A at the stage-1 row
    T hnp[K][V];                     // hn_s of the previous row, per stage
    T Lh[K][V], Lex[K][V], Ley[K][V];// latch s (1..K-1): stage s outputs of the row it processed last iteration
    T Cmu[K][V], Cep[K][V];          // delay line: coefficients of row i-d, d = 1..K-1
#pragma unroll
    for (int s = 0; s < K; ++s)
#pragma unroll
        for (int v = 0; v < V; ++v) hnp[s][v] = Lh[s][v] = Lex[s][v] = Ley[s][v] = Cmu[s][v] = Cep[s][v] = T(0);
    {   // prologue: hn_1 of row ta-K and ex0 of row ta-K+1
        T h[V], mu[V], exm[V], ey[V];
        ldvec<T, V>(h, at(P.h, ta - K)); ldvec<T, V>(mu, at(P.muc, ta - K)); ldvec<T, V>(exm, at(P.ex, ta - K));
        ldvec<T, V>(ex0c, at(P.ex, ta - K + 1)); ldvec<T, V>(ey, at(P.ey, ta - K));
        const T nb = __shfl_down_sync(full, ey[0], 1);
#pragma unroll
        for (int v = 0; v < V; ++v) {
            const T eyr = (v + 1 < V) ? ey[v + 1 < V ? v + 1 : 0] : nb;
            hnp[0][v] = add(h[v], mul(mu[v], sub(sub(ex0c[v], exm[v]), sub(eyr, ey[v]))));
        }
    }
    Row<T, V> cur;
    const int i_first = ta - K + 1, i_last = tb + K - 2;
    // Rows are prefetched RING-1 deep into a per-warp shared-memory ring with cp.async (LDGSTS): memory-level
    // parallelism without register cost (with a one-row register prefetch the K=4 body was latency-bound: ncu
    // long_scoreboard 65 % of stalls, DRAM 60 %).  Every lane copies and later reads back only its own 16 bytes.
    static_assert(V * sizeof(T) == 16, "the ring moves one 16-byte vector per lane per array");
    constexpr int RR = RING > 0 ? RING : 1;
    extern __shared__ __align__(16) unsigned char ring_raw[];
    T* ring = reinterpret_cast<T*>(ring_raw) + (size_t)(threadIdx.x >> 5) * (RR * 5 * 32 * V) + lane * V;
    auto issue = [&](int i) {                                // row i -> slot (i - i_first) mod RING
        if (i <= i_last) {
            T* slot = ring + (size_t)((i - i_first) % RR) * (5 * 32 * V);
            cp_async16(slot + 0 * 32 * V, at(P.h, i)); cp_async16(slot + 1 * 32 * V, at(P.muc, i));
            cp_async16(slot + 2 * 32 * V, at(P.epc, i)); cp_async16(slot + 3 * 32 * V, at(P.ex, i + 1));
            cp_async16(slot + 4 * 32 * V, at(P.ey, i));
        }
        cp_async_commit();                                   // (an empty group past the end keeps the count uniform)
    };
    auto load = [&](Row<T, V>& r, int i) {
        ldvec<T, V>(r.h, at(P.h, i)); ldvec<T, V>(r.mu, at(P.muc, i)); ldvec<T, V>(r.ep, at(P.epc, i));
        ldvec<T, V>(r.exn, at(P.ex, i + 1)); ldvec<T, V>(r.ey, at(P.ey, i));
    };
    if constexpr (RING > 0) {
#pragma unroll
        for (int d = 0; d < RING; ++d) issue(i_first + d);
    } else {
        load(cur, i_first);
    }
#if defined(FDTD_K_UNROLL)           // experiment hook (tools/sass_loop_mix.py): unroll so that the latch / delay-line copies rename
    constexpr int ROW_UNROLL = FDTD_K_UNROLL;
#pragma unroll ROW_UNROLL
#endif
    for (int i = i_first; i <= i_last; ++i) {                // stage s works on row i-(s-1)
        Row<T, V> nxt;
        if constexpr (RING > 0) {
            cp_async_wait<(RING > 0 ? RING - 1 : 0)>();      // the oldest group (row i) has landed
            const T* slot = ring + (size_t)((i - i_first) % RR) * (5 * 32 * V);
#pragma unroll
            for (int v = 0; v < V; ++v) {
                cur.h[v] = slot[0 * 32 * V + v]; cur.mu[v] = slot[1 * 32 * V + v]; cur.ep[v] = slot[2 * 32 * V + v];
                cur.exn[v] = slot[3 * 32 * V + v]; cur.ey[v] = slot[4 * 32 * V + v];
            }
        } else {
            if (i < i_last) load(nxt, i + 1);
        }
        // one leapfrog step of one row: fields (h, exc, ey) of that row, ex of the row below (exn), coefficients,
        // hn of the row above (hp); outputs the new (h, ex, ey) of the row
        T oh[V], oex[V], oey[V];
        auto stage = [&](const T (&h)[V], const T (&exc)[V], const T (&exn)[V], const T (&ey)[V], const T (&mu)[V],
                         const T (&ep)[V], T (&hp)[V]) {
            const T nb = __shfl_down_sync(full, ey[0], 1);
#pragma unroll
            for (int v = 0; v < V; ++v) {
                const T eyr = (v + 1 < V) ? ey[v + 1 < V ? v + 1 : 0] : nb;
                oh[v] = add(h[v], mul(mu[v], sub(sub(exn[v], exc[v]), sub(eyr, ey[v]))));
            }
            const T lf = __shfl_up_sync(full, oh[V - 1], 1);
#pragma unroll
            for (int v = 0; v < V; ++v) {
                const T hl = (v == 0) ? lf : oh[v > 0 ? v - 1 : 0];
                oex[v] = add(exc[v], mul(ep[v], sub(oh[v], hp[v])));
                oey[v] = sub(ey[v], mul(ep[v], sub(oh[v], hl)));
                hp[v] = oh[v];
            }
        };
        stage(cur.h, ex0c, cur.exn, cur.ey, cur.mu, cur.ep, hnp[0]);                 // stage 1, row i
#pragma unroll
        for (int s = 1; s < K; ++s) {                        // stage s+1, row i-s, fed by latch s and by stage s's fresh ex
            T ph[V], pex[V], pey[V];
#pragma unroll
            for (int v = 0; v < V; ++v) { ph[v] = oh[v]; pex[v] = oex[v]; pey[v] = oey[v]; }   // stage s, row i-s+1
            stage(Lh[s], Lex[s], pex, Ley[s], Cmu[s], Cep[s], hnp[s]);
#pragma unroll
            for (int v = 0; v < V; ++v) { Lh[s][v] = ph[v]; Lex[s][v] = pex[v]; Ley[s][v] = pey[v]; }
        }
        const int r = i - (K - 1);                           // row the last stage just produced
        if (r >= ta && lane >= MARGIN && lane < 32 - MARGIN) {
            const long long o = (long long)(r - P.row_off) * pitch + j;
            stvec<T, V>(P.hn + o, oh); stvec<T, V>(P.exn + o, oex); stvec<T, V>(P.eyn + o, oey);
        }
#pragma unroll
        for (int s = K - 1; s >= 2; --s)
#pragma unroll
            for (int v = 0; v < V; ++v) { Cmu[s][v] = Cmu[s - 1][v]; Cep[s][v] = Cep[s - 1][v]; }
#pragma unroll
        for (int v = 0; v < V; ++v) { Cmu[1][v] = cur.mu[v]; Cep[1][v] = cur.ep[v]; ex0c[v] = cur.exn[v]; }
        if constexpr (RING > 0) issue(i + RING);             // refill the slot just consumed (its values are in registers)
        else cur = nxt;
    }
}

// K <= 6 fits 168 registers without spills => 12 warps per SM instead of 8 (K = 6: 395 -> 420 Gcell/s); K = 7, 8 would spill
template <int K> struct KBounds { static constexpr int MIN_BLOCKS = (K <= 6) ? 3 : 1; };

template <typename T, int V, int K>
__global__ void __launch_bounds__(128, KBounds<K>::MIN_BLOCKS) stepK_stream(const __grid_constant__ StepParams<T> P,
                                                       const unsigned char* __restrict__ skip, int skip_stride) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tx = blockIdx.x * (blockDim.x >> 5) + warp;
    if (tx >= skip_stride) return;
    if (skip[(long long)blockIdx.y * skip_stride + tx]) return;
    const int ta = P.row_lo + blockIdx.y * P.tile_rows;
    const int tb = min(ta + P.tile_rows, P.row_hi);
    if (ta >= tb) return;
    run_tileK<T, V, K, KRing<K>::ROWS>(P, ta, tb, tx * KGeom<K, V>::OUT - KGeom<K, V>::MARGIN * V, lane);
}

// ---------------------------------------------------------------------------
// stepK_lean: the same K-stage pipeline with fewer issue slots per row -- the default plain-tile kernel since its B200
// runs in round 2 (K = 8: 552 vs 468 Gcell/s for stepK_stream, profiles/r2_sweep_b_summary.txt; fdtd_set_kstep_variant(h, 0)
// selects stepK_stream).  The row loop of stepK_stream<double,2,6> issues 367 instructions per row, of which only 132 are
// the DADD/DMUL of the update: 125 are register moves (latch and delay-line copies) and ~70 are address arithmetic
// ((i - row_off) * pitch per array and row, ring slot modulo) -- profiles/r1_stepK_unroll_sass_mix.txt.  Here
//   * the row loop is unrolled (by two; by four for K >= 7), so the latch copies become register renaming (ptxas: 125 -> 29 moves per row),
//   * every global address is a running pointer advanced by one pitch per row, and the ring slot is a wrapping counter.
// Same operations in the same order on the same operands => bit-identical to stepK_stream.
// ---------------------------------------------------------------------------
template <typename T, int V, int K, int RING, int UMU = 0>
__device__ __forceinline__ void run_tileK_lean(const StepParams<T>& P, int ta, int tb, int jw, int lane) {
    const unsigned full = 0xffffffffu;
    constexpr int MARGIN = KGeom<K, V>::MARGIN;
    static_assert(RING > 0 && V * sizeof(T) == 16, "the lean variant always prefetches through the cp.async ring");
    const int j = jw + lane * V;
    const long long pitch = P.pitch;
    const int i_first = ta - K + 1, i_last = tb + K - 2;
    const long long o_first = (long long)(i_first - P.row_off) * pitch + j;      // element offset of row i_first

    T ex0c[V];
    T hnp[K][V];
    T Lh[K][V], Lex[K][V], Ley[K][V];
    T Cmu[K][V], Cep[K][V];
#pragma unroll
    for (int s = 0; s < K; ++s)
#pragma unroll
        for (int v = 0; v < V; ++v) hnp[s][v] = Lh[s][v] = Lex[s][v] = Ley[s][v] = Cmu[s][v] = Cep[s][v] = T(0);
    {   // prologue: hn_1 of row ta-K and ex0 of row ta-K+1 (same as run_tileK)
        T h[V], mu[V], exm[V], ey[V];
        ldvec<T, V>(h, P.h + o_first - pitch); ldvec<T, V>(exm, P.ex + o_first - pitch);
        if constexpr (UMU) {
#pragma unroll
            for (int v = 0; v < V; ++v) mu[v] = P.mu_uniform;
        } else {
            ldvec<T, V>(mu, P.muc + o_first - pitch);
        }
        ldvec<T, V>(ex0c, P.ex + o_first); ldvec<T, V>(ey, P.ey + o_first - pitch);
        const T nb = __shfl_down_sync(full, ey[0], 1);
#pragma unroll
        for (int v = 0; v < V; ++v) {
            const T eyr = (v + 1 < V) ? ey[v + 1 < V ? v + 1 : 0] : nb;
            hnp[0][v] = add(h[v], mul(mu[v], sub(sub(ex0c[v], exm[v]), sub(eyr, ey[v]))));
        }
    }
    extern __shared__ __align__(16) unsigned char ring_raw[];
    constexpr int SLOT = 5 * 32 * V;                             // elements per ring slot (5 arrays x one warp row)
    T* const ring = reinterpret_cast<T*>(ring_raw) + (size_t)(threadIdx.x >> 5) * (RING * SLOT) + lane * V;
    // running state of the prefetcher: next row to issue, its global offset, its ring slot
    int i_issue = i_first;
    long long o_issue = o_first;
    int s_issue = 0;
    auto issue = [&]() {
        if (i_issue <= i_last) {
            T* slot = ring + s_issue * SLOT;
            cp_async16(slot + 0 * 32 * V, P.h + o_issue);
            if constexpr (!UMU) cp_async16(slot + 1 * 32 * V, P.muc + o_issue);
            cp_async16(slot + 2 * 32 * V, P.epc + o_issue); cp_async16(slot + 3 * 32 * V, P.ex + o_issue + pitch);
            cp_async16(slot + 4 * 32 * V, P.ey + o_issue);
        }
        cp_async_commit();
        ++i_issue; o_issue += pitch;
        s_issue = (s_issue + 1 == RING) ? 0 : s_issue + 1;
    };
#pragma unroll
    for (int d = 0; d < RING; ++d) issue();
    int s_read = 0;                                              // ring slot of row i
    long long o_store = o_first - (long long)(K - 1) * pitch;    // offset of row i-(K-1), the row the last stage produces
    // 2: the latch copies rename (K = 6: 367 -> 263 instructions per row); 2*(K-1) would also rename the coefficient delay
    // line (-> ~242) at five times the code size -- a build knob for the GPU sweep: FDTD_NVCC_EXTRA=-DFDTD_LEAN_UNROLL=10
    constexpr int ROW_UNROLL = (K >= 7) ? FDTD_LEAN_UNROLL_BIG : FDTD_LEAN_UNROLL;
#pragma unroll ROW_UNROLL
    for (int i = i_first; i <= i_last; ++i) {
        Row<T, V> cur;
        cp_async_wait<RING - 1>();
        {
            const T* slot = ring + s_read * SLOT;
#pragma unroll
            for (int v = 0; v < V; ++v) {
                cur.h[v] = slot[0 * 32 * V + v]; cur.ep[v] = slot[2 * 32 * V + v];
                if constexpr (UMU) cur.mu[v] = P.mu_uniform; else cur.mu[v] = slot[1 * 32 * V + v];
                cur.exn[v] = slot[3 * 32 * V + v]; cur.ey[v] = slot[4 * 32 * V + v];
            }
        }
        T oh[V], oex[V], oey[V];
        auto stage = [&](const T (&h)[V], const T (&exc)[V], const T (&exn)[V], const T (&ey)[V], const T (&mu)[V],
                         const T (&ep)[V], T (&hp)[V]) {
            const T nb = __shfl_down_sync(full, ey[0], 1);
#pragma unroll
            for (int v = 0; v < V; ++v) {
                const T eyr = (v + 1 < V) ? ey[v + 1 < V ? v + 1 : 0] : nb;
                oh[v] = add(h[v], mul(mu[v], sub(sub(exn[v], exc[v]), sub(eyr, ey[v]))));
            }
            const T lf = __shfl_up_sync(full, oh[V - 1], 1);
#pragma unroll
            for (int v = 0; v < V; ++v) {
                const T hl = (v == 0) ? lf : oh[v > 0 ? v - 1 : 0];
                oex[v] = add(exc[v], mul(ep[v], sub(oh[v], hp[v])));
                oey[v] = sub(ey[v], mul(ep[v], sub(oh[v], hl)));
                hp[v] = oh[v];
            }
        };
        stage(cur.h, ex0c, cur.exn, cur.ey, cur.mu, cur.ep, hnp[0]);
#pragma unroll
        for (int s = 1; s < K; ++s) {
            T ph[V], pex[V], pey[V];
#pragma unroll
            for (int v = 0; v < V; ++v) { ph[v] = oh[v]; pex[v] = oex[v]; pey[v] = oey[v]; }
            if constexpr (UMU) stage(Lh[s], Lex[s], pex, Ley[s], cur.mu, Cep[s], hnp[s]);       // the plane is one value
            else stage(Lh[s], Lex[s], pex, Ley[s], Cmu[s], Cep[s], hnp[s]);
#pragma unroll
            for (int v = 0; v < V; ++v) { Lh[s][v] = ph[v]; Lex[s][v] = pex[v]; Ley[s][v] = pey[v]; }
        }
        if (i - (K - 1) >= ta && lane >= MARGIN && lane < 32 - MARGIN) {
            stvec<T, V>(P.hn + o_store, oh); stvec<T, V>(P.exn + o_store, oex); stvec<T, V>(P.eyn + o_store, oey);
        }
        o_store += pitch;
#pragma unroll
        for (int s = K - 1; s >= 2; --s)
#pragma unroll
            for (int v = 0; v < V; ++v) { if constexpr (!UMU) Cmu[s][v] = Cmu[s - 1][v]; Cep[s][v] = Cep[s - 1][v]; }
#pragma unroll
        for (int v = 0; v < V; ++v) { if constexpr (!UMU) Cmu[1][v] = cur.mu[v]; Cep[1][v] = cur.ep[v]; ex0c[v] = cur.exn[v]; }
        s_read = (s_read + 1 == RING) ? 0 : s_read + 1;
        issue();                                                 // refill the slot just consumed
    }
}

// K = 6 at three CTAs per SM (168 registers) spills a dozen values per two rows in the lean variant; two CTAs per SM
// (255 registers, 8 warps instead of 12) is the other candidate: -DFDTD_LEAN_MIN_BLOCKS_K6=2
template <int K> struct KBoundsLean { static constexpr int MIN_BLOCKS = (K == 6) ? FDTD_LEAN_MIN_BLOCKS_K6 : KBounds<K>::MIN_BLOCKS; };

template <typename T, int V, int K>
__global__ void __launch_bounds__(128, KBoundsLean<K>::MIN_BLOCKS) stepK_lean(const __grid_constant__ StepParams<T> P,
                                                     const unsigned char* __restrict__ skip, int skip_stride) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tx = blockIdx.x * (blockDim.x >> 5) + warp;
    if (tx >= skip_stride) return;
    if (skip[(long long)blockIdx.y * skip_stride + tx]) return;
    const int ta = P.row_lo + blockIdx.y * P.tile_rows;
    const int tb = min(ta + P.tile_rows, P.row_hi);
    if (ta >= tb) return;
    constexpr int RING = KRing<K>::ROWS > 0 ? KRing<K>::ROWS : 4;      // K = 2 prefetches through the ring here, too
    run_tileK_lean<T, V, K, RING>(P, ta, tb, tx * KGeom<K, V>::OUT - KGeom<K, V>::MARGIN * V, lane);
}

// stepK_lean with a UNIFORM mu_coef plane (P.mu_uniform): four streams per row instead of five -- 7 words per cell and pass
// instead of 8 -- and no coefficient delay line for mu (32 registers at K = 8).  The host launches it only after a device
// scan has found every stored cell of the plane equal (fdtd2d.cu: scan_mu_uniform); the value is the plane's, so the bits
// are those of stepK_lean.
#ifdef FDTD_UMU_MAXNREG                    // build knob for the occupancy sweep: cap the registers of the K >= 7 instantiations
#define FDTD_UMU_BOUNDS(K) __maxnreg__((K) >= 7 ? FDTD_UMU_MAXNREG : 168)
#else
#define FDTD_UMU_BOUNDS(K) __launch_bounds__(128, KBoundsLean<K>::MIN_BLOCKS)
#endif
template <typename T, int V, int K>
__global__ void FDTD_UMU_BOUNDS(K) stepK_lean_umu(const __grid_constant__ StepParams<T> P,
                                                         const unsigned char* __restrict__ skip, int skip_stride) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tx = blockIdx.x * (blockDim.x >> 5) + warp;
    if (tx >= skip_stride) return;
    if (skip[(long long)blockIdx.y * skip_stride + tx]) return;
    const int ta = P.row_lo + blockIdx.y * P.tile_rows;
    const int tb = min(ta + P.tile_rows, P.row_hi);
    if (ta >= tb) return;
    constexpr int RING = KRing<K>::ROWS > 0 ? KRing<K>::ROWS : 4;
    run_tileK_lean<T, V, K, RING, 1>(P, ta, tb, tx * KGeom<K, V>::OUT - KGeom<K, V>::MARGIN * V, lane);
}

// ---------------------------------------------------------------------------
// stepK_skew_umu: stepK_lean_umu with the stage chain CUT.  In the lean pipeline stage s+1 reads the ex row that stage s
// produces in the same iteration, so the K stages of a row form one dependency chain of 7 K DADD/DMUL; with two warps per
// scheduler that leaves about six independent operations in flight and the FP64 pipe at 65 % (ncu: top stall "wait",
// profiles/r2m_ncu_full_stepK_lean_umu_f64_k8_raw.csv).  At a SKEWED boundary (bit s of SKEW: between stages s and s+1)
// the consumer runs one more row behind: it takes (h, ex, ey) of its row from what stage s left TWO iterations ago and the
// ex of the row below from what it left in the previous one -- nothing from this iteration -- so the chain falls apart into
// independent pieces that the scheduler interleaves.  Price per skewed boundary: one more latched row (3 V values) and
// one more entry of the coefficient delay line; the registers come from the mu delay line this variant does not have.
// Same operations on the same operands in the same order per cell: the bits of stepK_lean.
// ---------------------------------------------------------------------------
template <int SKEW> __host__ __device__ constexpr int skew_count(int upto) {       // skewed boundaries among 1 .. upto
    int n = 0;
    for (int b = 1; b <= upto; ++b) n += (SKEW >> b) & 1;
    return n;
}

template <typename T, int V, int K, int RING, int SKEW>
__device__ __forceinline__ void run_tileK_skew_umu(const StepParams<T>& P, int ta, int tb, int jw, int lane) {
    const unsigned full = 0xffffffffu;
    constexpr int MARGIN = KGeom<K, V>::MARGIN;
    constexpr int NSK = skew_count<SKEW>(K - 1);                 // rows the last stage trails behind the lean pipeline's
    static_assert(RING > 0 && V * sizeof(T) == 16, "prefetches through the cp.async ring");
    const int j = jw + lane * V;
    const long long pitch = P.pitch;
    const int i_first = ta - K + 1, i_last = tb + K - 2;
    const long long o_first = (long long)(i_first - P.row_off) * pitch + j;      // element offset of row i_first

    T ex0c[V], muU[V];
    T hnp[K][V];
    T Lh[K][V], Lex[K][V], Ley[K][V];                            // what stage s left in the previous iteration
    T Mh[K][V], Mex[K][V], Mey[K][V];                            // ... and in the one before (skewed boundaries only)
    T Cep[K + NSK][V];                                           // Cep[k] = eps_coef of row i - k
#pragma unroll
    for (int v = 0; v < V; ++v) muU[v] = P.mu_uniform;
#pragma unroll
    for (int s = 0; s < K; ++s)
#pragma unroll
        for (int v = 0; v < V; ++v) hnp[s][v] = Lh[s][v] = Lex[s][v] = Ley[s][v] = Mh[s][v] = Mex[s][v] = Mey[s][v] = T(0);
#pragma unroll
    for (int s = 0; s < K + NSK; ++s)
#pragma unroll
        for (int v = 0; v < V; ++v) Cep[s][v] = T(0);
    {   // prologue: hn_1 of row ta-K and ex0 of row ta-K+1 (same as run_tileK)
        T h[V], exm[V], ey[V];
        ldvec<T, V>(h, P.h + o_first - pitch); ldvec<T, V>(exm, P.ex + o_first - pitch);
        ldvec<T, V>(ex0c, P.ex + o_first); ldvec<T, V>(ey, P.ey + o_first - pitch);
        const T nb = __shfl_down_sync(full, ey[0], 1);
#pragma unroll
        for (int v = 0; v < V; ++v) {
            const T eyr = (v + 1 < V) ? ey[v + 1 < V ? v + 1 : 0] : nb;
            hnp[0][v] = add(h[v], mul(muU[v], sub(sub(ex0c[v], exm[v]), sub(eyr, ey[v]))));
        }
    }
    extern __shared__ __align__(16) unsigned char ring_raw[];
    constexpr int SLOT = 5 * 32 * V;                             // the lean kernel's slot layout (the mu part stays unused)
    T* const ring = reinterpret_cast<T*>(ring_raw) + (size_t)(threadIdx.x >> 5) * (RING * SLOT) + lane * V;
    int i_issue = i_first;
    long long o_issue = o_first;
    int s_issue = 0;
    auto issue = [&]() {
        if (i_issue <= i_last) {
            T* slot = ring + s_issue * SLOT;
            cp_async16(slot + 0 * 32 * V, P.h + o_issue);
            cp_async16(slot + 2 * 32 * V, P.epc + o_issue); cp_async16(slot + 3 * 32 * V, P.ex + o_issue + pitch);
            cp_async16(slot + 4 * 32 * V, P.ey + o_issue);
        }
        cp_async_commit();
        ++i_issue; o_issue += pitch;
        s_issue = (s_issue + 1 == RING) ? 0 : s_issue + 1;
    };
#pragma unroll
    for (int d = 0; d < RING; ++d) issue();
    int s_read = 0;
    long long o_store = o_first - (long long)(K - 1 + NSK) * pitch;      // row the last stage produces in iteration i_first
    constexpr int ROW_UNROLL = (K >= 7) ? FDTD_LEAN_UNROLL_BIG : FDTD_LEAN_UNROLL;
    // NSK more iterations than the lean loop: the rows beyond i_last that stage 1 then "loads" are whatever the ring holds;
    // nothing that is stored depends on them
#pragma unroll ROW_UNROLL
    for (int i = i_first; i <= i_last + NSK; ++i) {
        Row<T, V> cur;
        cp_async_wait<RING - 1>();
        {
            const T* slot = ring + s_read * SLOT;
#pragma unroll
            for (int v = 0; v < V; ++v) {
                cur.h[v] = slot[0 * 32 * V + v]; cur.ep[v] = slot[2 * 32 * V + v];
                cur.exn[v] = slot[3 * 32 * V + v]; cur.ey[v] = slot[4 * 32 * V + v];
            }
        }
        T oh[V], oex[V], oey[V];
        auto stage = [&](const T (&h)[V], const T (&exc)[V], const T (&exn)[V], const T (&ey)[V], const T (&mu)[V],
                         const T (&ep)[V], T (&hp)[V]) {
            const T nb = __shfl_down_sync(full, ey[0], 1);
#pragma unroll
            for (int v = 0; v < V; ++v) {
                const T eyr = (v + 1 < V) ? ey[v + 1 < V ? v + 1 : 0] : nb;
                oh[v] = add(h[v], mul(mu[v], sub(sub(exn[v], exc[v]), sub(eyr, ey[v]))));
            }
            const T lf = __shfl_up_sync(full, oh[V - 1], 1);
#pragma unroll
            for (int v = 0; v < V; ++v) {
                const T hl = (v == 0) ? lf : oh[v > 0 ? v - 1 : 0];
                oex[v] = add(exc[v], mul(ep[v], sub(oh[v], hp[v])));
                oey[v] = sub(ey[v], mul(ep[v], sub(oh[v], hl)));
                hp[v] = oh[v];
            }
        };
        stage(cur.h, ex0c, cur.exn, cur.ey, muU, cur.ep, hnp[0]);
#pragma unroll
        for (int s = 1; s < K; ++s) {
            T ph[V], pex[V], pey[V];
#pragma unroll
            for (int v = 0; v < V; ++v) { ph[v] = oh[v]; pex[v] = oex[v]; pey[v] = oey[v]; }
            const int d = s + skew_count<SKEW>(s);               // stage s+1 works on row i - d (a constant after unrolling)
            if ((SKEW >> s) & 1) {
                stage(Mh[s], Mex[s], Lex[s], Mey[s], muU, Cep[d], hnp[s]);
#pragma unroll
                for (int v = 0; v < V; ++v) { Mh[s][v] = Lh[s][v]; Mex[s][v] = Lex[s][v]; Mey[s][v] = Ley[s][v]; }
            } else {
                stage(Lh[s], Lex[s], pex, Ley[s], muU, Cep[d], hnp[s]);
            }
#pragma unroll
            for (int v = 0; v < V; ++v) { Lh[s][v] = ph[v]; Lex[s][v] = pex[v]; Ley[s][v] = pey[v]; }
        }
        if (i - (K - 1 + NSK) >= ta && lane >= MARGIN && lane < 32 - MARGIN) {
            stvec<T, V>(P.hn + o_store, oh); stvec<T, V>(P.exn + o_store, oex); stvec<T, V>(P.eyn + o_store, oey);
        }
        o_store += pitch;
#pragma unroll
        for (int s = K - 1 + NSK; s >= 2; --s)
#pragma unroll
            for (int v = 0; v < V; ++v) Cep[s][v] = Cep[s - 1][v];
#pragma unroll
        for (int v = 0; v < V; ++v) { Cep[1][v] = cur.ep[v]; ex0c[v] = cur.exn[v]; }
        s_read = (s_read + 1 == RING) ? 0 : s_read + 1;
        issue();
    }
}

#ifndef FDTD_SKEW_MASK
#define FDTD_SKEW_MASK 0x48                  // boundaries 3|4 and 6|7: three chains of 3 + 3 + 2 stages at K = 8
#endif
template <int K> struct KSkew { static constexpr int MASK = FDTD_SKEW_MASK & ((1 << K) - 2); };

template <typename T, int V, int K>
__global__ void __launch_bounds__(128, 1) stepK_skew_umu(const __grid_constant__ StepParams<T> P,
                                                        const unsigned char* __restrict__ skip, int skip_stride) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tx = blockIdx.x * (blockDim.x >> 5) + warp;
    if (tx >= skip_stride) return;
    if (skip[(long long)blockIdx.y * skip_stride + tx]) return;
    const int ta = P.row_lo + blockIdx.y * P.tile_rows;
    const int tb = min(ta + P.tile_rows, P.row_hi);
    if (ta >= tb) return;
    run_tileK_skew_umu<T, V, K, 4, KSkew<K>::MASK>(P, ta, tb, tx * KGeom<K, V>::OUT - KGeom<K, V>::MARGIN * V, lane);
}

// ---------------------------------------------------------------------------
// stepK_lean_split: stepK_lean on SPLIT field storage (dtype FDTD_F32X: every field value = fp32 hi + bf16 lo, 6 bytes;
// float64 coefficient planes, float64 arithmetic in the reference's order).  The plain tiles of a K-step group in this
// mode; stepK_edge<.., SPLIT = 1> takes the others.  Same lane geometry, stage pipeline and ring as stepK_lean, with
//   * a ring slot of 2 x 16 B (mu, ep) + 3 x 8 B (hi pairs of h, ex, ey) + 3 x 4 B (lo pairs) per lane,
//   * loads decoded (hi + lo, exact in double), the last stage's outputs encoded on their way to generation B,
//   * every value one call hands to the next -- a stage's (h, ex, ey) outputs -- rounded through the storage format
//     (split_round), because a launch of K calls must leave the bits K stored calls leave (oracle/split_storage.py).
//     Inside a call nothing is rounded: ex / ey are computed from the unrounded new h, as the reference computes them.
// Algorithmic traffic: 3 x 6 B read + 3 x 6 B written + 2 x 8 B coefficients = 52 B per cell per K steps.
// ---------------------------------------------------------------------------
#ifndef FDTD_SPLIT_XU_PER_STAGE
#define FDTD_SPLIT_XU_PER_STAGE 5          // of the 6 values a stage hands on (B200 sweep 0..6: 191 204 211 217 224 211 Gcell/s, profiles/r2l_*)
#endif
struct SplitRing {                                               // byte offsets inside one ring slot of one warp
    static constexpr int MU = 0, EP = 512, HI = 1024, LO = 1024 + 3 * 256, SLOT = 1024 + 3 * 256 + 3 * 128;
};

template <int K, int RING>
__device__ __forceinline__ void run_tileK_lean_split(const StepParams<double>& P, int ta, int tb, int jw, int lane) {
    typedef double T;
    constexpr int V = 2;
    const unsigned full = 0xffffffffu;
    constexpr int MARGIN = KGeom<K, V>::MARGIN;
    const int j = jw + lane * V;
    const long long pitch = P.pitch;
    const int i_first = ta - K + 1, i_last = tb + K - 2;
    const long long o_first = (long long)(i_first - P.row_off) * pitch + j;      // element offset of row i_first
    const long long lo_off = P.split_lo_off;
    const float* const h_hi = reinterpret_cast<const float*>(P.h);
    const float* const ex_hi = reinterpret_cast<const float*>(P.ex);
    const float* const ey_hi = reinterpret_cast<const float*>(P.ey);
    const unsigned short* const h_lo = reinterpret_cast<const unsigned short*>(reinterpret_cast<const char*>(P.h) + lo_off);
    const unsigned short* const ex_lo = reinterpret_cast<const unsigned short*>(reinterpret_cast<const char*>(P.ex) + lo_off);
    const unsigned short* const ey_lo = reinterpret_cast<const unsigned short*>(reinterpret_cast<const char*>(P.ey) + lo_off);
    auto decode2 = [](T (&x)[V], const float* hi, const unsigned short* lo) {
        x[0] = split_decode(hi[0], lo[0]); x[1] = split_decode(hi[1], lo[1]);
    };

    T ex0c[V];
    T hnp[K][V];
    T Lh[K][V], Lex[K][V], Ley[K][V];
    T Cmu[K][V], Cep[K][V];
#pragma unroll
    for (int s = 0; s < K; ++s)
#pragma unroll
        for (int v = 0; v < V; ++v) hnp[s][v] = Lh[s][v] = Lex[s][v] = Ley[s][v] = Cmu[s][v] = Cep[s][v] = T(0);
    {   // prologue: hn_1 of row ta-K and ex0 of row ta-K+1 (as run_tileK_lean)
        T h[V], mu[V], exm[V], ey[V];
        decode2(h, h_hi + o_first - pitch, h_lo + o_first - pitch);
        ldvec<T, V>(mu, P.muc + o_first - pitch);
        decode2(exm, ex_hi + o_first - pitch, ex_lo + o_first - pitch);
        decode2(ex0c, ex_hi + o_first, ex_lo + o_first);
        decode2(ey, ey_hi + o_first - pitch, ey_lo + o_first - pitch);
        const T nb = __shfl_down_sync(full, ey[0], 1);
#pragma unroll
        for (int v = 0; v < V; ++v) {
            const T eyr = (v + 1 < V) ? ey[v + 1 < V ? v + 1 : 0] : nb;
            hnp[0][v] = add(h[v], mul(mu[v], sub(sub(ex0c[v], exm[v]), sub(eyr, ey[v]))));
        }
    }
    extern __shared__ __align__(16) unsigned char ring_raw[];
    unsigned char* const ring = ring_raw + (size_t)(threadIdx.x >> 5) * (RING * SplitRing::SLOT);
    int i_issue = i_first;
    long long o_issue = o_first;
    int s_issue = 0;
    auto issue = [&]() {
        if (i_issue <= i_last) {
            unsigned char* slot = ring + s_issue * SplitRing::SLOT;
            cp_async16(slot + SplitRing::MU + lane * 16, P.muc + o_issue);
            cp_async16(slot + SplitRing::EP + lane * 16, P.epc + o_issue);
            cp_async8(slot + SplitRing::HI + 0 * 256 + lane * 8, h_hi + o_issue);
            cp_async8(slot + SplitRing::HI + 1 * 256 + lane * 8, ex_hi + o_issue + pitch);
            cp_async8(slot + SplitRing::HI + 2 * 256 + lane * 8, ey_hi + o_issue);
            cp_async4(slot + SplitRing::LO + 0 * 128 + lane * 4, h_lo + o_issue);
            cp_async4(slot + SplitRing::LO + 1 * 128 + lane * 4, ex_lo + o_issue + pitch);
            cp_async4(slot + SplitRing::LO + 2 * 128 + lane * 4, ey_lo + o_issue);
        }
        cp_async_commit();
        ++i_issue; o_issue += pitch;
        s_issue = (s_issue + 1 == RING) ? 0 : s_issue + 1;
    };
#pragma unroll
    for (int d = 0; d < RING; ++d) issue();
    int s_read = 0;
    long long o_store = o_first - (long long)(K - 1) * pitch;    // offset of row i-(K-1), the row the last stage produces
    float* const hn_hi = reinterpret_cast<float*>(P.hn);
    float* const exn_hi = reinterpret_cast<float*>(P.exn);
    float* const eyn_hi = reinterpret_cast<float*>(P.eyn);
    unsigned short* const hn_lo = reinterpret_cast<unsigned short*>(reinterpret_cast<char*>(P.hn) + lo_off);
    unsigned short* const exn_lo = reinterpret_cast<unsigned short*>(reinterpret_cast<char*>(P.exn) + lo_off);
    unsigned short* const eyn_lo = reinterpret_cast<unsigned short*>(reinterpret_cast<char*>(P.eyn) + lo_off);
    auto put2 = [](float* hi, unsigned short* lo, const T (&x)[V]) {
        float2 a; ushort2 b;
        split_encode(x[0], a.x, b.x); split_encode(x[1], a.y, b.y);
        *reinterpret_cast<float2*>(hi) = a; *reinterpret_cast<ushort2*>(lo) = b;
    };
    constexpr int ROW_UNROLL = FDTD_LEAN_UNROLL;
#pragma unroll ROW_UNROLL
    for (int i = i_first; i <= i_last; ++i) {
        Row<T, V> cur;
        cp_async_wait<RING - 1>();
        {
            const unsigned char* slot = ring + s_read * SplitRing::SLOT;
            const double2 m2 = *reinterpret_cast<const double2*>(slot + SplitRing::MU + lane * 16);
            const double2 e2 = *reinterpret_cast<const double2*>(slot + SplitRing::EP + lane * 16);
            cur.mu[0] = m2.x; cur.mu[1] = m2.y; cur.ep[0] = e2.x; cur.ep[1] = e2.y;
            const float2 a = *reinterpret_cast<const float2*>(slot + SplitRing::HI + 0 * 256 + lane * 8);
            const float2 b = *reinterpret_cast<const float2*>(slot + SplitRing::HI + 1 * 256 + lane * 8);
            const float2 c = *reinterpret_cast<const float2*>(slot + SplitRing::HI + 2 * 256 + lane * 8);
            const ushort2 la = *reinterpret_cast<const ushort2*>(slot + SplitRing::LO + 0 * 128 + lane * 4);
            const ushort2 lb = *reinterpret_cast<const ushort2*>(slot + SplitRing::LO + 1 * 128 + lane * 4);
            const ushort2 lc = *reinterpret_cast<const ushort2*>(slot + SplitRing::LO + 2 * 128 + lane * 4);
            cur.h[0] = split_decode(a.x, la.x); cur.h[1] = split_decode(a.y, la.y);
            cur.exn[0] = split_decode(b.x, lb.x); cur.exn[1] = split_decode(b.y, lb.y);
            cur.ey[0] = split_decode(c.x, lc.x); cur.ey[1] = split_decode(c.y, lc.y);
        }
        T oh[V], oex[V], oey[V];
        auto stage = [&](const T (&h)[V], const T (&exc)[V], const T (&exn)[V], const T (&ey)[V], const T (&mu)[V],
                         const T (&ep)[V], T (&hp)[V]) {
            const T nb = __shfl_down_sync(full, ey[0], 1);
#pragma unroll
            for (int v = 0; v < V; ++v) {
                const T eyr = (v + 1 < V) ? ey[v + 1 < V ? v + 1 : 0] : nb;
                oh[v] = add(h[v], mul(mu[v], sub(sub(exn[v], exc[v]), sub(eyr, ey[v]))));
            }
            const T lf = __shfl_up_sync(full, oh[V - 1], 1);
#pragma unroll
            for (int v = 0; v < V; ++v) {
                const T hl = (v == 0) ? lf : oh[v > 0 ? v - 1 : 0];
                oex[v] = add(exc[v], mul(ep[v], sub(oh[v], hp[v])));
                oey[v] = sub(ey[v], mul(ep[v], sub(oh[v], hl)));
                hp[v] = oh[v];                                    // the same call's next row reads the UNROUNDED new h
            }
        };
        stage(cur.h, ex0c, cur.exn, cur.ey, cur.mu, cur.ep, hnp[0]);
#pragma unroll
        for (int s = 1; s < K; ++s) {
            T ph[V], pex[V], pey[V];                             // what call s left in the arrays: through the storage format
#pragma unroll
            for (int v = 0; v < V; ++v) {
                // six values per stage: the first FDTD_SPLIT_XU_PER_STAGE of (h0, ex0, ey0, h1, ex1, ey1) take the bf16 half
                // through two XU conversions, the others through four more DADDs -- same bits, two pipes kept busy
                ph[v] = (3 * v + 0 < FDTD_SPLIT_XU_PER_STAGE) ? split_round(oh[v]) : split_round_fp64(oh[v]);
                pex[v] = (3 * v + 1 < FDTD_SPLIT_XU_PER_STAGE) ? split_round(oex[v]) : split_round_fp64(oex[v]);
                pey[v] = (3 * v + 2 < FDTD_SPLIT_XU_PER_STAGE) ? split_round(oey[v]) : split_round_fp64(oey[v]);
            }
            stage(Lh[s], Lex[s], pex, Ley[s], Cmu[s], Cep[s], hnp[s]);
#pragma unroll
            for (int v = 0; v < V; ++v) { Lh[s][v] = ph[v]; Lex[s][v] = pex[v]; Ley[s][v] = pey[v]; }
        }
        if (i - (K - 1) >= ta && lane >= MARGIN && lane < 32 - MARGIN) {
            put2(hn_hi + o_store, hn_lo + o_store, oh); put2(exn_hi + o_store, exn_lo + o_store, oex);
            put2(eyn_hi + o_store, eyn_lo + o_store, oey);
        }
        o_store += pitch;
#pragma unroll
        for (int s = K - 1; s >= 2; --s)
#pragma unroll
            for (int v = 0; v < V; ++v) { Cmu[s][v] = Cmu[s - 1][v]; Cep[s][v] = Cep[s - 1][v]; }
#pragma unroll
        for (int v = 0; v < V; ++v) { Cmu[1][v] = cur.mu[v]; Cep[1][v] = cur.ep[v]; ex0c[v] = cur.exn[v]; }
        s_read = (s_read + 1 == RING) ? 0 : s_read + 1;
        issue();                                                 // refill the slot just consumed
    }
}

template <int K>
__global__ void __launch_bounds__(128, 2) stepK_lean_split(const __grid_constant__ StepParams<double> P,
                                                           const unsigned char* __restrict__ skip, int skip_stride) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tx = blockIdx.x * (blockDim.x >> 5) + warp;
    if (tx >= skip_stride) return;
    if (skip[(long long)blockIdx.y * skip_stride + tx]) return;
    const int ta = P.row_lo + blockIdx.y * P.tile_rows;
    const int tb = min(ta + P.tile_rows, P.row_hi);
    if (ta >= tb) return;
    run_tileK_lean_split<K, 4>(P, ta, tb, tx * KGeom<K, 2>::OUT - KGeom<K, 2>::MARGIN * 2, lane);
}

// ---------------------------------------------------------------------------
// Small service kernels.
// ---------------------------------------------------------------------------
// host double plane (dense, ld = cols) staged on the device -> pitched T plane
template <typename T>
__global__ void plane_from_f64(T* dst, long long pitch, const double* src, long long src_ld, int rows, int cols) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= cols) return;
    for (int i = blockIdx.y; i < rows; i += gridDim.y) dst[(long long)i * pitch + j] = (T)src[(long long)i * src_ld + j];
}

// pitched T plane -> dense double plane, with the snapshot aliasing patches applied
struct Patch { int row, col; double value; };
struct PatchList { int n; Patch p[FDTD_MAX_SRC]; };

template <typename T>
__global__ void plane_to_f64(double* dst, long long dst_ld, const T* src, long long pitch, int rows, int cols,
                             int row_global0, const __grid_constant__ PatchList patches) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= cols) return;
    for (int i = blockIdx.y * blockDim.y + threadIdx.y; i < rows; i += gridDim.y * blockDim.y) {
        double v = (double)src[(long long)i * pitch + j];
        for (int k = 0; k < patches.n; ++k)
            if (patches.p[k].row == row_global0 + i && patches.p[k].col == j) v = patches.p[k].value;
        dst[(long long)i * dst_ld + j] = v;
    }
}

// every row_stride-th row / col_stride-th column of a pitched T plane -> dense double plane (the picture of the realtime
// view; no aliasing patches: plot_and_show draws the field update() returned, solver.py:319-325)
template <typename T>
__global__ void plane_view(double* dst, long long dst_ld, const T* src, long long pitch, int rows_out, int cols_out,
                           int row_stride, int col_stride) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= cols_out) return;
    for (int i = blockIdx.y * blockDim.y + threadIdx.y; i < rows_out; i += gridDim.y * blockDim.y)
        dst[(long long)i * dst_ld + j] = (double)src[(long long)i * row_stride * pitch + (long long)j * col_stride];
}

// Is a plane one value?  out[0] receives src[0]; *differs is set (never cleared) when any of rows x cols differs from it.
// Bitwise comparison: +0 and -0 count as different, NaN equals the same NaN.
template <typename T>
__global__ void plane_uniform(const T* src, long long pitch, int rows, int cols, T* out, int* differs) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= cols) return;
    typedef typename std::conditional<sizeof(T) == 8, unsigned long long, unsigned int>::type Bits;
    const Bits* const bits = reinterpret_cast<const Bits*>(src);
    const Bits first = bits[0];
    if (j == 0 && blockIdx.y == 0) out[0] = src[0];
    bool diff = false;
    for (int i = blockIdx.y; i < rows; i += gridDim.y) diff |= bits[(long long)i * pitch + j] != first;
    if (diff) *differs = 1;
}

__device__ __forceinline__ unsigned long long splitmix64(unsigned long long x) {
    x += 0x9E3779B97F4A7C15ull;
    unsigned long long z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

// SURVEY.md 8(d) C5 recipe; same operation order as oracle.restated.synthetic_coefficients
template <typename T>
__global__ void synth_coeffs(T* epc, T* muc, long long pitch, int lrows, int cols, int row_off, int NR, long long NC,
                             unsigned long long seed, double dtds, double eps0, double mu0) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= cols) return;
    for (int l = blockIdx.y; l < lrows; l += gridDim.y) {
        const int i = row_off + l;
        if (i < 0 || i >= NR) continue;
        const unsigned long long bits = splitmix64(seed + (unsigned long long)i * (unsigned long long)NC + (unsigned long long)j);
        const double u = __dmul_rn((double)(bits >> 11), 1.1102230246251565e-16);      // 2^-53
        const double eps_r = __dadd_rn(1.0, __dmul_rn(3.0, u));
        const double eps = __dmul_rn(eps_r, eps0);
        const double mu = __dmul_rn(1.0, mu0);
        epc[(long long)l * pitch + j] = (T)__dmul_rn(dtds, __ddiv_rn(1.0, eps));
        muc[(long long)l * pitch + j] = (T)__dmul_rn(dtds, __ddiv_rn(1.0, mu));
    }
}

struct ShapeDev { int kind, pad; long long a[6]; double eps, mu; };

__device__ __forceinline__ bool convex_hit(const ShapeDev& s, long long r, long long c) {
    // python: for j in range(-rad,-disp): cell (ci+i, cj+j+disp); for j in range(disp,rad): cell (ci+i, cj+j-disp);
    // i in range(-rad,rad), i*i+j*j < rad*rad; an index k in [-n,n) addresses k mod n, anything else is skipped
    const long long ci = s.a[0], cj = s.a[1], rad = s.a[2], disp = s.a[3], n = s.a[4];
    for (int wr = 0; wr < 2; ++wr) {
        const long long i = (wr ? r - n : r) - ci;
        if (i < -rad || i >= rad) continue;
        for (int wc = 0; wc < 2; ++wc) {
            const long long cc = wc ? c - n : c;
            const long long j1 = cc - cj - disp;                  // first cap: j in [-rad, -disp)
            if (j1 >= -rad && j1 < -disp && i * i + j1 * j1 < rad * rad) return true;
            const long long j2 = cc - cj + disp;                  // second cap: j in [disp, rad)
            if (j2 >= disp && j2 < rad && i * i + j2 * j2 < rad * rad) return true;
        }
    }
    return false;
}

template <typename T>
__global__ void paint_coeffs(T* epc, T* muc, long long pitch, int lrows, int cols, int row_off, int NR,
                             const ShapeDev* __restrict__ shapes, int n_shapes, double eps_bg, double mu_bg, double dtds) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= cols) return;
    for (int l = blockIdx.y; l < lrows; l += gridDim.y) {
        const long long i = (long long)row_off + l;
        if (i < 0 || i >= NR) continue;
        double eps = eps_bg, mu = mu_bg;
        for (int k = 0; k < n_shapes; ++k) {
            const ShapeDev& s = shapes[k];
            bool hit;
            if (s.kind == 0) hit = i >= s.a[0] && i < s.a[1] && j >= s.a[2] && j < s.a[3];
            else if (s.kind == 1) {
                const long long x = i - s.a[0], y = (long long)j - s.a[1], d2 = x * x + y * y, in = s.a[2] - s.a[3];
                hit = s.a[2] * s.a[2] > d2 && d2 > in * in && x < 0 && y > 0;
            } else hit = convex_hit(s, i, j);
            if (hit) { eps = s.eps; mu = s.mu; }
        }
        epc[(long long)l * pitch + j] = (T)__dmul_rn(dtds, __ddiv_rn(1.0, eps));
        muc[(long long)l * pitch + j] = (T)__dmul_rn(dtds, __ddiv_rn(1.0, mu));
    }
}

}  // namespace fdtd
