// stepK_edge: the K-step register pipeline of stepK_stream / stepK_lean with EVERY feature of Solver.update
// (/root/reference/solver.py:169-257) folded in as tile predicates -- absorbing / reflecting walls on all four sides
// (incl. the Gauss-Seidel order of the up/down walls and the corner precedence), hard point sources, TF/SF lines of all
// four directions with the stale `magnitude`, reflector rectangles and energy probes.
//
// The plain K-step kernels advance only tiles whose K-step dependency cone touches nothing special; this kernel takes
// all the other tiles of a temporal-blocking group (the wall bands, the wall / TF-SF columns, source and reflector
// tiles), so that a group is "plain tiles + edge tiles" of ONE pass over generation A and no cell needs a separate
// single-step launch chain.  It is the boundary-tile side of the same algorithm: same lane geometry (32 lanes x V
// columns loaded, the outer ceil(K/V) lanes on each side sacrificial), same stage pipeline, same arithmetic in the same
// order -- plus predicates.  Edge tiles are ~1 % of a large grid, so this variant trades speed for generality: plain
// guarded loads instead of the cp.async ring, and it keeps one more row of state per stage.
//
// Pipeline.  Iteration t loads row t+1 ("stage 0") and runs stage s = 1..K on row r = t-s+1.  Stage s reads
//   own row r      : L[s-1] = (h, ex, ey) after step s-1, C[s-1] = (mu, ep) of row r
//   the row below  : the FRESH outputs of stage s-1 this iteration, (h, ex) of row r+1
//   the row above  : its own hn_pre(r-1) of the previous iteration (hnp[s])
// and produces the final (h, ex, ey) of row r after step s.  The reference's wall statements read neighbours that are
// already updated (Gauss-Seidel), which in marching order means:
//   up wall   (solver.py:221-223)  h[0] needs h_old[1] (the fresh h of the row below: available); ex[0], ey[0] need the
//             NEW ex[1], ey[1]: row 0's E values wait in L[s] and are finished when the stage is at row 1, before any
//             consumer (stage s+1, or the store behind stage K) reads them;
//   down wall (:225-227)           h[NR-1], ey[NR-1], ex[NR] need values of the row above: each stage carries the
//             previous row's h_old / ex_pre / ey_pre;
//   left/right walls (:229-235)    read the up/down results of the neighbour column: applied after them (for row 0: at
//             the deferred point), neighbours by shuffle; reflector E zeroing (:238-242) after that.
// Rows < 0 and > NR and columns outside the arrays are computed on zeros and never stored; the masks of the update
// ranges (ex rows 1..NR-1, ey columns 1..NC-1) keep them out of every value that is.  Which rows a tile must march
// through so that stage K is exact on its own rows: DESIGN.md 5.
//
// Not handled (the host keeps such scenes off the K-step path): a TF/SF "right" line at column 0 (python's
// h[:, col-1] wraps to the far column).
//
// Tile classes.  A wall band needs no source loop, a wall column no row-edge logic, and the generic version costs about
// 2500 instructions per row and spills (ncu, profiles/r2_ncu_full_stepK_edge_f64_k8_raw.csv: 1.15 ms of SM slot time per
// K = 8 group on the C5 slab, a sixth of the group).  The kernel is therefore compiled per feature set F (EdgeFlags):
// the host planner lists, per tile, which features its K-step cone touches and launches each class with the cheapest
// instantiation that covers it.  Leaving a feature out is exact, not approximate: it removes predicates that are false
// for every cell the tile's stored values depend on.
#pragma once
#include "kernels.cuh"
#include "resident.cuh"

namespace fdtd {

struct EdgeTile { int tx, ta, tb, pad; };     // column tile (jw = tx * OUT - MARGIN * V), rows [ta, tb) of the index space

// what a tile's K-step cone touches (one kernel instantiation per combination the planner uses)
enum EdgeFlags : int {
    EDGE_ROWS = 1,     // a first / last row of the grid: rows 0, NR-1, NR (up / down walls, the update ranges of ex)
    EDGE_COLS = 2,     // a first / last column: columns 0, NC-1, NC (left / right walls, the update range of ey)
    EDGE_SRC = 4,      // a live pulse (point source cell, TF/SF line)
    EDGE_MISC = 8,     // reflector rectangles, probes
    EDGE_ALL = 15
};

// The instantiations that exist (cheapest first) and the one that serves a tile whose cone touches `need`.
constexpr int EDGE_CLASSES[5] = {EDGE_COLS, EDGE_ROWS, EDGE_SRC, EDGE_COLS | EDGE_SRC, EDGE_ALL};
inline int edge_class_for(int need) {
    for (int c : EDGE_CLASSES) if ((need & ~c) == 0) return c;
    return EDGE_ALL;
}

template <typename T>
struct EdgeRec {                               // the pulses ACTIVE in one call, in list order (what fill_params builds)
    int n_src, pad;
    T last_mag;
    SrcStep<T> src[FDTD_MAX_SRC];
};

// per-lane state of the down wall (solver.py:225-227 reads the row above): ex_pre, ey_pre and h_old of the previous row
// of every stage.  They are needed for one row per stage and tile, so they live in shared memory and are written only
// when the stage is at the rows in question (in registers they cost 3 (K+1) V values and made K = 8 spill).
template <typename T, int V, int K>
struct EdgeWallState { T ex[K + 1][V][32], ey[K + 1][V][32], ho[K + 1][V][32]; };

#if defined(FDTD_HOST_EMUL)
inline void edge_syncwarp() { (void)__shfl_sync(0xffffffffu, 0, 0); }       // the emulator's shuffle is a warp barrier
#else
__device__ __forceinline__ void edge_syncwarp() { __syncwarp(); }
#endif

// SPLIT = 1 (T = double): the fields live in split storage (kernels.cuh: fp32 hi plane + bf16 lo plane, dtype FDTD_F32X).
// Loads decode, stores encode, and every value a call would have stored is rounded to the storage format on its way to
// the next stage (split_round), so that a group of K calls gives exactly the bits of K single calls.
template <typename T, int V, int K, int F, int SPLIT = 0>
__device__ __forceinline__ void run_tile_edge(const StepParams<T>& P, const EdgeRec<T>* rec, EdgeWallState<T, V, K>* W,
                                              int ta, int tb, int jw, int lane) {
    constexpr bool ROWS = (F & EDGE_ROWS) != 0, COLS = (F & EDGE_COLS) != 0, SRC = (F & EDGE_SRC) != 0,
                   MISC = (F & EDGE_MISC) != 0;
    static_assert(!SPLIT || sizeof(T) == 8, "split storage computes in double");
    const unsigned full = 0xffffffffu;
    constexpr int MARGIN = KGeom<K, V>::MARGIN;
    const int NR = P.NR, NC = P.NC;
    const int j = jw + lane * V;
    const long long pitch = P.pitch;
    const bool col_in = j >= 0 && j + V <= pitch;                 // j and pitch are multiples of V: all or nothing
    // The lanes of the right margin are sacrificial because what lies right of the loaded columns is unknown -- unless
    // nothing lies there: a tile whose loaded columns reach past column NC owns its right margin as well (and the planner
    // lists no tile right of it).  This also gives the right wall the K+1 columns its Gauss-Seidel reads need on its left.
    const bool storing = lane >= MARGIN && (lane < 32 - MARGIN || (COLS && jw + 32 * V > NC));
    const T S = P.S, oms = P.oms;
    // does a wall column lie among the columns this warp loads?  (tile-uniform: no shuffles for walls that are elsewhere)
    const bool wl = COLS && P.absorb_l && jw <= 0, wr = COLS && P.absorb_r && jw + 32 * V > NC - 1 && jw <= NC;

    T Lh[K + 1][V], Lex[K + 1][V], Ley[K + 1][V];   // L[s]: stage s outputs of the row it processed last iteration (s = 0: loaded)
    T Cmu[K][V], Cep[K][V];                         // C[s]: coefficients of the row stage s+1 processes this iteration
    T hnp[K + 1][V];                                // hn_pre of stage s at its previous row
#pragma unroll
    for (int s = 0; s <= K; ++s)
#pragma unroll
        for (int v = 0; v < V; ++v) {
            Lh[s][v] = Lex[s][v] = Ley[s][v] = hnp[s][v] = T(0);
            if (s < K) Cmu[s][v] = Cep[s][v] = T(0);
            if (ROWS) W->ex[s][v][lane] = W->ey[s][v][lane] = W->ho[s][v][lane] = T(0);
        }

    auto row_in = [&](int i) { return i >= 0 && i <= NR && i >= P.row_off && i < P.row_off + P.lrows; };
    auto ldrow = [&](T (&r)[V], const T* a, int i) {          // a coefficient row
        if (col_in && row_in(i)) ldvec<T, V>(r, a + (long long)(i - P.row_off) * pitch + j);
        else {
#pragma unroll
            for (int v = 0; v < V; ++v) r[v] = T(0);
        }
    };
    auto ldfield = [&](T (&r)[V], const T* a, int i) {        // a row of h / ex / ey
        if (!SPLIT) { ldrow(r, a, i); return; }
        if (col_in && row_in(i)) {
            const long long at = (long long)(i - P.row_off) * pitch + j;
            const float* hi = reinterpret_cast<const float*>(a) + at;
            const unsigned short* lo = reinterpret_cast<const unsigned short*>(reinterpret_cast<const char*>(a) + P.split_lo_off) + at;
#pragma unroll
            for (int v = 0; v < V; ++v) r[v] = (T)split_decode(hi[v], lo[v]);
        } else {
#pragma unroll
            for (int v = 0; v < V; ++v) r[v] = T(0);
        }
    };
    auto stored = [&](T (&x)[V]) {                            // a call's final values, as the next call will read them
        if (SPLIT) {
#pragma unroll
            for (int v = 0; v < V; ++v) x[v] = (T)split_round((double)x[v]);
        }
    };
    // value of the column to the right / left of element v (own register or the neighbour lane's)
    auto right_of = [&](const T (&x)[V], T (&out)[V]) {
        const T nb = __shfl_down_sync(full, x[0], 1);
#pragma unroll
        for (int v = 0; v < V; ++v) out[v] = (v + 1 < V) ? x[v + 1 < V ? v + 1 : 0] : nb;
    };
    auto left_of = [&](const T (&x)[V], T (&out)[V]) {
        const T nb = __shfl_up_sync(full, x[V - 1], 1);
#pragma unroll
        for (int v = 0; v < V; ++v) out[v] = (v == 0) ? nb : x[v > 0 ? v - 1 : 0];
    };
    // left / right walls of ex (X) and ey (Y) of row i (solver.py:229-235: they read the up/down results of the
    // neighbour column), then the reflector zeroing of :238-242
    auto lr_and_rects = [&](T (&X)[V], T (&Y)[V], int i) {
        if (COLS && (wl || wr)) {
            T Xr[V], Yr[V], Xl[V], Yl[V];
            right_of(X, Xr); right_of(Y, Yr); left_of(X, Xl); left_of(Y, Yl);
#pragma unroll
            for (int v = 0; v < V; ++v) {
                const int c = j + v;
                if (wl && c == 0) { X[v] = add(mul(X[v], oms), mul(Xr[v], S)); Y[v] = add(mul(Y[v], oms), mul(Yr[v], S)); }
                else {
                    if (wr && c == NC - 1) X[v] = add(mul(X[v], oms), mul(Xl[v], S));
                    if (wr && c == NC) Y[v] = add(mul(Y[v], oms), mul(Yl[v], S));
                }
            }
        }
        if (MISC) {
#pragma unroll 1
            for (int q = 0; q < P.n_rect; ++q)
#pragma unroll
                for (int v = 0; v < V; ++v) {
                    if (inside(P.rex[q], i, j + v)) X[v] = T(0);
                    if (inside(P.rey[q], i, j + v)) Y[v] = T(0);
                }
        }
    };

    // rows the tile marches through: K above its first row for the K stages, one more because the down wall reads TWO
    // rows up on its first hop (ex[NR] <- ex_pre[NR-1] <- hn[NR-2]); below, the delayed store needs iteration tb+K-1, which
    // also gives the up wall its second row (ex[0] <- ex_pre[1] <- ex[2])
    const int t_first = max(ta - K - 1, 0), t_last = tb + K - 1;
    ldfield(Lh[0], P.h, t_first); ldfield(Lex[0], P.ex, t_first); ldfield(Ley[0], P.ey, t_first);
    ldrow(Cmu[0], P.muc, t_first); ldrow(Cep[0], P.epc, t_first);
    // the row below stage 1's row is loaded one iteration ahead of its use, so its latency hides behind K stages of work
    T nh[V], nex[V], ney[V], nmu[V], nep[V];
    ldfield(nh, P.h, t_first + 1); ldfield(nex, P.ex, t_first + 1); ldfield(ney, P.ey, t_first + 1);
    ldrow(nmu, P.muc, t_first + 1); ldrow(nep, P.epc, t_first + 1);

    for (int t = t_first; t <= t_last; ++t) {
        // "stage 0": the row below stage 1's row
        T ph[V], pex[V], pey[V], fmu[V], fep[V];
#pragma unroll
        for (int v = 0; v < V; ++v) { ph[v] = nh[v]; pex[v] = nex[v]; pey[v] = ney[v]; fmu[v] = nmu[v]; fep[v] = nep[v]; }
        if (t < t_last) {
            ldfield(nh, P.h, t + 2); ldfield(nex, P.ex, t + 2); ldfield(ney, P.ey, t + 2);
            ldrow(nmu, P.muc, t + 2); ldrow(nep, P.epc, t + 2);
        }
#pragma unroll
        for (int s = 1; s <= K; ++s) {
            const int r = t - (s - 1);                              // the row stage s works on
            const EdgeRec<T>& Q = rec[SRC ? s - 1 : 0];
            const T (&h)[V] = Lh[s - 1];
            const T (&exc)[V] = Lex[s - 1];
            const T (&ey)[V] = Ley[s - 1];
            const T (&mu)[V] = Cmu[s - 1];
            const T (&ep)[V] = Cep[s - 1];
            // hard point sources of this call (solver.py:176-180) on the own row and on the row below
            T ho[V], hb[V];
#pragma unroll
            for (int v = 0; v < V; ++v) { ho[v] = h[v]; hb[v] = ph[v]; }
            if (SRC) {
#pragma unroll 1
                for (int k = 0; k < Q.n_src; ++k) {
                    const SrcStep<T>& q = Q.src[k];
                    if (q.kind != 0) continue;
#pragma unroll
                    for (int v = 0; v < V; ++v)
                        if (q.col == j + v) { if (q.row == r) ho[v] = q.mag; if (q.row == r + 1) hb[v] = q.mag; }
                }
            }
            // new h before the walls (:182), TF/SF fix-ups in list order (:184-200), reflector zeroing (:203-204)
            T eyr[V], dex[V], hn[V];
            right_of(ey, eyr);
#pragma unroll
            for (int v = 0; v < V; ++v) {
                dex[v] = sub(pex[v], exc[v]);
                hn[v] = add(ho[v], mul(mu[v], sub(dex[v], sub(eyr[v], ey[v]))));
            }
            if (SRC) {
#pragma unroll 1
                for (int k = 0; k < Q.n_src; ++k) {
                    const SrcStep<T>& q = Q.src[k];
#pragma unroll
                    for (int v = 0; v < V; ++v) {
                        const int c = j + v;
                        if (q.kind == 1) { if (c == q.col) hn[v] = add(ho[v], mul(mu[v], sub(dex[v], sub(eyr[v], add(ey[v], q.magz))))); }
                        else if (q.kind == 2) { if (c == q.col) hn[v] = add(hn[v], mul(mu[v], Q.last_mag)); }
                        else if (q.kind == 3) { if (r == q.row + 1) hn[v] = sub(hn[v], mul(mu[v], Q.last_mag)); }
                        else if (q.kind == 4) { if (r == q.row) hn[v] = sub(hn[v], mul(mu[v], Q.last_mag)); }
                    }
                }
            }
            if (MISC) {
#pragma unroll 1
                for (int q = 0; q < P.n_rect; ++q)
#pragma unroll
                    for (int v = 0; v < V; ++v) if (inside(P.rh[q], r, j + v)) hn[v] = T(0);
            }
            // E update (:207-208) with its ranges, TF/SF "right" re-application on ey (:211-217)
            T hl[V], exq[V], eyq[V];
            left_of(hn, hl);
            const bool ex_row = !ROWS || (r >= 1 && r <= NR - 1);
#pragma unroll
            for (int v = 0; v < V; ++v) {
                const int c = j + v;
                exq[v] = ex_row ? add(exc[v], mul(ep[v], sub(hn[v], hnp[s][v]))) : exc[v];
                eyq[v] = (!COLS || (c >= 1 && c <= NC - 1)) ? sub(ey[v], mul(ep[v], sub(hn[v], hl[v]))) : ey[v];
            }
            if (SRC) {
#pragma unroll 1
                for (int k = 0; k < Q.n_src; ++k) {
                    const SrcStep<T>& q = Q.src[k];
                    if (q.kind != 1) continue;
#pragma unroll
                    for (int v = 0; v < V; ++v)
                        if (q.col == j + v && q.col > 0) eyq[v] = sub(eyq[v], mul(ep[v], sub(sub(hn[v], q.mag), hl[v])));
                }
            }
            T fx[V], fy[V], fh[V];
            if (ROWS) {
                // row 0's E values (waiting in L[s]) are finished now that this stage has the new ex[1], ey[1] (:222-223)
                if (r == 1) {
                    if (P.absorb_u) {
#pragma unroll
                        for (int v = 0; v < V; ++v) {
                            Lex[s][v] = add(mul(Lex[s][v], oms), mul(exq[v], S));
                            Ley[s][v] = add(mul(Ley[s][v], oms), mul(eyq[v], S));
                        }
                    }
                    lr_and_rects(Lex[s], Ley[s], 0);
                    stored(Lex[s]); stored(Ley[s]);
                }
                // own row: down wall (:226-227) from what this stage left in W at the row above
#pragma unroll
                for (int v = 0; v < V; ++v) { fx[v] = exq[v]; fy[v] = eyq[v]; }
                if (P.absorb_d && r == NR) {
#pragma unroll
                    for (int v = 0; v < V; ++v) fx[v] = add(mul(exc[v], oms), mul(W->ex[s][v][lane], S));
                }
                if (P.absorb_d && r == NR - 1) {
#pragma unroll
                    for (int v = 0; v < V; ++v) fy[v] = add(mul(eyq[v], oms), mul(W->ey[s][v][lane], S));
                }
                // ... then left/right walls and reflectors (row 0 waits for the next iteration)
                if (r != 0) lr_and_rects(fx, fy, r);
            } else {
#pragma unroll
                for (int v = 0; v < V; ++v) { fx[v] = exq[v]; fy[v] = eyq[v]; }
                lr_and_rects(fx, fy, r);
            }
            // h walls: all from the old h, later walls win (:221, :225, :229, :233)
#pragma unroll
            for (int v = 0; v < V; ++v) fh[v] = hn[v];
            if (ROWS) {
                if (P.absorb_u && r == 0) {
#pragma unroll
                    for (int v = 0; v < V; ++v) fh[v] = add(mul(ho[v], oms), mul(hb[v], S));
                }
                if (P.absorb_d && r == NR - 1) {
#pragma unroll
                    for (int v = 0; v < V; ++v) fh[v] = add(mul(ho[v], oms), mul(W->ho[s][v][lane], S));
                }
            }
            if (COLS && (wl || wr)) {
                T hor[V], hol[V];
                right_of(ho, hor); left_of(ho, hol);
#pragma unroll
                for (int v = 0; v < V; ++v) {
                    const int c = j + v;
                    if (wl && c == 0) fh[v] = add(mul(ho[v], oms), mul(hor[v], S));
                    if (wr && c == NC - 1) fh[v] = add(mul(ho[v], oms), mul(hol[v], S));
                }
            }
            // fh, fx, fy are what this call leaves in the arrays for row r (row 0's E values: one iteration later, above)
            stored(fh);
            if (!ROWS || r != 0) { stored(fx); stored(fy); }
            // L[s] holds the FINAL values of row r-1 after call s: probes (:246-255), then its consumer -- stage s+1
            // in the next trip of this loop, or the store below
            if (MISC && P.n_probe && storing && r - 1 >= ta && r - 1 < tb) {
#pragma unroll 1
                for (int k = 0; k < P.n_probe; ++k)
#pragma unroll
                    for (int v = 0; v < V; ++v)
                        if (P.probe_i[k] == r - 1 && P.probe_j[k] == j + v) {
                            double* o = P.probe_out + ((size_t)(s - 1) * P.n_probe + k) * 3;
                            o[0] = (double)Lh[s][v]; o[1] = (double)Lex[s][v]; o[2] = (double)Ley[s][v];
                        }
            }
            // what the down wall of the NEXT row will read from this one
            if (ROWS && P.absorb_d) {
                if (r == NR - 1) {
#pragma unroll
                    for (int v = 0; v < V; ++v) W->ex[s][v][lane] = exq[v];
                }
                if (r == NR - 2) {
#pragma unroll
                    for (int v = 0; v < V; ++v) { W->ey[s][v][lane] = eyq[v]; W->ho[s][v][lane] = ho[v]; }
                }
            }
            // stage s has consumed L[s-1] / C[s-1]: they take the previous stage's fresh row; ours becomes the fresh one
#pragma unroll
            for (int v = 0; v < V; ++v) {
                hnp[s][v] = hn[v];
                Lh[s - 1][v] = ph[v]; Lex[s - 1][v] = pex[v]; Ley[s - 1][v] = pey[v];
                ph[v] = fh[v]; pex[v] = fx[v]; pey[v] = fy[v];
            }
        }
        // the store behind stage K: row t-K, final in L[K]
        const int q = t - K;
        if (q >= ta && q < tb && storing) {
            T* const hq = P.hn + (long long)(q - P.row_off) * pitch;
            T* const exq_ = P.exn + (long long)(q - P.row_off) * pitch;
            T* const eyq_ = P.eyn + (long long)(q - P.row_off) * pitch;
            if (SPLIT) {
                const long long at = (long long)(q - P.row_off) * pitch;
                auto put = [&](T* base, int c, T x) {
                    float hi; unsigned short lo;
                    split_encode((double)x, hi, lo);
                    reinterpret_cast<float*>(base)[at + c] = hi;
                    reinterpret_cast<unsigned short*>(reinterpret_cast<char*>(base) + P.split_lo_off)[at + c] = lo;
                };
#pragma unroll
                for (int v = 0; v < V; ++v) {
                    const int c = j + v;
                    if (q < NR && c < NC) put(P.hn, c, Lh[K][v]);
                    if (c < NC) put(P.exn, c, Lex[K][v]);
                    if (q < NR && c <= NC) put(P.eyn, c, Ley[K][v]);
                }
            } else if (!ROWS && !COLS) {                         // every stored cell exists in all three arrays
                stvec<T, V>(hq + j, Lh[K]); stvec<T, V>(exq_ + j, Lex[K]); stvec<T, V>(eyq_ + j, Ley[K]);
            } else {
#pragma unroll
                for (int v = 0; v < V; ++v) {
                    const int c = j + v;
                    if (q < NR && c < NC) hq[c] = Lh[K][v];
                    if (c < NC) exq_[c] = Lex[K][v];
                    if (q < NR && c <= NC) eyq_[c] = Ley[K][v];
                }
            }
        }
#pragma unroll
        for (int v = 0; v < V; ++v) { Lh[K][v] = ph[v]; Lex[K][v] = pex[v]; Ley[K][v] = pey[v]; }
#pragma unroll
        for (int s = K - 1; s >= 1; --s)
#pragma unroll
            for (int v = 0; v < V; ++v) { Cmu[s][v] = Cmu[s - 1][v]; Cep[s][v] = Cep[s - 1][v]; }
#pragma unroll
        for (int v = 0; v < V; ++v) { Cmu[0][v] = fmu[v]; Cep[0][v] = fep[v]; }
    }
}

// One warp per edge tile.  P: generation A in h/ex/ey, generation B in hn/exn/eyn, probe_out = record of call
// first_step (call s of the group writes record s); R: the per-call pulse records of resident mode.
template <typename T, int V, int K, int F, int SPLIT = 0>
__global__ void __launch_bounds__(32) stepK_edge(const __grid_constant__ StepParams<T> P, const EdgeTile* __restrict__ tiles,
                                                 int n_tiles, const ResSources<T> R, long long first_step) {
#if defined(FDTD_HOST_EMUL)
    static EdgeRec<T> rec[K];
    static EdgeWallState<T, V, K> wall;
#else
    __shared__ EdgeRec<T> rec[(F & EDGE_SRC) ? K : 1];
    __shared__ EdgeWallState<T, V, (F & EDGE_ROWS) ? K : 0> wall_store;
    EdgeWallState<T, V, K>* const wall_p = reinterpret_cast<EdgeWallState<T, V, K>*>(&wall_store);
#endif
    const int lane = threadIdx.x & 31;
    if ((int)blockIdx.x >= n_tiles) return;
    if (F & EDGE_SRC) {
        for (int e = lane; e < K * FDTD_MAX_SRC; e += 32) {
            const int m = e / FDTD_MAX_SRC, k = e % FDTD_MAX_SRC;
            if (R.count && k < R.stride) rec[m].src[k] = R.src[(first_step + m) * R.stride + k];
        }
        if (lane < K) {
            rec[lane].n_src = R.count ? R.count[first_step + lane] : 0;
            rec[lane].pad = 0;
            rec[lane].last_mag = R.count ? R.last[first_step + lane] : T(0);
        }
        edge_syncwarp();
    }
    const EdgeTile tl = tiles[blockIdx.x];
#if defined(FDTD_HOST_EMUL)
    run_tile_edge<T, V, K, F, SPLIT>(P, rec, &wall, tl.ta, tl.tb, tl.tx * KGeom<K, V>::OUT - KGeom<K, V>::MARGIN * V, lane);
#else
    run_tile_edge<T, V, K, F, SPLIT>(P, rec, wall_p, tl.ta, tl.tb, tl.tx * KGeom<K, V>::OUT - KGeom<K, V>::MARGIN * V, lane);
#endif
}

}  // namespace fdtd
