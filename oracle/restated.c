/* TEST INFRASTRUCTURE ONLY -- plain-C restatement of one call of the reference's Solver.update
 * (/root/reference/solver.py:169-257), statement by statement, for parity checks on grids where the numpy
 * restatement (oracle/restated.py) gets slow.  Never linked into or called by the product.
 *
 * Pinned bitwise against oracle/restated.py (tests/test_c_oracle.py), which is itself pinned bitwise against the
 * live reference.  Build with -ffp-contract=off (no FMA) so that every + - * is rounded once, like a numpy ufunc:
 *     gcc -O2 -ffp-contract=off -fPIC -shared [-fopenmp] oracle/restated.c -o oracle/_build/liboracle.so
 *
 * Arrays are C-order doubles: h (NR,NC), ex (NR+1,NC), ey (NR,NC+1), hn (NR,NC) scratch/out, coefficients (NR,NC).
 * On return: hn holds the new h, ex/ey are updated in place, and h carries the hard point-source writes of
 * solver.py:176-180 (the reference mutates the OLD array, which is what saved snapshots alias).
 */
#include <stddef.h>

typedef struct {
    int kind;            /* 0 point, 1 right, 2 left, 3 up, 4 down (Pulse.get_direction) */
    long row, col;
    double mag;          /* pulse.magnitude(step) */
    double magz;         /* pulse.magnitude(step) * sqrt(MU/EPSILON) */
} oracle_src;

static long clip(long v, long len) {              /* python slice.indices(), step 1 */
    if (v < 0) { v += len; if (v < 0) v = 0; }
    if (v > len) v = len;
    return v;
}

static void zero_rect(double* a, long rows, long cols, const long* r) {
    const long r0 = clip(r[0], rows), r1 = clip(r[1], rows), c0 = clip(r[2], cols), c1 = clip(r[3], cols);
    for (long i = r0; i < r1; ++i)
        for (long j = c0; j < c1; ++j) a[i * cols + j] = 0.0;
}

void oracle_step(long NR, long NC, double* h, double* ex, double* ey, double* hn, const double* muc, const double* epc,
                 double S, double oms, const int* reflect, int n_src, const oracle_src* src, double last_mag,
                 int n_rect, const long* rects) {
    const long EC = NC + 1;                        /* ey row length */
#define H(i, j) h[(i) * NC + (j)]
#define HN(i, j) hn[(i) * NC + (j)]
#define EX(i, j) ex[(i) * NC + (j)]
#define EY(i, j) ey[(i) * EC + (j)]
#define MU_(i, j) muc[(i) * NC + (j)]
#define EP_(i, j) epc[(i) * NC + (j)]
    /* :176-180 hard point sources into the OLD h */
    for (int s = 0; s < n_src; ++s)
        if (src[s].kind == 0) H(src[s].row, src[s].col) = src[s].mag;
    /* :182 */
#pragma omp parallel for schedule(static)
    for (long i = 0; i < NR; ++i)
        for (long j = 0; j < NC; ++j)
            HN(i, j) = H(i, j) + MU_(i, j) * ((EX(i + 1, j) - EX(i, j)) - (EY(i, j + 1) - EY(i, j)));
    /* :184-200 TF/SF fix-ups in list order; left/up/down use the stale `magnitude` (last active pulse) */
    for (int s = 0; s < n_src; ++s) {
        const long c = src[s].col, r = src[s].row;
        if (src[s].kind == 1)
            for (long i = 0; i < NR; ++i)
                HN(i, c) = H(i, c) + MU_(i, c) * ((EX(i + 1, c) - EX(i, c)) - (EY(i, c + 1) - (EY(i, c) + src[s].magz)));
        else if (src[s].kind == 2)
            for (long i = 0; i < NR; ++i) HN(i, c) = HN(i, c) + MU_(i, c) * last_mag;
        else if (src[s].kind == 3)
            for (long j = 0; j < NC; ++j) HN(r + 1, j) = HN(r + 1, j) - MU_(r + 1, j) * last_mag;
        else if (src[s].kind == 4)
            for (long j = 0; j < NC; ++j) HN(r, j) = HN(r, j) - MU_(r, j) * last_mag;
    }
    /* :164-167, :203-204 */
    for (int k = 0; k < n_rect; ++k) zero_rect(hn, NR, NC, rects + 4 * k);
    /* :207-208 (rows 0, NR of ex and columns 0, NC of ey untouched) */
#pragma omp parallel for schedule(static)
    for (long i = 1; i < NR; ++i)
        for (long j = 0; j < NC; ++j) EX(i, j) = EX(i, j) + EP_(i, j) * (HN(i, j) - HN(i - 1, j));
#pragma omp parallel for schedule(static)
    for (long i = 0; i < NR; ++i)
        for (long j = 1; j < NC; ++j) EY(i, j) = EY(i, j) - EP_(i, j) * (HN(i, j) - HN(i, j - 1));
    /* :211-217 "right" re-applies on the already-updated ey; h[:, col-1] wraps at col 0 */
    for (int s = 0; s < n_src; ++s)
        if (src[s].kind == 1) {
            const long c = src[s].col, cm = (c == 0) ? NC - 1 : c - 1;
            for (long i = 0; i < NR; ++i) EY(i, c) = EY(i, c) - EP_(i, c) * (HN(i, c) - src[s].mag - HN(i, cm));
        }
    /* :220-235 absorbing walls, order up, down, left, right; E rows/columns read the CURRENT arrays */
    if (!reflect[0]) {
        for (long j = 0; j < NC; ++j) HN(0, j) = H(0, j) * oms + H(1, j) * S;
        for (long j = 0; j < NC; ++j) EX(0, j) = EX(0, j) * oms + EX(1, j) * S;
        for (long j = 0; j < EC; ++j) EY(0, j) = EY(0, j) * oms + EY(1, j) * S;
    }
    if (!reflect[1]) {
        for (long j = 0; j < NC; ++j) HN(NR - 1, j) = H(NR - 1, j) * oms + H(NR - 2, j) * S;
        for (long j = 0; j < NC; ++j) EX(NR, j) = EX(NR, j) * oms + EX(NR - 1, j) * S;
        for (long j = 0; j < EC; ++j) EY(NR - 1, j) = EY(NR - 1, j) * oms + EY(NR - 2, j) * S;
    }
    if (!reflect[2]) {
        for (long i = 0; i < NR; ++i) HN(i, 0) = H(i, 0) * oms + H(i, 1) * S;
        for (long i = 0; i <= NR; ++i) EX(i, 0) = EX(i, 0) * oms + EX(i, 1) * S;
        for (long i = 0; i < NR; ++i) EY(i, 0) = EY(i, 0) * oms + EY(i, 1) * S;
    }
    if (!reflect[3]) {
        for (long i = 0; i < NR; ++i) HN(i, NC - 1) = H(i, NC - 1) * oms + H(i, NC - 2) * S;
        for (long i = 0; i <= NR; ++i) EX(i, NC - 1) = EX(i, NC - 1) * oms + EX(i, NC - 2) * S;
        for (long i = 0; i < NR; ++i) EY(i, NC) = EY(i, NC) * oms + EY(i, NC - 1) * S;
    }
    /* :238-242 */
    for (int k = 0; k < n_rect; ++k) {
        zero_rect(ex, NR + 1, NC, rects + 4 * k);
        zero_rect(ey, NR, EC, rects + 4 * k);
    }
#undef H
#undef HN
#undef EX
#undef EY
#undef MU_
#undef EP_
}
