// =====================================================================================================
// TEST INFRASTRUCTURE ONLY.  CPU oracle = a from-scratch restatement of the reference's LBA arithmetic.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
// The product (pl_slam_plucker_b200/csrc) never links, imports or calls anything in this directory.
//
// PARITY PARTLY PINNED.  kongan/PL-SLAM-plucker ships no test, golden vector, fixture or recorded log for its LBA
//  and the application cannot be built here (Eigen, g2o, OpenCV, Boost, yaml-cpp, MRPT absent).
//   PINNED on the reference's own code: the vertex / edge arithmetic of profile G and the MapLine orthonormal helpers
//   (rows a6-a11, a17-a19 of SURVEY.md §8a) — g2o_types/g2o_types.h and src/mapFeatures.cpp compile UNMODIFIED against the
//   stand-in headers of oracle/ref_shim/ into oracle/_ref/libref_g2o_types.so; refmath.h agrees with it bit for bit on
//   seeded inputs (tests/test_ref_pin.py, committed vectors tests/golden/ref_g2o_types.npz).
//   UNPINNED (no reference artefact exists): the g2o Levenberg / Schur shell (g2o is not under /root/reference and no version
//   is pinned) and the hand-LM functions of src/mapHandler.cpp (OpenCV-entangled): those are pinned only by self-made checks
//   (finite differences, round trips, Schur == dense solve).
//
// What is restated, with the reference lines followed:
//   profile G     MapHandler::localBundleAdjustmentForPlukerWithG2O   src/mapHandler.cpp:5851-6323
//                 + vertices / edges                                  g2o_types/g2o_types.h:28-453
//                 + g2o core semantics (external, un-vendored, no version pinned):
//                   OptimizationAlgorithmLevenberg::solve/computeLambdaInit/computeScale, BlockSolver::buildSystem/
//                   setLambda/solve(Schur), BaseBinaryEdge::constructQuadraticForm, RobustKernelHuber::robustify,
//                   SparseOptimizer::optimize/initializeOptimization(level)
//   profile H_END MapHandler::levMarquardtOptimizationLBA             src/mapHandler.cpp:2334-3016
//   profile H_PLK MapHandler::levMarquardtOptimizationLBAForPluker    src/mapHandler.cpp:1618-2332
// The reduced camera system is solved by an envelope (skyline) LDL^T without pivoting: the same factorisation
// class as Eigen::SimplicialLDLT used by both reference paths (results agree to rounding).
// =====================================================================================================
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <algorithm>
#include <limits>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "../include/plba.h"
#include "refmath.h"

using namespace oracle;

namespace {

int g_threads = 1;
int nthreads() {
#ifdef _OPENMP
    return g_threads;
#else
    return 1;
#endif
}

// ---------------------------------------------------------------------------------------------------
// Block system shared by both profiles:  [Hpp  W ; W^T  Hll] [xp; xl] = [bp; bl]
// ---------------------------------------------------------------------------------------------------
struct BlockSystem {
    int n_free = 0, n_pt = 0, n_ls = 0, dl = 4;  // dl: line landmark dimension (4 orth, 6 endpoints)
    std::vector<double> Hpp, bp;                 // [n_free][36], [n_free][6]
    std::vector<double> Hll_pt, bl_pt;           // [n_pt][9], [n_pt][3]
    std::vector<double> Hll_ls, bl_ls;           // [n_ls][dl*dl], [n_ls][dl]
    // per edge: pose slot (-1 fixed / inactive) and W block (6 x d, row-major)
    std::vector<int> slot_p, slot_l;
    std::vector<double> W_p, W_l;                // [n_pobs][18], [n_lobs][6*dl]
    std::vector<int> ptr_p, ptr_l;               // CSR landmark -> edges
    void zero() {
        std::fill(Hpp.begin(), Hpp.end(), 0.0); std::fill(bp.begin(), bp.end(), 0.0);
        std::fill(Hll_pt.begin(), Hll_pt.end(), 0.0); std::fill(bl_pt.begin(), bl_pt.end(), 0.0);
        std::fill(Hll_ls.begin(), Hll_ls.end(), 0.0); std::fill(bl_ls.begin(), bl_ls.end(), 0.0);
        std::fill(W_p.begin(), W_p.end(), 0.0); std::fill(W_l.begin(), W_l.end(), 0.0);
        std::fill(slot_p.begin(), slot_p.end(), -1); std::fill(slot_l.begin(), slot_l.end(), -1);
    }
};

struct Skyline {
    int n = 0;
    std::vector<int> first;        // first stored column of row i
    std::vector<int64_t> rowptr;   // start of row i in val ; entry (i,j) at rowptr[i] + (j-first[i])
    std::vector<double> val;
    double &at(int i, int j) { return val[rowptr[i] + (j - first[i])]; }
};

static std::vector<int> csr_from_sorted(const int32_t *lm, int n_obs, int n_lm) {
    std::vector<int> ptr(n_lm + 1, 0);
    for (int i = 0; i < n_obs; i++) ptr[lm[i] + 1]++;
    for (int i = 0; i < n_lm; i++) ptr[i + 1] += ptr[i];
    return ptr;
}

static void build_skyline(const plba_problem &P, const std::vector<int> &ptr_p, const std::vector<int> &ptr_l, Skyline &S) {
    int nb = P.n_free;
    std::vector<int> minslot(nb);
    for (int s = 0; s < nb; s++) minslot[s] = s;
    auto scan = [&](const std::vector<int> &ptr, const int32_t *kf, int n_lm) {
        for (int l = 0; l < n_lm; l++) {
            int mn = INT32_MAX;
            for (int e = ptr[l]; e < ptr[l + 1]; e++) { int s = P.kf_slot[kf[e]]; if (s >= 0) mn = std::min(mn, s); }
            for (int e = ptr[l]; e < ptr[l + 1]; e++) { int s = P.kf_slot[kf[e]]; if (s >= 0) minslot[s] = std::min(minslot[s], mn); }
        }
    };
    scan(ptr_p, P.po_kf, P.n_pt);
    scan(ptr_l, P.lo_kf, P.n_ls);
    S.n = 6 * nb;
    S.first.resize(S.n); S.rowptr.resize(S.n + 1);
    int64_t off = 0;
    for (int i = 0; i < S.n; i++) { S.first[i] = 6 * minslot[i / 6]; S.rowptr[i] = off; off += (i - S.first[i] + 1); }
    S.rowptr[S.n] = off;
    S.val.assign(off, 0.0);
}

// LDL^T (no pivoting) inside the envelope, then solve.  Returns false on a zero pivot.
static bool skyline_ldlt_solve(Skyline &S, std::vector<double> &b) {
    int n = S.n;
    std::vector<double> d(n);
    for (int i = 0; i < n; i++) {
        int fi = S.first[i];
        double *ri = &S.val[S.rowptr[i]] - fi;   // ri[j] = entry (i,j)
        for (int j = fi; j < i; j++) {
            int fj = S.first[j];
            const double *rj = &S.val[S.rowptr[j]] - fj;
            double s = ri[j];
            for (int k = std::max(fi, fj); k < j; k++) s -= ri[k] * rj[k];   // ri[k] holds Y(i,k)=L(i,k)d_k, rj[k] holds L(j,k)
            ri[j] = s;
        }
        double di = ri[i];
        for (int j = fi; j < i; j++) { double y = ri[j]; double l = y / d[j]; di -= y * l; ri[j] = l; }
        d[i] = di;
        if (!(di != 0.0) || !std::isfinite(di)) return false;
    }
    // forward L y = b
    for (int i = 0; i < n; i++) { int fi = S.first[i]; const double *ri = &S.val[S.rowptr[i]] - fi; double s = b[i]; for (int j = fi; j < i; j++) s -= ri[j] * b[j]; b[i] = s; }
    for (int i = 0; i < n; i++) b[i] /= d[i];
    for (int i = n - 1; i >= 0; i--) { int fi = S.first[i]; const double *ri = &S.val[S.rowptr[i]] - fi; double xi = b[i]; for (int j = fi; j < i; j++) b[j] -= ri[j] * xi; }
    return true;
}

static void inv_small(const double *A, int d, double *out) {
    if (d == 3) { Mat<3, 3> M; for (int i = 0; i < 9; i++) M[i] = A[i]; Mat<3, 3> I = inverse(M); for (int i = 0; i < 9; i++) out[i] = I[i]; }
    else if (d == 4) { Mat<4, 4> M; for (int i = 0; i < 16; i++) M[i] = A[i]; Mat<4, 4> I = inverse(M); for (int i = 0; i < 16; i++) out[i] = I[i]; }
    else { Mat<6, 6> M; for (int i = 0; i < 36; i++) M[i] = A[i]; Mat<6, 6> I = inverse(M); for (int i = 0; i < 36; i++) out[i] = I[i]; }
}

// Landmark Schur complement + LDL^T + back-substitution (g2o BlockSolver::solve, Schur branch).
// multiplicative: H_ii *= (1+lambda) (profile H, src/mapHandler.cpp:2564-2565); else H_ii += lambda (g2o setLambda).
// If S_out / g_out are given the reduced system is exported (upper 6x6 blocks, dense n_free x n_free block grid).
static bool solve_schur(const plba_problem &P, const BlockSystem &B, Skyline &S, double lambda, bool multiplicative,
                        std::vector<double> &xp, std::vector<double> &xl_pt, std::vector<double> &xl_ls,
                        std::vector<double> *S_dense = nullptr, std::vector<double> *g_out = nullptr, bool solve = true) {
    const int nf = B.n_free;
    std::fill(S.val.begin(), S.val.end(), 0.0);
    for (int s = 0; s < nf; s++)
        for (int r = 0; r < 6; r++)
            for (int c = 0; c <= r; c++) {
                double v = B.Hpp[s * 36 + r * 6 + c];
                if (r == c) v = multiplicative ? v + lambda * v : v + lambda;
                S.at(6 * s + r, 6 * s + c) = v;
            }
    std::vector<double> coeff(6 * nf, 0.0);
    const int T = nthreads();
    std::vector<std::vector<double>> Sthr(T > 1 ? T : 0), Cthr(T > 1 ? T : 0);
    for (int t = 0; t < (int)Sthr.size(); t++) { Sthr[t].assign(S.val.size(), 0.0); Cthr[t].assign(6 * nf, 0.0); }

    auto process = [&](int n_lm, int d, const std::vector<int> &ptr, const std::vector<int> &slot, const std::vector<double> &W,
                       const std::vector<double> &Hll, const std::vector<double> &bl) {
#pragma omp parallel for schedule(static) num_threads(T) if (T > 1)
        for (int l = 0; l < n_lm; l++) {
            int tid = 0;
#ifdef _OPENMP
            tid = omp_get_thread_num();
#endif
            double *Sv = (T > 1) ? Sthr[tid].data() : S.val.data();
            double *Cv = (T > 1) ? Cthr[tid].data() : coeff.data();
            double D[36], Dinv[36], db[6];
            for (int i = 0; i < d * d; i++) D[i] = Hll[(size_t)l * d * d + i];
            for (int i = 0; i < d; i++) D[i * d + i] = multiplicative ? D[i * d + i] + lambda * D[i * d + i] : D[i * d + i] + lambda;
            inv_small(D, d, Dinv);
            for (int i = 0; i < d; i++) { double s = 0; for (int j = 0; j < d; j++) s += Dinv[i * d + j] * bl[(size_t)l * d + j]; db[i] = s; }
            // merge edges that hit the same pose block (g2o keeps ONE Hpl block per (pose, landmark))
            int slots[64]; double Wm[64][36]; int ns = 0;
            for (int e = ptr[l]; e < ptr[l + 1]; e++) {
                int s = slot[e]; if (s < 0) continue;
                int k = 0; for (; k < ns; k++) if (slots[k] == s) break;
                if (k == ns) { if (ns >= 64) { fprintf(stderr, "oracle: track too long\n"); abort(); } slots[ns] = s; for (int i = 0; i < 6 * d; i++) Wm[ns][i] = 0; ns++; }
                for (int i = 0; i < 6 * d; i++) Wm[k][i] += W[(size_t)e * 6 * d + i];
            }
            for (int a = 0; a < ns; a++) {
                double Bd[36];
                for (int r = 0; r < 6; r++) for (int c = 0; c < d; c++) { double s = 0; for (int k = 0; k < d; k++) s += Wm[a][r * d + k] * Dinv[k * d + c]; Bd[r * d + c] = s; }
                for (int r = 0; r < 6; r++) { double s = 0; for (int k = 0; k < d; k++) s += Wm[a][r * d + k] * db[k]; Cv[6 * slots[a] + r] += s; }
                for (int b = 0; b < ns; b++) {
                    if (slots[b] < slots[a]) continue;   // upper blocks only (i1 <= i2)
                    for (int r = 0; r < 6; r++) for (int c = 0; c < 6; c++) {
                        if (slots[a] == slots[b] && c > r) continue;
                        double s = 0; for (int k = 0; k < d; k++) s += Bd[r * d + k] * Wm[b][c * d + k];
                        // block (a,b)(r,c) lives in the lower skyline at (6*sb + c, 6*sa + r); diagonal block: (r,c) with c<=r
                        int64_t idx;
                        if (slots[a] == slots[b]) idx = S.rowptr[6 * slots[a] + r] + (6 * slots[a] + c - S.first[6 * slots[a] + r]);
                        else idx = S.rowptr[6 * slots[b] + c] + (6 * slots[a] + r - S.first[6 * slots[b] + c]);
                        Sv[idx] -= s;
                    }
                }
            }
        }
    };
    process(B.n_pt, 3, B.ptr_p, B.slot_p, B.W_p, B.Hll_pt, B.bl_pt);
    process(B.n_ls, B.dl, B.ptr_l, B.slot_l, B.W_l, B.Hll_ls, B.bl_ls);
    for (size_t t = 0; t < Sthr.size(); t++) {
        for (size_t i = 0; i < S.val.size(); i++) S.val[i] += Sthr[t][i];
        for (int i = 0; i < 6 * nf; i++) coeff[i] += Cthr[t][i];
    }
    xp.assign(6 * nf, 0.0);
    for (int i = 0; i < 6 * nf; i++) xp[i] = B.bp[i] - coeff[i];
    if (g_out) *g_out = xp;
    if (S_dense) {
        S_dense->assign((size_t)nf * nf * 36, 0.0);
        for (int i = 0; i < S.n; i++)
            for (int j = S.first[i]; j <= i; j++) {
                int bi = i / 6, bj = j / 6, r = i % 6, c = j % 6;   // lower entry (i,j): block row bi >= block col bj
                double v = S.at(i, j);
                // export as upper block (bj, bi) entry (c, r); diagonal blocks filled symmetrically
                (*S_dense)[((size_t)bj * nf + bi) * 36 + c * 6 + r] = v;
                if (bi == bj) (*S_dense)[((size_t)bj * nf + bi) * 36 + r * 6 + c] = v;
            }
    }
    if (!solve) return true;
    bool ok = skyline_ldlt_solve(S, xp);
    // back-substitution  x_l = Dinv (b_l - W^T x_p)
    auto backsub = [&](int n_lm, int d, const std::vector<int> &ptr, const std::vector<int> &slot, const std::vector<double> &W,
                       const std::vector<double> &Hll, const std::vector<double> &bl, std::vector<double> &xl) {
        xl.assign((size_t)n_lm * d, 0.0);
#pragma omp parallel for schedule(static) num_threads(T) if (T > 1)
        for (int l = 0; l < n_lm; l++) {
            double D[36], Dinv[36], c[6];
            for (int i = 0; i < d * d; i++) D[i] = Hll[(size_t)l * d * d + i];
            for (int i = 0; i < d; i++) D[i * d + i] = multiplicative ? D[i * d + i] + lambda * D[i * d + i] : D[i * d + i] + lambda;
            inv_small(D, d, Dinv);
            for (int i = 0; i < d; i++) c[i] = bl[(size_t)l * d + i];
            for (int e = ptr[l]; e < ptr[l + 1]; e++) {
                int s = slot[e]; if (s < 0) continue;
                for (int k = 0; k < d; k++) { double a = 0; for (int r = 0; r < 6; r++) a += W[(size_t)e * 6 * d + r * d + k] * xp[6 * s + r]; c[k] -= a; }
            }
            for (int i = 0; i < d; i++) { double s = 0; for (int j = 0; j < d; j++) s += Dinv[i * d + j] * c[j]; xl[(size_t)l * d + i] = s; }
        }
    };
    backsub(B.n_pt, 3, B.ptr_p, B.slot_p, B.W_p, B.Hll_pt, B.bl_pt, xl_pt);
    backsub(B.n_ls, B.dl, B.ptr_l, B.slot_l, B.W_l, B.Hll_ls, B.bl_ls, xl_ls);
    return ok;
}

// Literal full-system variant: dense N x N matrix, H_ii damping, LDL^T without pivoting (what profile H does,
// src/mapHandler.cpp:2343,2564-2568).  Used to pin solve_schur on small problems.
static bool solve_dense(const BlockSystem &B, double lambda, bool multiplicative,
                        std::vector<double> &xp, std::vector<double> &xl_pt, std::vector<double> &xl_ls) {
    const int nf = B.n_free, dl = B.dl;
    const int N = 6 * nf + 3 * B.n_pt + dl * B.n_ls;
    std::vector<double> H((size_t)N * N, 0.0), g(N, 0.0);
    for (int s = 0; s < nf; s++) { for (int r = 0; r < 6; r++) { for (int c = 0; c < 6; c++) H[(size_t)(6 * s + r) * N + 6 * s + c] = B.Hpp[s * 36 + r * 6 + c]; g[6 * s + r] = B.bp[6 * s + r]; } }
    auto put = [&](int n_lm, int d, int base, const std::vector<int> &ptr, const std::vector<int> &slot, const std::vector<double> &W,
                   const std::vector<double> &Hll, const std::vector<double> &bl) {
        for (int l = 0; l < n_lm; l++) {
            int j0 = base + d * l;
            for (int r = 0; r < d; r++) { for (int c = 0; c < d; c++) H[(size_t)(j0 + r) * N + j0 + c] = Hll[(size_t)l * d * d + r * d + c]; g[j0 + r] = bl[(size_t)l * d + r]; }
            for (int e = ptr[l]; e < ptr[l + 1]; e++) {
                int s = slot[e]; if (s < 0) continue;
                for (int r = 0; r < 6; r++) for (int c = 0; c < d; c++) { double w = W[(size_t)e * 6 * d + r * d + c]; H[(size_t)(6 * s + r) * N + j0 + c] += w; H[(size_t)(j0 + c) * N + 6 * s + r] += w; }
            }
        }
    };
    put(B.n_pt, 3, 6 * nf, B.ptr_p, B.slot_p, B.W_p, B.Hll_pt, B.bl_pt);
    put(B.n_ls, dl, 6 * nf + 3 * B.n_pt, B.ptr_l, B.slot_l, B.W_l, B.Hll_ls, B.bl_ls);
    for (int i = 0; i < N; i++) H[(size_t)i * N + i] = multiplicative ? H[(size_t)i * N + i] + lambda * H[(size_t)i * N + i] : H[(size_t)i * N + i] + lambda;
    // dense LDL^T on the lower triangle
    std::vector<double> d(N);
    for (int i = 0; i < N; i++) {
        double *ri = &H[(size_t)i * N];
        for (int j = 0; j < i; j++) { const double *rj = &H[(size_t)j * N]; double s = ri[j]; for (int k = 0; k < j; k++) s -= ri[k] * rj[k]; ri[j] = s; }
        double di = ri[i];
        for (int j = 0; j < i; j++) { double y = ri[j], l = y / d[j]; di -= y * l; ri[j] = l; }
        d[i] = di;
        if (di == 0.0 || !std::isfinite(di)) return false;
    }
    for (int i = 0; i < N; i++) { const double *ri = &H[(size_t)i * N]; double s = g[i]; for (int j = 0; j < i; j++) s -= ri[j] * g[j]; g[i] = s; }
    for (int i = 0; i < N; i++) g[i] /= d[i];
    for (int i = N - 1; i >= 0; i--) { const double *ri = &H[(size_t)i * N]; double xi = g[i]; for (int j = 0; j < i; j++) g[j] -= ri[j] * xi; }
    xp.assign(g.begin(), g.begin() + 6 * nf);
    xl_pt.assign(g.begin() + 6 * nf, g.begin() + 6 * nf + 3 * B.n_pt);
    xl_ls.assign(g.begin() + 6 * nf + 3 * B.n_pt, g.end());
    return true;
}

static M4 T_from_rows(const double *r12) {
    M4 T = M4::Identity();
    for (int r = 0; r < 3; r++) for (int c = 0; c < 4; c++) T(r, c) = r12[r * 4 + c];
    return T;
}
static void rows_from_T(const M4 &T, double *r12) { for (int r = 0; r < 3; r++) for (int c = 0; c < 4; c++) r12[r * 4 + c] = T(r, c); }

static void init_system(const plba_problem &P, int dl, BlockSystem &B) {
    B.n_free = P.n_free; B.n_pt = P.n_pt; B.n_ls = P.n_ls; B.dl = dl;
    B.Hpp.assign((size_t)P.n_free * 36, 0.0); B.bp.assign((size_t)P.n_free * 6, 0.0);
    B.Hll_pt.assign((size_t)P.n_pt * 9, 0.0); B.bl_pt.assign((size_t)P.n_pt * 3, 0.0);
    B.Hll_ls.assign((size_t)P.n_ls * dl * dl, 0.0); B.bl_ls.assign((size_t)P.n_ls * dl, 0.0);
    B.slot_p.assign(P.n_pobs, -1); B.slot_l.assign(P.n_lobs, -1);
    B.W_p.assign((size_t)P.n_pobs * 18, 0.0); B.W_l.assign((size_t)P.n_lobs * 6 * dl, 0.0);
    B.ptr_p = csr_from_sorted(P.po_lm, P.n_pobs, P.n_pt);
    B.ptr_l = csr_from_sorted(P.lo_lm, P.n_lobs, P.n_ls);
}

static int validate(const plba_problem &P) {
    if (P.n_kf < 0 || P.n_free < 0 || P.n_pt < 0 || P.n_ls < 0 || P.n_pobs < 0 || P.n_lobs < 0) return PLBA_E_ARG;
    for (int i = 0; i < P.n_pobs; i++) { if (P.po_lm[i] < 0 || P.po_lm[i] >= P.n_pt || P.po_kf[i] < 0 || P.po_kf[i] >= P.n_kf) return PLBA_E_ARG; if (i && P.po_lm[i] < P.po_lm[i - 1]) return PLBA_E_ARG; }
    for (int i = 0; i < P.n_lobs; i++) { if (P.lo_lm[i] < 0 || P.lo_lm[i] >= P.n_ls || P.lo_kf[i] < 0 || P.lo_kf[i] >= P.n_kf) return PLBA_E_ARG; if (i && P.lo_lm[i] < P.lo_lm[i - 1]) return PLBA_E_ARG; }
    for (int i = 0; i < P.n_kf; i++) if (P.kf_slot[i] < -1 || P.kf_slot[i] >= P.n_free) return PLBA_E_ARG;
    return PLBA_OK;
}

static void push_trace(plba_result *res, const plba_trace_rec &r) {
    if (res->trace && res->n_trace < res->trace_cap) res->trace[res->n_trace] = r;
    res->n_trace++;
}

// =====================================================================================================
// Profile G   
// =====================================================================================================
struct GRun {
    const plba_problem &P; const plba_options &O; plba_result *res; int window;
    std::vector<M4> Tcw;            // all KFs: estimate = T_kf_w^-1 (src/mapHandler.cpp:5940,5960)
    std::vector<V3> pts;            // :5982
    std::vector<V4> orth;           // :6040-6046
    std::vector<V2> e_p, e_l;       // cached _error per edge
    std::vector<uint8_t> lvl_p, lvl_l;
    std::vector<double> om_p, om_l; // information = (float)(1/sigma2)  (:6009-6010, Q13)
    BlockSystem B; Skyline S;
    bool dense_solve = false;
    GRun(const plba_problem &p, const plba_options &o, plba_result *r, int w) : P(p), O(o), res(r), window(w) {}

    V2 pobs(int i) const { V2 v; v[0] = P.po_uv[2 * i]; v[1] = P.po_uv[2 * i + 1]; return v; }
    V4 lobs(int i) const { V4 v; for (int k = 0; k < 4; k++) v[k] = P.lo_ab[4 * i + k]; return v; }

    void computeActiveErrors() {
        const int T = nthreads();
#pragma omp parallel for schedule(static) num_threads(T) if (T > 1)
        for (int i = 0; i < P.n_pobs; i++) if (!lvl_p[i]) e_p[i] = pointEdgeError(P.cam, Tcw[P.po_kf[i]], pts[P.po_lm[i]], pobs(i));
#pragma omp parallel for schedule(static) num_threads(T) if (T > 1)
        for (int i = 0; i < P.n_lobs; i++) if (!lvl_l[i]) e_l[i] = lineEdgeError(P.cam, Tcw[P.lo_kf[i]], orth[P.lo_lm[i]], lobs(i));
    }
    double chi2_p(int i) const { return om_p[i] * e_p[i].squaredNorm(); }
    double chi2_l(int i) const { return om_l[i] * e_l[i].squaredNorm(); }
    double activeRobustChi2(bool kernels) const {
        double chi = 0; double rho[3];
        for (int i = 0; i < P.n_pobs; i++) if (!lvl_p[i]) { double c = chi2_p(i); if (kernels) { huberRobustify(O.huber_delta, c, rho); c = rho[0]; } chi += c; }
        for (int i = 0; i < P.n_lobs; i++) if (!lvl_l[i]) { double c = chi2_l(i); if (kernels) { huberRobustify(O.huber_delta, c, rho); c = rho[0]; } chi += c; }
        return chi;
    }
    // BlockSolver::buildSystem: linearizeOplus + constructQuadraticForm per active edge
    void buildSystem(bool kernels) {
        B.zero();
        const int T = nthreads();
        std::vector<std::vector<double>> Hthr(T), bthr(T);
        for (int t = 0; t < T; t++) { Hthr[t].assign(B.Hpp.size(), 0.0); bthr[t].assign(B.bp.size(), 0.0); }
        const bool q12 = (O.quirks == PLBA_QUIRKS_FAITHFUL);
#pragma omp parallel for schedule(static) num_threads(T) if (T > 1)
        for (int l = 0; l < P.n_pt; l++) {
            int tid = 0;
#ifdef _OPENMP
            tid = omp_get_thread_num();
#endif
            for (int i = B.ptr_p[l]; i < B.ptr_p[l + 1]; i++) {
                if (lvl_p[i]) continue;
                PointEdgeLin L; pointEdgeLinearize(P.cam, Tcw[P.po_kf[i]], pts[l], L);
                double rho[3] = {0, 1, 0};
                if (kernels) huberRobustify(O.huber_delta, chi2_p(i), rho);
                double wom = rho[1] * om_p[i];                       // robustInformation(rho) = rho[1] * information
                V2 omega_r = e_p[i] * (-om_p[i]) * rho[1];          // omega_r = -omega*error ; omega_r *= rho[1]
                Mat<3, 1> bl = L.Jxi.T() * omega_r; Mat<3, 3> Hl = L.Jxi.T() * L.Jxi * wom;
                for (int k = 0; k < 3; k++) B.bl_pt[3 * l + k] += bl[k];
                for (int k = 0; k < 9; k++) B.Hll_pt[9 * l + k] += Hl[k];
                int s = P.kf_slot[P.po_kf[i]];
                if (s >= 0) {
                    Mat<6, 1> bpv = L.Jxj.T() * omega_r; Mat<6, 6> Hp = L.Jxj.T() * L.Jxj * wom; Mat<6, 3> W = L.Jxj.T() * L.Jxi * wom;
                    for (int k = 0; k < 6; k++) bthr[tid][6 * s + k] += bpv[k];
                    for (int k = 0; k < 36; k++) Hthr[tid][36 * s + k] += Hp[k];
                    for (int k = 0; k < 18; k++) B.W_p[(size_t)i * 18 + k] = W[k];
                    B.slot_p[i] = s;
                }
            }
        }
#pragma omp parallel for schedule(static) num_threads(T) if (T > 1)
        for (int l = 0; l < P.n_ls; l++) {
            int tid = 0;
#ifdef _OPENMP
            tid = omp_get_thread_num();
#endif
            for (int i = B.ptr_l[l]; i < B.ptr_l[l + 1]; i++) {
                if (lvl_l[i]) continue;
                LineEdgeLin L; lineEdgeLinearize(P.cam, Tcw[P.lo_kf[i]], orth[l], lobs(i), q12, L);
                double rho[3] = {0, 1, 0};
                if (kernels) huberRobustify(O.huber_delta, chi2_l(i), rho);
                double wom = rho[1] * om_l[i];
                V2 omega_r = e_l[i] * (-om_l[i]) * rho[1];
                Mat<4, 1> bl = L.Jxi.T() * omega_r; Mat<4, 4> Hl = L.Jxi.T() * L.Jxi * wom;
                for (int k = 0; k < 4; k++) B.bl_ls[4 * l + k] += bl[k];
                for (int k = 0; k < 16; k++) B.Hll_ls[16 * l + k] += Hl[k];
                int s = P.kf_slot[P.lo_kf[i]];
                if (s >= 0) {
                    Mat<6, 1> bpv = L.Jxj.T() * omega_r; Mat<6, 6> Hp = L.Jxj.T() * L.Jxj * wom; Mat<6, 4> W = L.Jxj.T() * L.Jxi * wom;
                    for (int k = 0; k < 6; k++) bthr[tid][6 * s + k] += bpv[k];
                    for (int k = 0; k < 36; k++) Hthr[tid][36 * s + k] += Hp[k];
                    for (int k = 0; k < 24; k++) B.W_l[(size_t)i * 24 + k] = W[k];
                    B.slot_l[i] = s;
                }
            }
        }
        for (int t = 0; t < T; t++) { for (size_t k = 0; k < B.Hpp.size(); k++) B.Hpp[k] += Hthr[t][k]; for (size_t k = 0; k < B.bp.size(); k++) B.bp[k] += bthr[t][k]; }
    }
    double maxDiagonal() const {   // OptimizationAlgorithmLevenberg::computeLambdaInit
        double m = 0;
        for (int s = 0; s < P.n_free; s++) for (int r = 0; r < 6; r++) m = std::max(m, std::fabs(B.Hpp[36 * s + 7 * r]));
        for (int l = 0; l < P.n_pt; l++) for (int r = 0; r < 3; r++) m = std::max(m, std::fabs(B.Hll_pt[9 * l + 4 * r]));
        for (int l = 0; l < P.n_ls; l++) for (int r = 0; r < 4; r++) m = std::max(m, std::fabs(B.Hll_ls[16 * l + 5 * r]));
        return m;
    }
    // returns number of outer iterations done
    int stage(int stage_id, bool kernels, int n_outer) {
        double lambda = 0, ni = 2;
        int it = 0;
        for (; it < n_outer; it++) {
            computeActiveErrors();
            double currentChi = activeRobustChi2(kernels), tempChi = currentChi;
            buildSystem(kernels);
            if (it == 0) { lambda = O.lm_tau * maxDiagonal(); ni = 2; }
            double rho = 0; int qmax = 0;
            do {
                std::vector<M4> Tb = Tcw; std::vector<V3> pb = pts; std::vector<V4> ob = orth;   // push()
                std::vector<double> xp, xlp, xll;
                bool ok2 = dense_solve ? solve_dense(B, lambda, false, xp, xlp, xll) : solve_schur(P, B, S, lambda, false, xp, xlp, xll);
                res->n_trials++;
                // update: oplus on every free vertex
                for (int k = 0; k < P.n_kf; k++) { int s = P.kf_slot[k]; if (s < 0) continue; V6 d; for (int i = 0; i < 6; i++) d[i] = xp[6 * s + i]; Tcw[k] = poseOplusG2O(Tcw[k], d); }
                for (int l = 0; l < P.n_pt; l++) for (int i = 0; i < 3; i++) pts[l][i] += xlp[3 * l + i];
                for (int l = 0; l < P.n_ls; l++) { V4 d; for (int i = 0; i < 4; i++) d[i] = xll[4 * l + i]; orth[l] = updateOrthCoord(orth[l], d); }
                computeActiveErrors();
                tempChi = activeRobustChi2(kernels);
                if (!ok2) tempChi = std::numeric_limits<double>::max();
                rho = currentChi - tempChi;
                double scale = 0;   // computeScale(): sum_j x_j (lambda x_j + b_j)
                for (int j = 0; j < 6 * P.n_free; j++) scale += xp[j] * (lambda * xp[j] + B.bp[j]);
                for (size_t j = 0; j < xlp.size(); j++) scale += xlp[j] * (lambda * xlp[j] + B.bl_pt[j]);
                for (size_t j = 0; j < xll.size(); j++) scale += xll[j] * (lambda * xll[j] + B.bl_ls[j]);
                scale += 1e-3;
                rho /= scale;
                plba_trace_rec tr{}; tr.window = window; tr.stage = stage_id; tr.iter = it; tr.trial = qmax; tr.chi = currentChi; tr.chi_new = tempChi;
                tr.rho = rho; tr.lambda = lambda; tr.scale = scale;
                if (rho > 0 && std::isfinite(tempChi)) {
                    double alpha = 1. - std::pow((2 * rho - 1), 3);
                    alpha = std::min(alpha, 2. / 3.);
                    double scaleFactor = std::max(1. / 3., alpha);
                    lambda *= scaleFactor; ni = 2; currentChi = tempChi;
                    tr.accepted = 1;
                } else {
                    lambda *= ni; ni *= 2;
                    Tcw = Tb; pts = pb; orth = ob;   // pop()
                    tr.accepted = 0;
                }
                qmax++;
                bool again = (rho < 0 && qmax < O.lm_max_trials);
                if (!again && (qmax == O.lm_max_trials || rho == 0)) tr.stop = 1;
                push_trace(res, tr);
                if (!again) break;
            } while (true);
            if (qmax == O.lm_max_trials || rho == 0) { it++; break; }   // Terminate
        }
        return it;
    }

    int run() {
        const int nk = P.n_kf;
        Tcw.resize(nk);
        for (int k = 0; k < nk; k++) Tcw[k] = inverse_se3(T_from_rows(P.kf_T_wc + 12 * k));   // Q16: reference uses the general 4x4 inverse
        pts.resize(P.n_pt); for (int l = 0; l < P.n_pt; l++) for (int i = 0; i < 3; i++) pts[l][i] = P.pt_xyz[3 * l + i];
        orth.resize(P.n_ls); for (int l = 0; l < P.n_ls; l++) { V6 pl; for (int i = 0; i < 6; i++) pl[i] = P.ls_plk[6 * l + i]; orth[l] = changePlukerToOrth(pl); }
        e_p.resize(P.n_pobs); e_l.resize(P.n_lobs); lvl_p.assign(P.n_pobs, 0); lvl_l.assign(P.n_lobs, 0);
        om_p.resize(P.n_pobs); om_l.resize(P.n_lobs);
        for (int i = 0; i < P.n_pobs; i++) { float f = (float)(1.0 / (P.po_sig2 ? P.po_sig2[i] : 1.0)); om_p[i] = (double)f; }
        for (int i = 0; i < P.n_lobs; i++) { float f = (float)(1.0 / (P.lo_sig2 ? P.lo_sig2[i] : 1.0)); om_l[i] = (double)f; }
        init_system(P, 4, B);
        build_skyline(P, B.ptr_p, B.ptr_l, S);

        stage(0, true, O.iters_stage1);                                     // :6121-6122
        for (int i = 0; i < P.n_pobs; i++) {                                // :6125-6134
            bool depth_pos = pointPc(Tcw[P.po_kf[i]], pts[P.po_lm[i]])[2] > 0.0;
            if (chi2_p(i) > O.chi2_gate || !depth_pos) lvl_p[i] = 1;
        }
        for (int i = 0; i < P.n_lobs; i++) if (chi2_l(i) > O.chi2_gate) lvl_l[i] = 1;   // :6138-6147
        stage(1, false, O.iters_stage2);                                    // :6151-6152 (all kernels removed, Q15)

        for (int i = P.n_pobs - 1; i >= 0; i--) {                           // :6156-6217
            if (lvl_p[i]) e_p[i] = pointEdgeError(P.cam, Tcw[P.po_kf[i]], pts[P.po_lm[i]], pobs(i));
            bool depth_pos = pointPc(Tcw[P.po_kf[i]], pts[P.po_lm[i]])[2] > 0.0;
            double c = chi2_p(i);
            uint8_t f = lvl_p[i] ? PLBA_OBS_LEVEL1 : 0;
            if (c > O.chi2_gate || !depth_pos) f |= PLBA_OBS_BAD;
            if (!depth_pos) f |= PLBA_OBS_NEGDEPTH;
            if (res->po_flags) res->po_flags[i] = f;
            if (res->po_chi2) res->po_chi2[i] = c;
        }
        for (int i = P.n_lobs - 1; i >= 0; i--) {                           // :6224-6291
            if (lvl_l[i]) e_l[i] = lineEdgeError(P.cam, Tcw[P.lo_kf[i]], orth[P.lo_lm[i]], lobs(i));
            double c = chi2_l(i);
            uint8_t f = lvl_l[i] ? PLBA_OBS_LEVEL1 : 0;
            if (c > O.chi2_gate) f |= PLBA_OBS_BAD;
            if (res->lo_flags) res->lo_flags[i] = f;
            if (res->lo_chi2) res->lo_chi2[i] = c;
        }
        // write-back :6297-6319
        if (res->kf_T_wc) for (int k = 0; k < nk; k++) {
            if (P.kf_slot[k] >= 0) rows_from_T(inverse_se3(Tcw[k]), res->kf_T_wc + 12 * k);
            else for (int i = 0; i < 12; i++) res->kf_T_wc[12 * k + i] = P.kf_T_wc[12 * k + i];
        }
        if (res->pt_xyz) for (int l = 0; l < P.n_pt; l++) for (int i = 0; i < 3; i++) res->pt_xyz[3 * l + i] = pts[l][i];
        for (int l = 0; l < P.n_ls; l++) {
            if (res->ls_orth) for (int i = 0; i < 4; i++) res->ls_orth[4 * l + i] = orth[l][i];
            if (res->ls_plk) { V6 pl = changeOrthToPluker(orth[l]); for (int i = 0; i < 6; i++) res->ls_plk[6 * l + i] = pl[i]; }
        }
        if (res->pt_inlier) for (int l = 0; l < P.n_pt; l++) res->pt_inlier[l] = 1;
        if (res->ls_inlier) for (int l = 0; l < P.n_ls; l++) res->ls_inlier[l] = 1;
        return PLBA_OK;
    }
};

// =====================================================================================================
// Profile H   
// =====================================================================================================
struct HTerm { double J_p[6]; double J_l[6]; double r; double w; };

// Point term, src/mapHandler.cpp:2368-2411 (pass 0) == :2610-2658 (loop).  Tiw = inverse_se3(T_kf_w).
static void h_point_term(const double cam[4], const M4 &Tiw, const V3 &Xwj, const V2 &p_obs, double homog_th, HTerm &t) {
    M3 R = Tiw.block<3, 3>(0, 0);
    V3 Xwi = R * Xwj + Tiw.block<3, 1>(0, 3);
    V2 p_prj = projection(cam, Xwi);
    V2 p_err = p_obs - p_prj;
    double p_err_norm = p_err.norm();
    double gx = Xwi[0], gy = Xwi[1], gz = Xwi[2];
    double gz2 = gz * gz;
    gz2 = 1.0 / std::max(homog_th, gz2);
    double fx = cam[0], fy = cam[1];
    double dx = p_err[0], dy = p_err[1];
    double fxdx = fx * dx, fydy = fy * dy;
    V6 J;
    J[0] = +gz2 * fxdx * gz;
    J[1] = +gz2 * fydy * gz;
    J[2] = -gz2 * (fxdx * gx + fydy * gy);
    J[3] = -gz2 * (fxdx * gx * gy + fydy * gy * gy + fydy * gz * gz);
    J[4] = +gz2 * (fxdx * gx * gx + fxdx * gz * gz + fydy * gx * gy);
    J[5] = +gz2 * (fydy * gx * gz - fxdx * gy * gz);
    J = J / std::max(homog_th, p_err_norm);
    Mat<1, 3> Jl0;
    Jl0(0, 0) = +gz2 * fxdx * gz; Jl0(0, 1) = +gz2 * fydy * gz; Jl0(0, 2) = -gz2 * (fxdx * gx + fydy * gy);
    Mat<1, 3> Jl = (Jl0 * R) / std::max(homog_th, p_err_norm);
    for (int i = 0; i < 6; i++) t.J_p[i] = J[i];
    for (int i = 0; i < 3; i++) t.J_l[i] = Jl(0, i);
    t.r = p_err_norm; t.w = robustWeightCauchy(p_err_norm);
}

// Endpoint-line term, src/mapHandler.cpp:2450-2524 (pass 0) / :2694-2767 (loop).  th = homogTh() or the literal 1e-7.
// q5_fixed: use the observed line coefficients (lx,ly) as upstream pl-slam does instead of (e0,e1) (Q5).
static void h_endline_term(const double cam[4], const M4 &Tiw, const V3 &Pwj, const V3 &Qwj, const double *l_obs, double th, bool q5_fixed, HTerm &t) {
    M3 R = Tiw.block<3, 3>(0, 0);
    V3 tt = Tiw.block<3, 1>(0, 3);
    V3 Pwi = R * Pwj + tt, Qwi = R * Qwj + tt;
    V2 p_prj = projection(cam, Pwi), q_prj = projection(cam, Qwi);
    V2 l_err;
    l_err[0] = l_obs[0] * p_prj[0] + l_obs[1] * p_prj[1] + l_obs[2];
    l_err[1] = l_obs[0] * q_prj[0] + l_obs[1] * q_prj[1] + l_obs[2];
    double l_err_norm = l_err.norm();
    double fx = cam[0], fy = cam[1];
    double lx = q5_fixed ? l_obs[0] : l_err[0];
    double ly = q5_fixed ? l_obs[1] : l_err[1];
    double fxlx = fx * lx, fyly = fy * ly;
    auto endpoint = [&](const V3 &G, V6 &Jp, Mat<1, 3> &Jl0) {
        double gx = G[0], gy = G[1], gz = G[2];
        double gz2 = gz * gz;
        gz2 = 1.0 / std::max(th, gz2);
        Jp[0] = +gz2 * fxlx * gz;
        Jp[1] = +gz2 * fyly * gz;
        Jp[2] = -gz2 * (fxlx * gx + fyly * gy);
        Jp[3] = -gz2 * (fxlx * gx * gy + fyly * gy * gy + fyly * gz * gz);
        Jp[4] = +gz2 * (fxlx * gx * gx + fxlx * gz * gz + fyly * gx * gy);
        Jp[5] = +gz2 * (fyly * gx * gz - fxlx * gy * gz);
        Jl0(0, 0) = +gz2 * fxlx * gz; Jl0(0, 1) = +gz2 * fyly * gz; Jl0(0, 2) = -gz2 * (fxlx * gx + fyly * gy);
    };
    V6 JP, JQ; Mat<1, 3> JlP0, JlQ0;
    endpoint(Pwi, JP, JlP0);
    Mat<1, 3> JlP = (JlP0 * R) * l_err[0] / std::max(th, l_err_norm);
    endpoint(Qwi, JQ, JlQ0);
    Mat<1, 3> JlQ = (JlQ0 * R) * l_err[1] / std::max(th, l_err_norm);
    V6 J = (JP * l_err[0] + JQ * l_err[1]) / std::max(th, l_err_norm);
    for (int i = 0; i < 6; i++) t.J_p[i] = J[i];
    for (int i = 0; i < 3; i++) { t.J_l[i] = JlP(0, i); t.J_l[3 + i] = JlQ(0, i); }
    t.r = l_err_norm; t.w = robustWeightCauchy(l_err_norm);
}

// Plücker-line term of H_PLK, src/mapHandler.cpp:1741-1812 (pass 0) / :2001-2075 (loop).
// faithful: Q6 (fenmu used where 1/fenmu is meant), Q7 (sign of d n/d theta3), Q8 (line Jacobian sign).
static void h_plkline_term(const double cam[4], const M4 &Tiw, const V6 &NDw, const double *l_obs, double th, bool fixed, HTerm &t) {
    M3 Rw = getOrhtRFromPluker(NDw);
    M2 Ww = getOrthWFromPluker(NDw);
    Mat<6, 4> jacobianPO = jacobianFromPlukerToOrth(Rw, Ww, fixed ? +1.0 : -1.0);
    V6 NDc = getTransformMatrixForPluker(Tiw) * NDw;
    V3 l = plukerK(cam) * NDc.block<3, 1>(0, 0);
    double fenmu = std::sqrt(l[0] * l[0] + l[1] * l[1]);
    V2 l_err;
    l_err[0] = (l_obs[0] * l[0] + l_obs[1] * l[1] + l[2]) / fenmu;
    l_err[1] = (l_obs[2] * l[0] + l_obs[3] * l[1] + l[2]) / fenmu;
    double l_err_norm = l_err.norm();
    double a0 = l_obs[0], b0 = l_obs[1], a1 = l_obs[2], b1 = l_obs[3];
    double lx = l[0], ly = l[1];
    Mat<1, 3> f0, f1;
    if (!fixed) {
        f0(0, 0) = a0 * fenmu - lx * l_err[0] * fenmu * fenmu; f0(0, 1) = b0 * fenmu - ly * l_err[0] * fenmu * fenmu; f0(0, 2) = fenmu;
        f1(0, 0) = a1 * fenmu - lx * l_err[1] * fenmu * fenmu; f1(0, 1) = b1 * fenmu - ly * l_err[1] * fenmu * fenmu; f1(0, 2) = fenmu;
    } else {
        f0(0, 0) = -lx * l_err[0] / (fenmu * fenmu) + a0 / fenmu; f0(0, 1) = -ly * l_err[0] / (fenmu * fenmu) + b0 / fenmu; f0(0, 2) = 1.0 / fenmu;
        f1(0, 0) = -lx * l_err[1] / (fenmu * fenmu) + a1 / fenmu; f1(0, 1) = -ly * l_err[1] / (fenmu * fenmu) + b1 / fenmu; f1(0, 2) = 1.0 / fenmu;
    }
    Mat<3, 6> KL; KL.setBlock<3, 3>(0, 0, plukerK(cam));
    M3 DR = Tiw.block<3, 3>(0, 0); V3 Dt = Tiw.block<3, 1>(0, 3);
    V3 nw = NDw.block<3, 1>(0, 0), dw = NDw.block<3, 1>(3, 0);
    M6 RT;
    RT.setBlock<3, 3>(0, 3, -vechat(DR * nw) - vechat(Dt) * vechat(DR * dw));
    RT.setBlock<3, 3>(0, 0, -vechat(DR * dw));
    Mat<1, 6> jac0 = f0 * KL * RT, jac1 = f1 * KL * RT;
    Mat<1, 6> Jp = (jac0 * l_err[0] + jac1 * l_err[1]) / std::max(th, l_err_norm);
    M6 TM = getTransformMatrixForPluker(Tiw);
    Mat<1, 4> jl0 = f0 * KL * TM * jacobianPO, jl1 = f1 * KL * TM * jacobianPO;
    Mat<1, 4> Jl = (jl0 * l_err[0] + jl1 * l_err[1]) / std::max(th, l_err_norm);
    // Q8: point terms carry -dr/dx with X += H^-1 g; the reference leaves the line rows at +dr/dx.
    double sgn = fixed ? -1.0 : 1.0;
    for (int i = 0; i < 6; i++) t.J_p[i] = sgn * Jp(0, i);
    for (int i = 0; i < 4; i++) t.J_l[i] = sgn * Jl(0, i);
    t.r = l_err_norm; t.w = robustWeightCauchy(l_err_norm);
}

struct HRun {
    const plba_problem &P; const plba_options &O; plba_result *res; int window;
    bool plk;                       // H_PLK (4-dof lines) or H_END (6-dof)
    bool dense_solve = false;
    int dl;
    std::vector<V6> Xkf; std::vector<V3> Xpt; std::vector<double> Xls;   // X vector pieces
    BlockSystem B; Skyline S;
    double point_error = 0, line_error = 0, err = 0;
    HRun(const plba_problem &p, const plba_options &o, plba_result *r, int w) : P(p), O(o), res(r), window(w) { plk = (o.profile == PLBA_PROFILE_H_PLK); dl = plk ? 4 : 6; }

    void accumulate(int d, const HTerm &t, int s, double *Hll, double *bl, double *W, int *slot_out) {
        for (int r = 0; r < d; r++) { bl[r] += t.J_l[r] * t.r * t.w; for (int c = 0; c < d; c++) Hll[r * d + c] += t.J_l[r] * t.J_l[c] * t.w; }
        if (s >= 0) {
            for (int r = 0; r < 6; r++) {
                B.bp[6 * s + r] += t.J_p[r] * t.r * t.w;
                for (int c = 0; c < 6; c++) B.Hpp[36 * s + r * 6 + c] += t.J_p[r] * t.J_p[c] * t.w;
                for (int c = 0; c < d; c++) W[r * d + c] = t.J_l[c] * t.J_p[r] * t.w;     // Haux^T
            }
            *slot_out = s;
        }
    }
    // one linearisation pass.  pass0: everything from MAP values (src/mapHandler.cpp:2361-2550); else from X (:2603-2793)
    void linearize(bool pass0) {
        const bool fixed = (O.quirks == PLBA_QUIRKS_FIXED);
        B.zero(); err = 0; point_error = 0; line_error = 0;
        std::vector<M4> Tloop(P.n_kf), Tmap(P.n_kf);
        for (int k = 0; k < P.n_kf; k++) {
            M4 Tm = T_from_rows(P.kf_T_wc + 12 * k);
            Tmap[k] = inverse_se3(Tm);
            int s = P.kf_slot[k];
            Tloop[k] = (s >= 0) ? inverse_se3(expmap_se3(Xkf[s])) : Tmap[k];
        }
        for (int i = 0; i < P.n_pobs; i++) {
            int l = P.po_lm[i], k = P.po_kf[i], s = P.kf_slot[k];
            V3 Xw; if (pass0) for (int a = 0; a < 3; a++) Xw[a] = P.pt_xyz[3 * l + a]; else Xw = Xpt[l];
            V2 ob; ob[0] = P.po_uv[2 * i]; ob[1] = P.po_uv[2 * i + 1];
            HTerm t; h_point_term(P.cam, pass0 ? Tmap[k] : Tloop[k], Xw, ob, O.homog_th, t);
            accumulate(3, t, s, &B.Hll_pt[9 * l], &B.bl_pt[3 * l], &B.W_p[(size_t)i * 18], &B.slot_p[i]);
            err += t.r * t.r * t.w; point_error += t.r * t.r * t.w;
        }
        for (int i = 0; i < P.n_lobs; i++) {
            int l = P.lo_lm[i], k = P.lo_kf[i], s = P.kf_slot[k];
            // Q4: inside the loop the line terms keep using the MAP pose (:2700, :2010)
            const M4 &Tiw = (pass0 || !fixed) ? Tmap[k] : Tloop[k];
            double th = (pass0 || fixed) ? O.homog_th : 0.0000001;   // loop uses the literal (:2717-2760)
            HTerm t;
            if (!plk) {
                V3 Pw, Qw;
                if (pass0) { for (int a = 0; a < 3; a++) { Pw[a] = P.ls_end[6 * l + a]; Qw[a] = P.ls_end[6 * l + 3 + a]; } }
                else if (!fixed) { for (int a = 0; a < 3; a++) { Pw[a] = Xls[3 * l + a]; Qw[a] = Xls[3 * l + a]; } }   // Q3 (:2697-2698)
                else { for (int a = 0; a < 3; a++) { Pw[a] = Xls[6 * l + a]; Qw[a] = Xls[6 * l + 3 + a]; } }
                h_endline_term(P.cam, Tiw, Pw, Qw, P.lo_ab + 4 * i, th, fixed, t);
            } else {
                V6 NDw;
                if (pass0) for (int a = 0; a < 6; a++) NDw[a] = P.ls_plk[6 * l + a];
                else { V4 o; for (int a = 0; a < 4; a++) o[a] = Xls[4 * l + a]; NDw = changeOrthToPluker(o); }
                h_plkline_term(P.cam, Tiw, NDw, P.lo_ab + 4 * i, O.homog_th, fixed, t);
            }
            accumulate(dl, t, s, &B.Hll_ls[(size_t)dl * dl * l], &B.bl_ls[(size_t)dl * l], &B.W_l[(size_t)i * 6 * dl], &B.slot_l[i]);
            err += t.r * t.r * t.w; line_error += t.r * t.r * t.w;
        }
    }
    double Hmax() const {   // :2555-2560
        double m = 0.0;
        auto upd = [&](double h) { if (h > m || h < -m) m = std::fabs(h); };
        for (int s = 0; s < P.n_free; s++) for (int r = 0; r < 6; r++) upd(B.Hpp[36 * s + 7 * r]);
        for (int l = 0; l < P.n_pt; l++) for (int r = 0; r < 3; r++) upd(B.Hll_pt[9 * l + 4 * r]);
        for (int l = 0; l < P.n_ls; l++) for (int r = 0; r < dl; r++) upd(B.Hll_ls[(size_t)dl * dl * l + (dl + 1) * r]);
        return m;
    }
    bool solve(double lambda, std::vector<double> &xp, std::vector<double> &xlp, std::vector<double> &xll) {
        res->n_trials++;
        return dense_solve ? solve_dense(B, lambda, true, xp, xlp, xll) : solve_schur(P, B, S, lambda, true, xp, xlp, xll);
    }
    void retract(const std::vector<double> &xp, const std::vector<double> &xlp, const std::vector<double> &xll) {
        for (int s = 0; s < P.n_free; s++) {   // :2571-2578
            V6 d; for (int i = 0; i < 6; i++) d[i] = xp[6 * s + i];
            M4 Tprev = expmap_se3(Xkf[s]);
            M4 Tcurr = Tprev * inverse_se3(expmap_se3(d));
            Xkf[s] = logmap_se3(Tcurr);
        }
        for (int l = 0; l < P.n_pt; l++) for (int i = 0; i < 3; i++) Xpt[l][i] += xlp[3 * l + i];
        if (!plk) for (size_t i = 0; i < Xls.size(); i++) Xls[i] += xll[i];
        else for (int l = 0; l < P.n_ls; l++) {   // :1880-1891
            V4 D, dD; for (int i = 0; i < 4; i++) { D[i] = Xls[4 * l + i]; dD[i] = xll[4 * l + i]; }
            V4 p = updateOrthCoord(D, dD);
            for (int i = 0; i < 4; i++) Xls[4 * l + i] = p[i];
        }
    }
    static double vnorm(const std::vector<double> &a, const std::vector<double> &b, const std::vector<double> &c) {
        double s = 0; for (double v : a) s += v * v; for (double v : b) s += v * v; for (double v : c) s += v * v; return std::sqrt(s);
    }
    int run() {
        const bool fixed = (O.quirks == PLBA_QUIRKS_FIXED);
        Xkf.resize(P.n_free);
        for (int k = 0; k < P.n_kf; k++) {
            int s = P.kf_slot[k]; if (s < 0) continue;
            if (P.x_pose) for (int i = 0; i < 6; i++) Xkf[s][i] = P.x_pose[6 * s + i];
            else Xkf[s] = logmap_se3(T_from_rows(P.kf_T_wc + 12 * k));
        }
        Xpt.resize(P.n_pt); for (int l = 0; l < P.n_pt; l++) for (int i = 0; i < 3; i++) Xpt[l][i] = P.pt_xyz[3 * l + i];
        Xls.resize((size_t)dl * P.n_ls);
        std::vector<V4> orth0(P.n_ls);
        for (int l = 0; l < P.n_ls; l++) {
            if (!plk) for (int i = 0; i < 6; i++) Xls[6 * l + i] = P.ls_end[6 * l + i];
            else { V6 pl; for (int i = 0; i < 6; i++) pl[i] = P.ls_plk[6 * l + i]; orth0[l] = changePlukerToOrth(pl); for (int i = 0; i < 4; i++) Xls[4 * l + i] = orth0[l][i]; }   // :1577
        }
        init_system(P, dl, B);
        build_skyline(P, B.ptr_p, B.ptr_l, S);
        double err_prev = 999999999.9;
        double lambda = O.lambda_lba_lm, lambda_k = O.lambda_lba_k;
        const int max_iters = O.max_iters_lba;
        const int Npt = P.n_pt, Nls = P.n_ls;
        std::vector<double> xp, xlp, xll;
        // Global BA shell (levMarquardtOptimizationGBA, src/mapHandler.cpp:3128-3728): same terms as H_END, different shell
        const bool gba = (O.shell == PLBA_SHELL_GBA);
        const double eps = std::numeric_limits<double>::epsilon();
        const double th_err_change = gba ? eps : O.min_error_change, th_err = gba ? eps : O.min_error, th_dx = gba ? eps : O.min_error_change;   // :3664, :3694
        const double norm_fixed = gba ? (double)(P.n_pobs + P.n_lobs) : (double)(Npt + Nls);

        linearize(true);
        // Q1: err /= (Npt_obs + Nls_obs) with both counters left at 0  (:2551)
        if (!fixed) err /= (double)0; else err /= norm_fixed;     // (GBA: `err` is not even initialised, :3142; restated as 0 + sum)
        {
            double hm = Hmax();                                   // :2555-2561
            if (gba && !fixed) hm = (double)(long long)hm;        // `int Hmax` (:3386-3391): the scale is truncated to an integer
            lambda *= hm;
        }
        solve(lambda, xp, xlp, xll);                              // :2564-2568
        retract(xp, xlp, xll);                                    // :2571-2586
        {
            plba_trace_rec tr{}; tr.window = window; tr.iter = 0; tr.accepted = 1; tr.chi = err; tr.lambda = lambda;
            tr.dx_norm = vnorm(xp, xlp, xll); tr.err_pt = point_error; tr.err_ls = line_error; push_trace(res, tr);
        }
        err_prev = err;
        for (int iters = 1; iters < max_iters; iters++) {         // :2594
            linearize(false);
            // :2796 ; H_PLK divides by zero in every iteration (:2108, Q1)
            if ((plk || gba) && !fixed) err /= (double)0; else err /= norm_fixed;      // GBA: err /= (Npt_obs+Nls_obs) = 0 every iteration (:3662)
            plba_trace_rec tr{}; tr.window = window; tr.iter = iters; tr.chi = err; tr.lambda = lambda; tr.err_pt = point_error; tr.err_ls = line_error;
            if (std::fabs(err - err_prev) < th_err_change || err < th_err) { tr.stop = 1; push_trace(res, tr); break; }   // :2798 / :3664
            solve(lambda, xp, xlp, xll);                          // :2801-2806
            if (err > err_prev) lambda /= lambda_k;               // :2809-2811 (Q2)
            else { lambda *= lambda_k; retract(xp, xlp, xll); tr.accepted = 1; }
            tr.dx_norm = vnorm(xp, xlp, xll);
            if (tr.dx_norm < th_dx) { tr.stop = 2; push_trace(res, tr); break; }   // :2834 / :3694
            push_trace(res, tr);
            err_prev = err;
        }
        // write-back :2849-2882 (vo_status treated as VO_PROCESSING, Q17)
        if (res->kf_T_wc) for (int k = 0; k < P.n_kf; k++) {
            int s = P.kf_slot[k];
            if (s >= 0) rows_from_T(expmap_se3(Xkf[s]), res->kf_T_wc + 12 * k);
            else for (int i = 0; i < 12; i++) res->kf_T_wc[12 * k + i] = P.kf_T_wc[12 * k + i];
        }
        if (res->x_pose) for (int s = 0; s < P.n_free; s++) for (int i = 0; i < 6; i++) res->x_pose[6 * s + i] = Xkf[s][i];
        for (int l = 0; l < Npt; l++) {
            double d2 = 0; for (int i = 0; i < 3; i++) { double d = Xpt[l][i] - P.pt_xyz[3 * l + i]; d2 += d * d; }
            if (res->pt_inlier) res->pt_inlier[l] = (!gba && std::sqrt(d2) > 0.01) ? 0 : 1;      // GBA has no inlier rule (:3705-3726)
            if (res->pt_xyz) for (int i = 0; i < 3; i++) res->pt_xyz[3 * l + i] = Xpt[l][i];
        }
        for (int l = 0; l < Nls; l++) {
            if (!plk) {
                double d2 = 0; for (int i = 0; i < 6; i++) { double d = Xls[6 * l + i] - P.ls_end[6 * l + i]; d2 += d * d; }
                if (res->ls_inlier) res->ls_inlier[l] = (!gba && std::sqrt(d2) > 0.01) ? 0 : 1;
                if (res->ls_end) for (int i = 0; i < 6; i++) res->ls_end[6 * l + i] = Xls[6 * l + i];
            } else {
                // :2185-2196: DX = X_line - orthNDw ; inlier rule on ||DX|| ; Q9: NDw = changeOrthToPluker(DX)
                V4 X4, DX; double d2 = 0;
                for (int i = 0; i < 4; i++) { X4[i] = Xls[4 * l + i]; DX[i] = X4[i] - orth0[l][i]; d2 += DX[i] * DX[i]; }
                if (res->ls_inlier) res->ls_inlier[l] = (std::sqrt(d2) > 0.01) ? 0 : 1;
                if (res->ls_orth) for (int i = 0; i < 4; i++) res->ls_orth[4 * l + i] = X4[i];
                if (res->ls_plk) { V6 pl = changeOrthToPluker(fixed ? X4 : DX); for (int i = 0; i < 6; i++) res->ls_plk[6 * l + i] = pl[i]; }
            }
        }
        return PLBA_OK;
    }
};

}  // namespace

// =====================================================================================================
// C entry points (ctypes)
// =====================================================================================================
extern "C" {

void plba_oracle_set_threads(int n) {
#ifdef _OPENMP
    g_threads = n > 0 ? n : omp_get_max_threads();
#else
    (void)n; g_threads = 1;
#endif
}
int plba_oracle_get_threads(void) { return nthreads(); }

// flags: bit0 = literal dense full-system solve instead of Schur (small problems only)
int plba_oracle_solve(const plba_problem *prob, const plba_options *opt, plba_result *res, int flags) {
    res->n_trace = 0; res->n_trials = 0;
    int v = validate(*prob);
    if (v != PLBA_OK) { res->status = v; return v; }
    if (prob->n_pobs + prob->n_lobs == 0) { res->status = PLBA_DISCARDED; return PLBA_DISCARDED; }   // src/mapHandler.cpp:1496-1500
    int rc;
    if (opt->profile == PLBA_PROFILE_G) { GRun g(*prob, *opt, res, 0); g.dense_solve = flags & 1; rc = g.run(); }
    else { HRun h(*prob, *opt, res, 0); h.dense_solve = flags & 1; rc = h.run(); }
    res->status = rc;
    return rc;
}

int plba_oracle_solve_batch(int n, const plba_problem *probs, const plba_options *opt, plba_result *res, int flags) {
    int rc_all = PLBA_OK;
    const int T = nthreads();
    const int saved = g_threads;
    // windows are independent: parallelise across them, each window single-threaded inside
    g_threads = 1;
#pragma omp parallel for schedule(dynamic) num_threads(T) if (T > 1)
    for (int w = 0; w < n; w++) {
        int rc = plba_oracle_solve(&probs[w], opt, &res[w], flags);
        if (res[w].trace) for (int i = 0; i < std::min(res[w].n_trace, res[w].trace_cap); i++) res[w].trace[i].window = w;
        if (rc < PLBA_DISCARDED) {
#pragma omp critical
            rc_all = rc;
        }
    }
    g_threads = saved;
    return rc_all;
}

// One linearisation + Schur of profile G at the INITIAL estimate with Huber on: exports the reduced camera system
// (dense grid of 6x6 blocks, upper part filled) and g_red, plus chi2.  Used by the sharding tests (sum over shards == whole).
int plba_oracle_reduced_system_G(const plba_problem *prob, const plba_options *opt, double lambda, double *S_out, double *g_out, double *chi_out) {
    plba_result dummy{}; GRun g(*prob, *opt, &dummy, 0);
    const plba_problem &P = *prob;
    g.Tcw.resize(P.n_kf);
    for (int k = 0; k < P.n_kf; k++) g.Tcw[k] = inverse_se3(T_from_rows(P.kf_T_wc + 12 * k));
    g.pts.resize(P.n_pt); for (int l = 0; l < P.n_pt; l++) for (int i = 0; i < 3; i++) g.pts[l][i] = P.pt_xyz[3 * l + i];
    g.orth.resize(P.n_ls); for (int l = 0; l < P.n_ls; l++) { V6 pl; for (int i = 0; i < 6; i++) pl[i] = P.ls_plk[6 * l + i]; g.orth[l] = changePlukerToOrth(pl); }
    g.e_p.resize(P.n_pobs); g.e_l.resize(P.n_lobs); g.lvl_p.assign(P.n_pobs, 0); g.lvl_l.assign(P.n_lobs, 0);
    g.om_p.assign(P.n_pobs, 1.0); g.om_l.assign(P.n_lobs, 1.0);
    for (int i = 0; i < P.n_pobs; i++) if (P.po_sig2) g.om_p[i] = (double)(float)(1.0 / P.po_sig2[i]);
    for (int i = 0; i < P.n_lobs; i++) if (P.lo_sig2) g.om_l[i] = (double)(float)(1.0 / P.lo_sig2[i]);
    init_system(P, 4, g.B);
    build_skyline(P, g.B.ptr_p, g.B.ptr_l, g.S);
    // a shard may not reach the whole envelope: use a full lower triangle so shards are summable
    g.S.n = 6 * P.n_free; g.S.first.assign(g.S.n, 0); g.S.rowptr.resize(g.S.n + 1);
    int64_t off = 0; for (int i = 0; i < g.S.n; i++) { g.S.rowptr[i] = off; off += i + 1; } g.S.rowptr[g.S.n] = off; g.S.val.assign(off, 0.0);
    g.computeActiveErrors();
    if (chi_out) *chi_out = g.activeRobustChi2(true);
    g.buildSystem(true);
    std::vector<double> xp, a, b, Sd, gv;
    solve_schur(P, g.B, g.S, lambda, false, xp, a, b, &Sd, &gv, false);
    if (S_out) for (size_t i = 0; i < Sd.size(); i++) S_out[i] = Sd[i];
    if (g_out) for (size_t i = 0; i < gv.size(); i++) g_out[i] = gv[i];
    return PLBA_OK;
}

// ---- unit-level exports for tests (finite differences, round trips) ----
void plba_oracle_expmap_se3(const double *x6, double *T16) { V6 x; for (int i = 0; i < 6; i++) x[i] = x6[i]; M4 T = expmap_se3(x); for (int i = 0; i < 16; i++) T16[i] = T[i]; }
void plba_oracle_logmap_se3(const double *T16, double *x6) { M4 T; for (int i = 0; i < 16; i++) T[i] = T16[i]; V6 x = logmap_se3(T); for (int i = 0; i < 6; i++) x6[i] = x[i]; }
void plba_oracle_inverse_se3(const double *T16, double *o16) { M4 T; for (int i = 0; i < 16; i++) T[i] = T16[i]; M4 o = inverse_se3(T); for (int i = 0; i < 16; i++) o16[i] = o[i]; }
void plba_oracle_pluker_to_orth(const double *pl6, double *o4) { V6 p; for (int i = 0; i < 6; i++) p[i] = pl6[i]; V4 o = changePlukerToOrth(p); for (int i = 0; i < 4; i++) o4[i] = o[i]; }
void plba_oracle_orth_to_pluker(const double *o4, double *pl6) { V4 o; for (int i = 0; i < 4; i++) o[i] = o4[i]; V6 p = changeOrthToPluker(o); for (int i = 0; i < 6; i++) pl6[i] = p[i]; }
// U, W and d(Plücker)/d(orth) of a line; q7 != 0 selects the MapLine:: copy with the flipped sign (src/mapFeatures.cpp:260, quirk Q7)
void plba_oracle_orth_UW_jac(const double *pl6, int q7, double *U9, double *W4, double *J24) {
    V6 p; for (int i = 0; i < 6; i++) p[i] = pl6[i];
    M3 U = getOrhtRFromPluker(p); M2 W = getOrthWFromPluker(p); Mat<6, 4> J = jacobianFromPlukerToOrth(U, W, q7 ? -1.0 : 1.0);
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) U9[3 * i + j] = U(i, j);
    for (int i = 0; i < 2; i++) for (int j = 0; j < 2; j++) W4[2 * i + j] = W(i, j);
    for (int i = 0; i < 6; i++) for (int j = 0; j < 4; j++) J24[4 * i + j] = J(i, j);
}
void plba_oracle_transform_pluker(const double *T16, const double *pl6, double *out6) {
    M4 T; for (int i = 0; i < 16; i++) T[i] = T16[i]; V6 p; for (int i = 0; i < 6; i++) p[i] = pl6[i];
    V6 q = getTransformMatrixForPluker(T) * p; for (int i = 0; i < 6; i++) out6[i] = q[i];
}
void plba_oracle_update_orth(const double *D4, const double *d4, double *out4) { V4 D, d; for (int i = 0; i < 4; i++) { D[i] = D4[i]; d[i] = d4[i]; } V4 p = updateOrthCoord(D, d); for (int i = 0; i < 4; i++) out4[i] = p[i]; }
void plba_oracle_pose_oplus(const double *T16, const double *d6, double *o16) { M4 T; for (int i = 0; i < 16; i++) T[i] = T16[i]; V6 d; for (int i = 0; i < 6; i++) d[i] = d6[i]; M4 o = poseOplusG2O(T, d); for (int i = 0; i < 16; i++) o16[i] = o[i]; }
// EdgePosePoint: e[2], Jxi[2x3], Jxj[2x6]
void plba_oracle_point_edge(const double *cam, const double *Tcw16, const double *Pw3, const double *obs2, double *e2, double *Jxi6, double *Jxj12) {
    M4 T; for (int i = 0; i < 16; i++) T[i] = Tcw16[i]; V3 P; for (int i = 0; i < 3; i++) P[i] = Pw3[i]; V2 ob; ob[0] = obs2[0]; ob[1] = obs2[1];
    V2 e = pointEdgeError(cam, T, P, ob); PointEdgeLin L; pointEdgeLinearize(cam, T, P, L);
    e2[0] = e[0]; e2[1] = e[1]; for (int i = 0; i < 6; i++) Jxi6[i] = L.Jxi[i]; for (int i = 0; i < 12; i++) Jxj12[i] = L.Jxj[i];
}
// EdgePoseLine: e[2], Jxi[2x4], Jxj[2x6]
void plba_oracle_line_edge(const double *cam, const double *Tcw16, const double *orth4, const double *obs4, int faithful, double *e2, double *Jxi8, double *Jxj12) {
    M4 T; for (int i = 0; i < 16; i++) T[i] = Tcw16[i]; V4 o, ob; for (int i = 0; i < 4; i++) { o[i] = orth4[i]; ob[i] = obs4[i]; }
    LineEdgeLin L; lineEdgeLinearize(cam, T, o, ob, faithful != 0, L);
    e2[0] = L.e[0]; e2[1] = L.e[1]; for (int i = 0; i < 8; i++) Jxi8[i] = L.Jxi[i]; for (int i = 0; i < 12; i++) Jxj12[i] = L.Jxj[i];
}
// Profile-H scalar terms: J_p[6], J_l[6], r, w.   kind: 0 point, 1 endpoint line, 2 Plücker line
void plba_oracle_h_term(int kind, const double *cam, const double *Tiw16, const double *lm6, const double *obs4, double th, int fixed, double *Jp6, double *Jl6, double *rw2) {
    M4 T; for (int i = 0; i < 16; i++) T[i] = Tiw16[i];
    HTerm t{};
    if (kind == 0) { V3 X; for (int i = 0; i < 3; i++) X[i] = lm6[i]; V2 ob; ob[0] = obs4[0]; ob[1] = obs4[1]; h_point_term(cam, T, X, ob, th, t); }
    else if (kind == 1) { V3 Pw, Qw; for (int i = 0; i < 3; i++) { Pw[i] = lm6[i]; Qw[i] = lm6[3 + i]; } h_endline_term(cam, T, Pw, Qw, obs4, th, fixed != 0, t); }
    else { V6 nd; for (int i = 0; i < 6; i++) nd[i] = lm6[i]; h_plkline_term(cam, T, nd, obs4, th, fixed != 0, t); }
    for (int i = 0; i < 6; i++) { Jp6[i] = t.J_p[i]; Jl6[i] = t.J_l[i]; }
    rw2[0] = t.r; rw2[1] = t.w;
}

}  // extern "C"
