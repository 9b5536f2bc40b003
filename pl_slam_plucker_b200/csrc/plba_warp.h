// Warp-autonomous assembly / update kernels (the default path; the CTA-chunk kernels of plba_kernels.h remain for windows
// that contain a landmark with more than 32 observations).
//
// Work decomposition.  After the signature sort (plba_upload) landmarks seen by exactly the same keyframe sequence are
// adjacent: a RUN.  A work ITEM is a slice of a run (<= a few passes); one warp owns an item from its first load to its
// last red.global.add.  Inside an item every landmark has the same k observations in the same keyframes, so
//   * lane t of a pass handles observation (landmark t / k, track position t % k): observation loads are contiguous
//     (128-bit, coalesced), no per-observation index is read at all (the structure comes from the item descriptor),
//   * the pose a lane works with never changes during the item,
//   * the Schur tasks (pose pair (i <= j) of the track x two halves of three block rows) are the same for every landmark
//     of the item: a lane keeps its 3x6 block IN REGISTERS across all passes and leaves the SM once per item.
// Phases of a pass are separated by __syncwarp only (no block barrier anywhere in the loop): the 16 resident warps of an SM
// drift apart and hide each other's latencies; the per-landmark work (H_ll, its inverse, v = H_ll^-1 b_l) is done
// redundantly by all lanes of the landmark, which costs the same FP64-pipe slots as doing it on one lane and saves the
// shared-memory round trip.
#pragma once
#include "plba_kernels.h"

namespace plba {

#ifndef PLBA_W_CTAS
#define PLBA_W_CTAS 3      // resident 4-warp CTAs per SM the kernels are compiled for (register budget = 65536 / (128 * PLBA_W_CTAS))
#endif
#ifndef PLBA_WU_CTAS
#define PLBA_WU_CTAS 4      // the update kernel holds no Schur blocks: 128 registers suffice (measured: -5 % at config 5)
#endif
#ifndef PLBA_W_PASSES
#define PLBA_W_PASSES 8
#endif
#ifndef PLBA_W_WARPS
#define PLBA_W_WARPS 4
#endif
enum { WARPS_PER_CTA = PLBA_W_WARPS, WNT = 32 * WARPS_PER_CTA, W_CTAS_PER_SM = PLBA_W_CTAS, W_MAX_TRACK = 32, W_ITEM_PASSES_MAX = PLBA_W_PASSES };

// per-warp shared memory: [component][lane] so that lane-indexed accesses are conflict free
template <int PROF, int LT>
struct WSmem {
    typedef KT<PROF, LT> K;
    enum { NA = K::RANK * 6, NB = K::RANK * K::D, NE = K::RANK, DOUBLES = (NA + 2 * NB + NE + K::D) * 32, INTS = 64 };
    double *A, *B, *TA, *E, *V;
    int *slot, *fpos;
    static PLBA_HD size_t bytes() { return sizeof(double) * DOUBLES + sizeof(int) * INTS; }
    PLBA_HD explicit WSmem(unsigned char *raw) {
        double *p = (double *)raw;
        A = p; p += NA * 32; B = p; p += NB * 32; TA = p; p += NB * 32; E = p; p += NE * 32; V = p; p += K::D * 32;
        slot = (int *)p; fpos = slot + 32;
    }
};
template <int PROF, int LT> struct WRec;
template <int PROF> struct WSmemMax {      // per warp: the larger of the update kernel's [component][lane] tile and the assembly kernel's record array
    static PLBA_HD size_t bytes();
};

// ---- landmark-only quantities, in registers (the warp path's form of lm_precompute) -------------------------------
template <int LT> struct LmD;
template <> struct LmD<LT_POINT> { double Pw[3]; };
template <> struct LmD<LT_LINE_ORTH> { LinePre L; double o[4]; };
template <> struct LmD<LT_LINE_END> { double Pw[3], Qw[3]; };

template <int PROF, int LT> struct LmLoad;
template <int PROF> struct LmLoad<PROF, LT_POINT> {
    static PLBA_HD void run(const DevP &, const WinCtrl &, int, int lm, const double *state, LmD<LT_POINT> &d) {
        for (int i = 0; i < 3; i++) d.Pw[i] = state[(size_t)3 * lm + i];
    }
};
template <int PROF> struct LmLoad<PROF, LT_LINE_ORTH> {
    static PLBA_HD void run(const DevP &P, const WinCtrl &ctl, int, int lm, const double *state, LmD<LT_LINE_ORTH> &d) {
        for (int i = 0; i < 4; i++) d.o[i] = state[(size_t)4 * lm + i];
        if (PROF == PLBA_PROFILE_H_PLK && ctl.iter == 0) {      // pass 0 reads map NDw (:1744)
            double pl[6];
            for (int i = 0; i < 6; i++) pl[i] = P.lns_map[(size_t)6 * lm + i];
            line_pre_from_plk(pl, d.L);
        } else {
            // Plücker vector, U and W of the line at the CURRENT state: written once per landmark and state by k_reset / the update
            // kernel (lane = landmark, all 32 lanes busy with the trigonometry) instead of once per observation and kernel
            const double *c = P.lpre[ctl.cur] + (size_t)LPRE_N * lm;
            for (int i = 0; i < 3; i++) { d.L.n[i] = c[i]; d.L.d[i] = c[3 + i]; d.L.u1[i] = c[6 + i]; d.L.u2[i] = c[9 + i]; d.L.u3[i] = c[12 + i]; }
            d.L.w1 = c[15]; d.L.w2 = c[16];
        }
    }
};
template <int PROF> struct LmLoad<PROF, LT_LINE_END> {
    static PLBA_HD void run(const DevP &P, const WinCtrl &ctl, int win, int lm, const double *state, LmD<LT_LINE_END> &d) {
        // inside the loop the reference reads BOTH endpoints from offset 3*loc (Q3, :2697-2698)
        const bool q3 = (!P.fixed_quirks && ctl.iter > 0);
        const int l0 = P.win_ls0[win];
        const size_t q3off = (size_t)6 * l0 + (size_t)3 * (lm - l0);      // endpoint lines are never re-ordered (see plba_upload)
        for (int i = 0; i < 3; i++) {
            d.Pw[i] = q3 ? state[q3off + i] : state[(size_t)6 * lm + i];
            d.Qw[i] = q3 ? state[q3off + i] : state[(size_t)6 * lm + 3 + i];
        }
    }
};

template <int LT> struct ObsLoad;
template <> struct ObsLoad<LT_POINT> {
    double uv[2];
    PLBA_HD void load(const DevP &P, int o) { const plba_d2 v = *(const plba_d2 *)(P.po_uv + (size_t)2 * o); uv[0] = v.x; uv[1] = v.y; }
};
template <> struct ObsLoad<LT_LINE_ORTH> {
    double ab[4];
    PLBA_HD void load(const DevP &P, int o) {
        const plba_d2 a0 = *(const plba_d2 *)(P.lo_ab + (size_t)4 * o), a1 = *(const plba_d2 *)(P.lo_ab + (size_t)4 * o + 2);
        ab[0] = a0.x; ab[1] = a0.y; ab[2] = a1.x; ab[3] = a1.y;
    }
};
template <> struct ObsLoad<LT_LINE_END> : ObsLoad<LT_LINE_ORTH> {};

template <int N> PLBA_HD void wrec_ld(const double *p, double *r) {      // N even, p 16-byte aligned
#pragma unroll
    for (int i = 0; i < N / 2; i++) { const plba_d2 v = ((const plba_d2 *)p)[i]; r[2 * i] = v.x; r[2 * i + 1] = v.y; }
}
template <int N> PLBA_HD void wrec_st(double *p, const double *r) {
#pragma unroll
    for (int i = 0; i < N / 2; i++) { plba_d2 v; v.x = r[2 * i]; v.y = r[2 * i + 1]; ((plba_d2 *)p)[i] = v; }
}

// one observation -> scaled rows At = sqrt(w) J_pose, Bt = sqrt(w) J_lm, et = sqrt(w) e, in registers (same arithmetic as
// obs_linearize; see there for the citations).  Returns false for an edge gated out of the active set (rows are zero).
template <int PROF, int LT>
PLBA_HD bool obs_lin_w(const DevP &P, const WinCtrl &ctl, int o, int kf, const LmD<LT> &lmd, const ObsLoad<LT> &ob, double *A, double *B, double *e, double &cost) {
    typedef KT<PROF, LT> K;
    typedef ObsAcc<LT> OA;
    double wsq = 0.0;
    bool active = true;
    cost = 0.0;
    if (PROF == PLBA_PROFILE_G) {
        if (ctl.stage == 1 && OA::lvl(P)[o]) active = false;
        if (active) {
            // the pose as six 128-bit loads (L1 hits, lanes of one keyframe share them): half the load instructions of twelve scalar reads
            double T[12];
            wrec_ld<12>(P.poseT[ctl.cur] + (size_t)12 * kf, T);
            if constexpr (LT == LT_POINT) {
                g_point_lin(P.cam, T, lmd.Pw, ob.uv, e, A, B);
            } else if constexpr (LT == LT_LINE_ORTH) {
                const LmD<LT_LINE_ORTH> &d = lmd; const ObsLoad<LT_LINE_ORTH> &q = ob;
                double head[3], tail[3];
                if (P.fixed_quirks) { for (int i = 0; i < 3; i++) { head[i] = d.L.n[i]; tail[i] = d.L.d[i]; } }
                else { for (int i = 0; i < 3; i++) { head[i] = d.o[i]; tail[i] = d.o[1 + i]; } }   // Q12
                g_line_lin(P.cam, T, d.L, head, tail, q.ab, e, A, B);
            }
            const double om = OA::om(P)[o];
            const double chi2 = om * (e[0] * e[0] + e[1] * e[1]);
            double rho0 = chi2, rho1 = 1.0;
            if (ctl.stage == 0) huber(P.huber_delta, chi2, rho0, rho1);
            cost = rho0;
            { const double ww = rho1 * om; wsq = (ww > 1e-290 && ww < 1e290) ? ww * plba_rsqrt_fast(ww) : sqrt(ww); }      // sqrt(rho1 Omega)
        }
    } else {
        const bool pass0 = (ctl.iter == 0);
        double r, w, Jp[6], Jl[6];
        if constexpr (LT == LT_POINT) {
            const double *T = P.poseT[ctl.cur] + (size_t)12 * kf;
            h_point(P.cam, T, lmd.Pw, ob.uv, P.homog_th, Jp, Jl, r, w);
        } else {
            // Q4: inside the loop the line terms keep the MAP pose (:2700, :2010)
            const double *T = (pass0 || P.fixed_quirks) ? P.poseT[ctl.cur] + (size_t)12 * kf : P.kf_Tmap + (size_t)12 * kf;
            if constexpr (LT == LT_LINE_END) {
                const double th = (pass0 || P.fixed_quirks) ? P.homog_th : 0.0000001;
                h_endline(P.cam, T, lmd.Pw, lmd.Qw, ob.ab, th, P.fixed_quirks != 0, Jp, Jl, r, w);
            } else {
                h_plkline(P.cam, T, lmd.L, ob.ab, P.homog_th, P.fixed_quirks != 0, Jp, Jl, r, w);
            }
        }
        wsq = sqrt(w);
        for (int c = 0; c < 6; c++) A[c] = Jp[c];
        for (int c = 0; c < K::D; c++) B[c] = Jl[c];
        e[0] = -r;                 // g += J r w  ==  b = -J^T (w e) with e = -r
        cost = w * r * r;
    }
    // (an edge gated out of the active set leaves zero rows: one branch instead of a select per entry — 40 selects per observation)
    if (active) {
        for (int i = 0; i < K::RANK * 6; i++) A[i] *= wsq;
        for (int i = 0; i < K::RANK * K::D; i++) B[i] *= wsq;
        for (int i = 0; i < K::RANK; i++) e[i] *= wsq;
    } else {
        for (int i = 0; i < K::RANK * 6; i++) A[i] = 0.0;
        for (int i = 0; i < K::RANK * K::D; i++) B[i] = 0.0;
        for (int i = 0; i < K::RANK; i++) e[i] = 0.0;
    }
    return active;
}

// inputs of one observation of a pass (measurement, information, level flag, landmark state / line cache) pulled towards L1
template <int PROF, int LT>
PLBA_D void prefetch_obs_w(const DevP &P, const WinCtrl &ctl, int o, int lm, const double *state) {
    typedef ObsAcc<LT> OA;
    if (LT == LT_POINT) { plba_prefetch_l1(P.po_uv + (size_t)2 * o); plba_prefetch_l1(state + (size_t)3 * lm); }
    else {
        plba_prefetch_l1(P.lo_ab + (size_t)4 * o); plba_prefetch_l1(state + (size_t)KT<PROF, LT>::D * lm);
        if (LT == LT_LINE_ORTH) { const double *c = P.lpre[ctl.cur] + (size_t)LPRE_N * lm; plba_prefetch_l1(c); plba_prefetch_l1(c + 16); }
    }
    plba_prefetch_l1(OA::om(P) + o);
    if (PROF == PLBA_PROFILE_G && ctl.stage == 1) plba_prefetch_l1(OA::lvl(P) + o);
}

// ---- assembly: shared-memory record of one observation (one per lane and pass) ---------------------------------------
// Array-of-structures, every sub-array 16-byte aligned, so that the Schur phase (which re-reads the records of the two poses of its
// pair for every landmark) moves them with 128-bit shared-memory loads: half the instructions and half the wavefronts of the
// [component][lane] layout, and the kernel's first limiter was the shared-memory pipe (ncu: LSU wavefronts 66 % of peak at config 5,
// FP64 pipe 36 %).  The record stride is even with an odd half, so that the 128-bit accesses of eight consecutive lanes fall into
// eight different 16-byte bank groups (conflict-free stores of the linearisation phase).
template <int PROF, int LT>
struct WRec {
    typedef KT<PROF, LT> K;
    // A (the pose rows) is stored as [column half][row of the residual][3]: a Schur task that owns three columns of a 6x6 block reads
    // "its" half of A_b as one aligned run (two residual rows: 6 doubles, no padding; one row: 3 + one pad)
    enum { RANK = K::RANK, D = K::D, NAH = (RANK * 3 + 1) & ~1, AHS = NAH / RANK, NA = 2 * NAH, NB = RANK * D,      // AHS: row stride inside a half (3, or 4 with a pad when RANK * 3 is odd)
           NBp = (NB + 1) & ~1, NEp = (RANK + 1) & ~1, NVp = (D + 1) & ~1,
           NVp_ = NVp, OA = 0, OB = NA, OT = OB + NBp, OE = OT + NBp, OV = OE + NEp, SIZE = OV + NVp, STRIDE = ((SIZE / 2) % 2 == 1) ? SIZE : SIZE + 2,
           NACC = 18, NG = 3 };
    static PLBA_HD size_t bytes() { return sizeof(double) * STRIDE * 32 + sizeof(int) * 64; }
};

// Schur task of a lane: pose pair (i <= j) of the free track positions x one COLUMN half of the 6x6 block (columns hcol .. hcol + 2).
// Both halves need M = Ta B^T (4 D FMAs), but each computes only its three columns of M A_b and of the block: 60 FMAs per landmark
// and half against the 72 of the round-1 split by rows (where both halves computed all of M A_b), with 18 accumulators per lane
// (whole blocks, 36 accumulators, spill around the linearisation at 168 registers: measured 2.10 against 1.92 ms at config 5).
struct WTask { int on, pa, pb, sa, sb, half, diag, slice; };
PLBA_HD void wtask_decode(int task, int nf, const int *fpos, const int *slot, int slot0, WTask &t) {
    const int p = task >> 1;
    t.half = task & 1;
    int q = p, i = 0;
    while (q >= nf - i) { q -= nf - i; i++; }
    const int j = i + q;
    t.pa = fpos[i]; t.pb = fpos[j];
    t.sa = slot[t.pa] - slot0; t.sb = slot[t.pb] - slot0;
    t.diag = (i == j) ? 1 : 0;
}

// accumulate the landmarks m = first, first + step, ... < nlp of the pass into the lane's 6x3 half block blk[r * 3 + cc] (and, on the
// diagonal, its three entries of g / diag(H_pp))
template <int PROF, int LT, int mode>
PLBA_D void wtask_accumulate(const double *rec, const WTask &t, int k, int first, int step, int nlp, double *blk, double *gv, double *hd) {
    typedef WRec<PROF, LT> R;
    const int D = R::D, RANK = R::RANK;
    for (int m = first; m < nlp; m += step) {
        const double *ra = rec + (size_t)(m * k + t.pa) * R::STRIDE, *rb = rec + (size_t)(m * k + t.pb) * R::STRIDE;
        if ((PROF != PLBA_PROFILE_G || mode == 0) && t.diag) {        // diag(H_pp): lambda init (all profiles) and the hand LM's multiplicative damping
            double Ah[R::NAH];
            wrec_ld<R::NAH>(ra + R::OA + t.half * R::NAH, Ah);
#pragma unroll
            for (int cc = 0; cc < 3; cc++) {
#pragma unroll
                for (int kk = 0; kk < RANK; kk++) hd[cc] += Ah[kk * R::AHS + cc] * Ah[kk * R::AHS + cc];
            }
        }
        if (mode == 0) continue;
        double MA[RANK * 3];
        {
            double M[RANK * RANK], Ta[R::NBp], Bb[R::NBp], Abh[R::NAH];
            wrec_ld<R::NBp>(ra + R::OT, Ta); wrec_ld<R::NBp>(rb + R::OB, Bb); wrec_ld<R::NAH>(rb + R::OA + t.half * R::NAH, Abh);
#pragma unroll
            for (int kk = 0; kk < RANK; kk++) {
#pragma unroll
                for (int k2 = 0; k2 < RANK; k2++) {
                    double sum = (t.diag && kk == k2) ? -1.0 : 0.0;          // diagonal: A^T (Ta B^T - I) A, subtracted below
#pragma unroll
                    for (int mm = 0; mm < D; mm++) sum += Ta[kk * D + mm] * Bb[k2 * D + mm];
                    M[kk * RANK + k2] = sum;
                }
            }
            if (t.diag) {
                // g_a += -At^T (et + Bt v): this lane's three entries
                double ev[R::NEp], V[R::NVp];
                wrec_ld<R::NEp>(ra + R::OE, ev); wrec_ld<R::NVp>(ra + R::OV, V);
#pragma unroll
                for (int kk = 0; kk < RANK; kk++) {
#pragma unroll
                    for (int mm = 0; mm < D; mm++) ev[kk] += Bb[kk * D + mm] * V[mm];
                }
#pragma unroll
                for (int cc = 0; cc < 3; cc++) {
#pragma unroll
                    for (int kk = 0; kk < RANK; kk++) gv[cc] -= Abh[kk * R::AHS + cc] * ev[kk];      // (diagonal task: record b == record a)
                }
            }
#pragma unroll
            for (int kk = 0; kk < RANK; kk++) {
#pragma unroll
                for (int cc = 0; cc < 3; cc++) {
                    double sum = 0;
#pragma unroll
                    for (int k2 = 0; k2 < RANK; k2++) sum += M[kk * RANK + k2] * Abh[k2 * R::AHS + cc];
                    MA[kk * 3 + cc] = sum;
                }
            }
        }
#pragma unroll
        for (int h = 0; h < 2; h++) {
            double Aa[R::NAH];
            wrec_ld<R::NAH>(ra + R::OA + h * R::NAH, Aa);
#pragma unroll
            for (int rr = 0; rr < 3; rr++) {
#pragma unroll
                for (int kk = 0; kk < RANK; kk++) {
                    const double av = Aa[kk * R::AHS + rr];
#pragma unroll
                    for (int cc = 0; cc < 3; cc++) blk[(3 * h + rr) * 3 + cc] -= av * MA[kk * 3 + cc];
                }
            }
        }
    }
}

// the lane's half block leaves the SM.  blk holds MINUS the Schur term, so S(a,b) += blk (upper storage; transposed if the pair is
// stored the other way round), g, diag(H_pp).  Two observations of one landmark in the same keyframe (sa == sb off the track
// diagonal: rare) put blk + blk^T onto the diagonal block.
template <int PROF, int mode>
PLBA_D void wtask_flush(const DevP &P, const WTask &t, const SWin &swin, int slot0, const double *blk, const double *gv, const double *hd) {
    const int hcol = 3 * t.half;
    if (mode == 0) {
        if (t.diag) {
#pragma unroll
            for (int cc = 0; cc < 3; cc++) plba_atomic_add(&P.hpp_diag_init[(size_t)6 * (slot0 + t.sa) + hcol + cc], hd[cc]);
        }
        return;
    }
    const bool tr = (t.sa > t.sb);
    const int ra = tr ? t.sb : t.sa, cb = tr ? t.sa : t.sb;
    const SBlk sb_ = s_block(swin, ra, cb);                        // dense upper storage, or the solver's node form (large banded windows)
    const long long sr = tr ? sb_.sc : sb_.sr, sc = tr ? sb_.sr : sb_.sc;
    double *p0 = sb_.p + hcol * sc;
    const bool upper_only = (t.sa == t.sb);                        // diagonal block of S: only its upper triangle is stored
#pragma unroll
    for (int r = 0; r < 6; r++) {
#pragma unroll
        for (int cc = 0; cc < 3; cc++) {
            if (!upper_only || r <= hcol + cc) plba_atomic_add(p0 + r * sr + cc * sc, blk[r * 3 + cc]);
        }
    }
    if (t.diag) {
#pragma unroll
        for (int cc = 0; cc < 3; cc++) {
            plba_atomic_add(&P.gs[(size_t)6 * (slot0 + t.sa) + hcol + cc], gv[cc]);
            if (PROF != PLBA_PROFILE_G) plba_atomic_add(&P.hpp_diag[(size_t)6 * (slot0 + t.sa) + hcol + cc], hd[cc]);
        }
    } else if (upper_only) {
        // transposed copy of an off-track-diagonal block that landed on the diagonal of S
#pragma unroll
        for (int r = 0; r < 6; r++) {
#pragma unroll
            for (int cc = 0; cc < 3; cc++) {
                if (hcol + cc <= r) plba_atomic_add(sb_.p + (size_t)(hcol + cc) * sb_.sr + r * sb_.sc, blk[r * 3 + cc]);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// assembly: mode 0 = diagonal pass for the initial lambda (computeLambdaInit / Hmax, :2555-2561); mode 1 = full
// ---------------------------------------------------------------------------------------------------------
template <int PROF, int LT, int mode>
PLBA_D void assemble_items_w(const DevP &Pin, const WItem *items, int n_items, unsigned char *wraw) {
    PLBA_PARAMS_REF(P, Pin);
    typedef KT<PROF, LT> K;
    typedef ObsAcc<LT> OA;
    typedef WRec<PROF, LT> R;
    const int D = K::D, RANK = K::RANK;
    double *rec = (double *)wraw;                                 // [32] records of the pass
    int *slot_s = (int *)(rec + (size_t)R::STRIDE * 32), *fpos_s = slot_s + 32;
    LANE_VAR(double, cost_l); LANE_VAR(double, maxd_l);
    LANE_ARR(double, blk, R::NACC); LANE_ARR(double, gv, R::NG); LANE_ARR(double, hd, R::NG);
    LANE_VAR(WTask, tk);
    LANE_VAR(int, kf_l); LANE_VAR(int, m_l);
    int cur_win = -1;
    WPHASE_BEGIN
        LANE_BIND(cost_l); LANE_BIND(maxd_l);
        cost_l = 0.0; maxd_l = 0.0;
    WPHASE_END
    for (;;) {
        // next work item: one atomic per warp (items differ in cost, a static deal leaves the slowest warp far behind)
        WPHASE_BEGIN
            if (lane == 0) slot_s[48] = plba_atomic_fetch_add_i(&P.counters[LT == LT_POINT ? CNT_WORK_PT : CNT_WORK_LS], 1);
        WPHASE_END
        const int ii = slot_s[48];
        WPHASE_BEGIN
        WPHASE_END
        if (ii >= n_items) break;
        const WItem it = items[ii];
        const WinCtrl ctl = P.ctrl[it.win];             // the controller only runs between launches: a private copy is safe
        if (ctl.done) continue;
        if (mode == 0 && !ctl.need_init) continue;
        if (it.k == 0) continue;                                  // landmarks without observations add nothing to the system
        if (it.win != cur_win) {
            if (cur_win >= 0) {
                WPHASE_BEGIN
                    LANE_BIND(cost_l); LANE_BIND(maxd_l);
                    double *accw = P.acc + (size_t)4 * cur_win;
                    if (mode == 1) PLBA_WARP_FLUSH_ADD((PROF == PLBA_PROFILE_G) ? &accw[ACC_CHI_LIN] : (LT == LT_POINT ? &accw[ACC_ERR_PT] : &accw[ACC_ERR_LS]), cost_l);
                    else PLBA_WARP_FLUSH_MAX(&P.accmax[cur_win], maxd_l);
                    cost_l = 0.0; maxd_l = 0.0;
                WPHASE_END
            }
            cur_win = it.win;
        }
        const int k = it.k, nf = it.nfree;
        const int lpp = 32 / k;                                  // landmarks per pass
        const int npass = (it.n_lm + lpp - 1) / lpp;
        const int ntask = nf * (nf + 1);                          // (pairs i <= j) x 2 column halves
        const int rounds = (ntask + 31) >> 5;
        const bool keep = (rounds == 1);                          // the lane's block stays in registers for the whole item
        const int nslice = keep ? 32 / ntask : 1;                 // few tasks: the landmarks of a pass are dealt over several lanes per task
        const int slot0 = P.win_slot0[it.win];
        const double *state = OA::state(P, ctl.cur);
        const SWin swin = s_window(P, it.win);
        WPHASE_BEGIN
            LANE_BIND(kf_l); LANE_BIND(m_l);
            m_l = lane / k;
            kf_l = OA::kf(P)[it.ob0 + (lane - m_l * k)];
            if (lane < k) slot_s[lane] = P.kf_slot[kf_l];
            if (lane < nf) fpos_s[lane] = OA::freepos(P)[it.fp0 + lane];
        WPHASE_END
        if (keep) {
            WPHASE_BEGIN
                LANE_BIND(tk); LANE_BIND(blk); LANE_BIND(gv); LANE_BIND(hd);
                const int slice = lane / ntask, task = lane - slice * ntask;
                tk.on = (slice < nslice) ? 1 : 0; tk.slice = slice;
                if (tk.on) wtask_decode(task, nf, fpos_s, slot_s, slot0, tk);
#pragma unroll
                for (int i = 0; i < R::NACC; i++) blk[i] = 0.0;
#pragma unroll
                for (int i = 0; i < R::NG; i++) { gv[i] = 0.0; hd[i] = 0.0; }
            WPHASE_END
        }
        for (int pass = 0; pass < npass; pass++) {
            const int lmb = pass * lpp;
            const int nlp = (it.n_lm - lmb < lpp) ? it.n_lm - lmb : lpp;
            const int nvalid = nlp * k;
            // ---- per observation: residual, Jacobians, robust weight (subsystem 1) ----
            WPHASE_BEGIN
                LANE_BIND(cost_l); LANE_BIND(kf_l); LANE_BIND(m_l);
                if (lane < nvalid) {
                    const int o = it.ob0 + lmb * k + lane;
                    LmD<LT> lmd; ObsLoad<LT> ob;
                    ob.load(P, o);
                    LmLoad<PROF, LT>::run(P, ctl, it.win, it.lm0 + lmb + m_l, state, lmd);
                    double A[RANK * 6], B[R::NBp], e[R::NEp], cost;
                    if (R::NBp > RANK * D) B[R::NBp - 1] = 0.0;
                    if (R::NEp > RANK) e[R::NEp - 1] = 0.0;
                    obs_lin_w<PROF, LT>(P, ctl, o, kf_l, lmd, ob, A, B, e, cost);
                    if (mode == 1) cost_l += cost;
                    double *r = rec + (size_t)lane * R::STRIDE;
                    double Ah[R::NA];
#pragma unroll
                    for (int h = 0; h < 2; h++) {
#pragma unroll
                        for (int kk = 0; kk < RANK; kk++) {
#pragma unroll
                            for (int cc = 0; cc < 3; cc++) Ah[h * R::NAH + kk * R::AHS + cc] = A[kk * 6 + 3 * h + cc];
                        }
                        if (R::NAH > RANK * 3) Ah[h * R::NAH + R::NAH - 1] = 0.0;
                    }
                    wrec_st<R::NA>(r + R::OA, Ah); wrec_st<R::NBp>(r + R::OB, B); wrec_st<R::NEp>(r + R::OE, e);
                }
            WPHASE_END
            // ---- per landmark (redundantly on each of its lanes): H_ll, b_l, damping, inverse; Ta = Bt H_ll^-1, v = H_ll^-1 b_l (subsystem 2) ----
            WPHASE_BEGIN
                LANE_BIND(maxd_l); LANE_BIND(m_l);
                if (lane < nvalid) {
                    const int t0 = m_l * k;
                    double H[D * D], bl[D];
#pragma unroll
                    for (int i = 0; i < D * D; i++) H[i] = 0.0;
#pragma unroll
                    for (int i = 0; i < D; i++) bl[i] = 0.0;
                    for (int t = t0; t < t0 + k; t++) {
                        double b[R::NBp], ee[R::NEp];
                        wrec_ld<R::NBp>(rec + (size_t)t * R::STRIDE + R::OB, b); wrec_ld<R::NEp>(rec + (size_t)t * R::STRIDE + R::OE, ee);
#pragma unroll
                        for (int kk = 0; kk < RANK; kk++) {
#pragma unroll
                            for (int r = 0; r < D; r++) {
                                bl[r] -= b[kk * D + r] * ee[kk];
#pragma unroll
                                for (int c = r; c < D; c++) H[r * D + c] += b[kk * D + r] * b[kk * D + c];
                            }
                        }
                    }
                    double maxd = 0.0;
#pragma unroll
                    for (int r = 0; r < D; r++) {
                        if (fabs(H[r * D + r]) > maxd) maxd = fabs(H[r * D + r]);
#pragma unroll
                        for (int c = 0; c < r; c++) H[r * D + c] = H[c * D + r];
                    }
                    if (mode == 0) { if (maxd > maxd_l) maxd_l = maxd; }
                    else {
                        damp_invert<PROF, D>(ctl, H);
                        double *r = rec + (size_t)lane * R::STRIDE;
                        double b[R::NBp], ta[R::NBp], v[R::NVp];
                        wrec_ld<R::NBp>(r + R::OB, b);
                        if (R::NBp > RANK * D) ta[R::NBp - 1] = 0.0;
                        if (R::NVp > D) v[R::NVp - 1] = 0.0;
#pragma unroll
                        for (int kk = 0; kk < RANK; kk++) {
#pragma unroll
                            for (int c = 0; c < D; c++) {
                                double sum = 0;
#pragma unroll
                                for (int mm = 0; mm < D; mm++) sum += b[kk * D + mm] * H[mm * D + c];
                                ta[kk * D + c] = sum;
                            }
                        }
#pragma unroll
                        for (int rr = 0; rr < D; rr++) {
                            double sum = 0;
#pragma unroll
                            for (int c = 0; c < D; c++) sum += H[rr * D + c] * bl[c];
                            v[rr] = sum;
                        }
                        wrec_st<R::NBp>(r + R::OT, ta); wrec_st<R::NVp>(r + R::OV, v);
                    }
                }
            WPHASE_END
            // ---- Schur tasks (subsystem 3): one round when the lane keeps its block for the whole item, else a round per 32 tasks ----
#pragma unroll 1
            for (int r = 0; r < rounds; r++) {
                WPHASE_BEGIN
                    LANE_BIND(tk); LANE_BIND(blk); LANE_BIND(gv); LANE_BIND(hd); LANE_BIND(m_l);
                    if (r == 0 && pass + 1 < npass && lane < lpp * k) {        // the next pass's inputs travel while this pass's Schur tasks run
                        const int lmn = lmb + lpp + m_l;
                        if (lmn < it.n_lm) prefetch_obs_w<PROF, LT>(P, ctl, it.ob0 + (lmb + lpp) * k + lane, it.lm0 + lmn, state);
                    }
                    if (!keep) {
                        const int task = r * 32 + lane;
                        tk.on = (task < ntask) ? 1 : 0; tk.slice = 0;
                        if (tk.on) wtask_decode(task, nf, fpos_s, slot_s, slot0, tk);
#pragma unroll
                        for (int i = 0; i < R::NACC; i++) blk[i] = 0.0;
#pragma unroll
                        for (int i = 0; i < R::NG; i++) { gv[i] = 0.0; hd[i] = 0.0; }
                    }
                    if (tk.on) {
                        wtask_accumulate<PROF, LT, mode>(rec, tk, k, tk.slice, nslice, nlp, blk, gv, hd);
                        if (!keep || pass == npass - 1) wtask_flush<PROF, mode>(P, tk, swin, slot0, blk, gv, hd);
                    }
                WPHASE_END
            }
        }
    }
    if (cur_win >= 0) {
        WPHASE_BEGIN
            LANE_BIND(cost_l); LANE_BIND(maxd_l);
            double *accw = P.acc + (size_t)4 * cur_win;
            if (mode == 1) PLBA_WARP_FLUSH_ADD((PROF == PLBA_PROFILE_G) ? &accw[ACC_CHI_LIN] : (LT == LT_POINT ? &accw[ACC_ERR_PT] : &accw[ACC_ERR_LS]), cost_l);
            else PLBA_WARP_FLUSH_MAX(&P.accmax[cur_win], maxd_l);
        WPHASE_END
    }
}

template <int PROF> PLBA_HD size_t WSmemMax<PROF>::bytes() {
    size_t m = WSmem<PROF, LT_POINT>::bytes();
    const size_t b = WSmem<PROF, LineOf<PROF>::LT>::bytes(), c = WRec<PROF, LT_POINT>::bytes(), d = WRec<PROF, LineOf<PROF>::LT>::bytes();
    if (b > m) m = b;
    if (c > m) m = c;
    if (d > m) m = d;
    return (m + 15) & ~(size_t)15;
}

template <int PROF>
PLBA_KERNEL void PLBA_BOUNDS(WNT, W_CTAS_PER_SM) k_assemble_w(const DevP *Pp, int mode) {
    if (mode == 0 && Pp->counters[CNT_NEED_INIT] == 0) { PLBA_COUNT_LAUNCH(Pp); return; }        // no window waits for its initial lambda (the counter only changes between launches)
    PLBA_SMEM(raw);
    PLBA_PARAMS(P, Pp);
    unsigned char *wraw = raw + (size_t)PLBA_WARP_IN_CTA * WSmemMax<PROF>::bytes();
    if (mode == 0) {
        assemble_items_w<PROF, LT_POINT, 0>(P, P.witems_pt, P.n_witems_pt, wraw);
        assemble_items_w<PROF, LineOf<PROF>::LT, 0>(P, P.witems_ls, P.n_witems_ls, wraw);
    } else {
        assemble_items_w<PROF, LT_POINT, 1>(P, P.witems_pt, P.n_witems_pt, wraw);
        assemble_items_w<PROF, LineOf<PROF>::LT, 1>(P, P.witems_ls, P.n_witems_ls, wraw);
    }
    // the last CTA out re-arms the work counters for the next launch
    PHASE_BEGIN
    PHASE_END
    PHASE_BEGIN
        if (tid == 0 && plba_atomic_fetch_add_i(&P.counters[CNT_WTICKET], 1) == PLBA_NB - 1) {
            P.counters[CNT_WORK_PT] = 0; P.counters[CNT_WORK_LS] = 0; P.counters[CNT_WTICKET] = 0;
        }
    PHASE_END
}

// ---------------------------------------------------------------------------------------------------------
// update: re-linearise, back-substitute x_l = H_ll^-1 (b_l - W^T x_p), retract, evaluate the new cost (subsystem 4)
// ---------------------------------------------------------------------------------------------------------
template <int PROF, int LT>
PLBA_D void update_items_w(const DevP &Pin, const WItem *items, int n_items, unsigned char *wraw) {
    PLBA_PARAMS_REF(P, Pin);
    typedef KT<PROF, LT> K;
    typedef ObsAcc<LT> OA;
    const int D = K::D, RANK = K::RANK;
    WSmem<PROF, LT> sm(wraw);
    double *U = sm.TA;                                           // [RANK][32]: J_pose x_p of each observation
    double *XL = sm.A, *PL = sm.A + 4 * 32;                       // orthonormal lines: [4][32] step and [6][32] new Plücker vector per landmark of the item
    (void)XL; (void)PL;                                          //   (the A rows stay in registers in this kernel: their shared-memory tile is free)
    LANE_VAR(double, chi_l); LANE_VAR(double, sc_l); LANE_VAR(double, d2_l);
    LANE_ARR(double, xpl, 6);
    LANE_VAR(int, kf_l); LANE_VAR(int, m_l); LANE_VAR(int, slot_l); LANE_VAR(int, act_l);
    int cur_win = -1;
    WPHASE_BEGIN
        LANE_BIND(chi_l); LANE_BIND(sc_l); LANE_BIND(d2_l);
        chi_l = 0.0; sc_l = 0.0; d2_l = 0.0;
    WPHASE_END
    for (;;) {
        // next work item: one atomic per warp (items differ in cost, a static deal leaves the slowest warp far behind)
        WPHASE_BEGIN
            if (lane == 0) sm.slot[48] = plba_atomic_fetch_add_i(&P.counters[LT == LT_POINT ? CNT_WORK_PT : CNT_WORK_LS], 1);
        WPHASE_END
        const int ii = sm.slot[48];
        WPHASE_BEGIN
        WPHASE_END
        if (ii >= n_items) break;
        const WItem it = items[ii];
        const WinCtrl ctl = P.ctrl[it.win];             // the controller only runs between launches: a private copy is safe
        if (ctl.done) continue;
        if (it.win != cur_win) {
            if (cur_win >= 0) {
                WPHASE_BEGIN
                    LANE_BIND(chi_l); LANE_BIND(sc_l); LANE_BIND(d2_l);
                    double *acc = P.accB + (size_t)4 * cur_win;
                    if (PROF == PLBA_PROFILE_G) PLBA_WARP_FLUSH_ADD(&acc[ACC_CHI_NEW], chi_l);
                    PLBA_WARP_FLUSH_ADD(&acc[ACC_SCALE], sc_l);
                    PLBA_WARP_FLUSH_ADD(&acc[ACC_DX2], d2_l);
                    chi_l = 0.0; sc_l = 0.0; d2_l = 0.0;
                WPHASE_END
            }
            cur_win = it.win;
        }
        // k == 0: landmarks without observations still go through the damped solve of their (zero) block, lane = landmark
        // (g2o: x_l = 0; the hand LM: 0/0, which the reference then adds to the state)
        const int k = it.k;
        const int lpp = k ? 32 / k : 32;
        const int npass = (it.n_lm + lpp - 1) / lpp;
        const double *state = OA::state(P, ctl.cur);
        double *state_new = OA::state(P, ctl.cur ^ 1);
        WPHASE_BEGIN
            LANE_BIND(kf_l); LANE_BIND(m_l); LANE_BIND(slot_l); LANE_BIND(xpl); LANE_BIND(act_l);
            m_l = k ? lane / k : lane;
            kf_l = k ? OA::kf(P)[it.ob0 + (lane - m_l * k)] : 0;
            slot_l = k ? P.kf_slot[kf_l] : -1;
            act_l = 0;
#pragma unroll
            for (int c = 0; c < 6; c++) xpl[c] = (slot_l >= 0) ? P.xp[(size_t)6 * slot_l + c] : 0.0;
        WPHASE_END
        for (int pass = 0; pass < npass; pass++) {
            const int lmb = pass * lpp;
            const int nlp = (it.n_lm - lmb < lpp) ? it.n_lm - lmb : lpp;
            const int nvalid = k ? nlp * k : nlp;
            WPHASE_BEGIN
                LANE_BIND(sc_l); LANE_BIND(kf_l); LANE_BIND(m_l); LANE_BIND(slot_l); LANE_BIND(xpl); LANE_BIND(act_l);
                if (k && lane < nvalid) {
                    const int o = it.ob0 + lmb * k + lane;
                    LmD<LT> lmd; ObsLoad<LT> ob;
                    ob.load(P, o);
                    LmLoad<PROF, LT>::run(P, ctl, it.win, it.lm0 + lmb + m_l, state, lmd);
                    double A[RANK * 6], B[RANK * D], e[RANK], cost;
                    act_l = obs_lin_w<PROF, LT>(P, ctl, o, kf_l, lmd, ob, A, B, e, cost) ? 1 : 0;
#pragma unroll
                    for (int kk = 0; kk < RANK; kk++) {
                        double u = 0.0;
#pragma unroll
                        for (int c = 0; c < 6; c++) u += A[kk * 6 + c] * xpl[c];
                        if (slot_l < 0) u = 0.0;
                        U[kk * 32 + lane] = u;
                        sc_l -= u * e[kk];                     // x_p^T b_p restricted to this edge
                    }
#pragma unroll
                    for (int i = 0; i < RANK * D; i++) sm.B[i * 32 + lane] = B[i];
#pragma unroll
                    for (int i = 0; i < RANK; i++) sm.E[i * 32 + lane] = e[i];
                }
            WPHASE_END
            WPHASE_BEGIN
                LANE_BIND(chi_l); LANE_BIND(sc_l); LANE_BIND(d2_l); LANE_BIND(kf_l); LANE_BIND(m_l); LANE_BIND(act_l);
                if (k && pass + 1 < npass && lane < lpp * k) {                 // the next pass's inputs travel while this pass's landmarks are solved
                    const int lmn = lmb + lpp + m_l;
                    if (lmn < it.n_lm) prefetch_obs_w<PROF, LT>(P, ctl, it.ob0 + (lmb + lpp) * k + lane, it.lm0 + lmn, state);
                }
                if (lane < nvalid) {
                    const int t0 = m_l * k;
                    const int lm = it.lm0 + lmb + m_l;
                    const bool leader = (k == 0) || (lane == t0);
                    double H[D * D], bl[D], rhs[D];
#pragma unroll
                    for (int i = 0; i < D * D; i++) H[i] = 0.0;
#pragma unroll
                    for (int i = 0; i < D; i++) { bl[i] = 0.0; rhs[i] = 0.0; }
                    for (int t = t0; t < t0 + k; t++) {
#pragma unroll
                        for (int kk = 0; kk < RANK; kk++) {
                            double b[D];
#pragma unroll
                            for (int c = 0; c < D; c++) b[c] = sm.B[(kk * D + c) * 32 + t];
                            const double ek = sm.E[kk * 32 + t], u = U[kk * 32 + t];
#pragma unroll
                            for (int r = 0; r < D; r++) {
                                bl[r] -= b[r] * ek; rhs[r] -= b[r] * u;
#pragma unroll
                                for (int c = r; c < D; c++) H[r * D + c] += b[r] * b[c];
                            }
                        }
                    }
#pragma unroll
                    for (int r = 0; r < D; r++) {
                        rhs[r] += bl[r];
#pragma unroll
                        for (int c = 0; c < r; c++) H[r * D + c] = H[c * D + r];
                    }
                    damp_invert<PROF, D>(ctl, H);
                    double xl[D];
#pragma unroll
                    for (int r = 0; r < D; r++) {
                        double s = 0;
#pragma unroll
                        for (int c = 0; c < D; c++) s += H[r * D + c] * rhs[c];
                        xl[r] = s;
                    }
                    if (leader) {
#pragma unroll
                        for (int r = 0; r < D; r++) { sc_l += xl[r] * (ctl.lambda * xl[r] + bl[r]); d2_l += xl[r] * xl[r]; }
                    }
                    if constexpr (LT == LT_LINE_ORTH) {
                        // orthonormal lines: the retraction and the new Plücker vector are trigonometry per LANDMARK; the step is
                        // parked in shared memory and retracted below with one lane per landmark (items hold <= 32 line landmarks)
                        if (leader) {
#pragma unroll
                            for (int i = 0; i < D; i++) XL[i * 32 + lmb + m_l] = xl[i];
                        }
                    } else {
                    // retraction
                    double cur[D], nw[D];
#pragma unroll
                    for (int i = 0; i < D; i++) cur[i] = state[(size_t)D * lm + i];
#pragma unroll
                    for (int i = 0; i < D; i++) nw[i] = cur[i] + xl[i];
                    if (leader && (PROF == PLBA_PROFILE_G || ctl.apply)) {
#pragma unroll
                        for (int i = 0; i < D; i++) state_new[(size_t)D * lm + i] = nw[i];
                    }
                    if constexpr (PROF == PLBA_PROFILE_G) if (act_l) {
                        // new cost at the trial state (computeActiveErrors + activeRobustChi2 after update)
                        const int o = it.ob0 + lmb * k + lane;
                        const double *T = P.poseT[ctl.cur ^ 1] + (size_t)12 * kf_l;
                        double e[2];
                        ObsLoad<LT> ob; ob.load(P, o);
                        double zc; g_point_error(P.cam, T, nw, ob.uv, e, zc);
                        const double chi2 = OA::om(P)[o] * (e[0] * e[0] + e[1] * e[1]);
                        OA::chi2(P)[o] = chi2;                       // the cached _error of the edge (e->chi2())
                        double rho0 = chi2, rho1;
                        if (ctl.stage == 0) huber(P.huber_delta, chi2, rho0, rho1);
                        chi_l += rho0;
                    }
                    }
                }
            WPHASE_END
        }
        if constexpr (LT == LT_LINE_ORTH) {
            // ---- lane = landmark: updateOrthCoord (a11), the new state, its Plücker vector / U / W for the next linearisation ----
            WPHASE_BEGIN
                if (lane < it.n_lm) {
                    const int lm = it.lm0 + lane;
                    double cur[4], xl[4], nw[4], pl[6];
#pragma unroll
                    for (int i = 0; i < 4; i++) { cur[i] = state[(size_t)4 * lm + i]; xl[i] = XL[i * 32 + lane]; }
                    orth_update(cur, xl, nw);
                    orth_to_plk_sc(nw, pl);
#pragma unroll
                    for (int i = 0; i < 6; i++) PL[i * 32 + lane] = pl[i];
                    if (PROF == PLBA_PROFILE_G || ctl.apply) {
#pragma unroll
                        for (int i = 0; i < 4; i++) state_new[(size_t)4 * lm + i] = nw[i];
                        LinePre Ln; line_pre_from_plk(pl, Ln);
                        double *c = P.lpre[ctl.cur ^ 1] + (size_t)LPRE_N * lm;
#pragma unroll
                        for (int i = 0; i < 3; i++) { c[i] = Ln.n[i]; c[3 + i] = Ln.d[i]; c[6 + i] = Ln.u1[i]; c[9 + i] = Ln.u2[i]; c[12 + i] = Ln.u3[i]; }
                        c[15] = Ln.w1; c[16] = Ln.w2;
                    }
                }
            WPHASE_END
            if constexpr (PROF == PLBA_PROFILE_G) {
                // ---- new cost at the trial state (computeActiveErrors + activeRobustChi2 after update), lane = observation again ----
                for (int pass = 0; pass < npass && k; pass++) {
                    const int lmb = pass * lpp;
                    const int nlp = (it.n_lm - lmb < lpp) ? it.n_lm - lmb : lpp;
                    WPHASE_BEGIN
                        LANE_BIND(chi_l); LANE_BIND(kf_l); LANE_BIND(m_l);
                        if (lane < nlp * k) {
                            const int o = it.ob0 + lmb * k + lane;
                            if (!(ctl.stage == 1 && OA::lvl(P)[o])) {          // the active set of obs_lin_w
                                const double *T = P.poseT[ctl.cur ^ 1] + (size_t)12 * kf_l;
                                const int l = lmb + m_l;
                                const double n3[3] = {PL[l], PL[32 + l], PL[64 + l]}, d3[3] = {PL[96 + l], PL[128 + l], PL[160 + l]};
                                double e[2];
                                ObsLoad<LT> ob; ob.load(P, o);
                                g_line_error(P.cam, T, n3, d3, ob.ab, e);
                                const double chi2 = OA::om(P)[o] * (e[0] * e[0] + e[1] * e[1]);
                                OA::chi2(P)[o] = chi2;
                                double rho0 = chi2, rho1;
                                if (ctl.stage == 0) huber(P.huber_delta, chi2, rho0, rho1);
                                chi_l += rho0;
                            }
                        }
                    WPHASE_END
                }
            }
        }
    }
    if (cur_win >= 0) {
        WPHASE_BEGIN
            LANE_BIND(chi_l); LANE_BIND(sc_l); LANE_BIND(d2_l);
            double *acc = P.accB + (size_t)4 * cur_win;
            if (PROF == PLBA_PROFILE_G) PLBA_WARP_FLUSH_ADD(&acc[ACC_CHI_NEW], chi_l);
            PLBA_WARP_FLUSH_ADD(&acc[ACC_SCALE], sc_l);
            PLBA_WARP_FLUSH_ADD(&acc[ACC_DX2], d2_l);
        WPHASE_END
    }
}

template <int PROF>
PLBA_KERNEL void PLBA_BOUNDS(WNT, PLBA_WU_CTAS) k_update_w(const DevP *Pp, int flags) {
    PLBA_SMEM(raw);
    PLBA_PARAMS(P, Pp);
    unsigned char *wraw = raw + (size_t)PLBA_WARP_IN_CTA * WSmemMax<PROF>::bytes();
    if (P.S_clear_doubles) clear_consumed_S(P, PLBA_BID, PLBA_NB);
    update_items_w<PROF, LT_POINT>(P, P.witems_pt, P.n_witems_pt, wraw);
    update_items_w<PROF, LineOf<PROF>::LT>(P, P.witems_ls, P.n_witems_ls, wraw);
    // the last CTA to arrive re-arms the work counters and (fused mode) runs the controller for every window
    int *last = (int *)raw;
    PHASE_BEGIN
    PHASE_END
    PHASE_BEGIN
        if (tid == 0) { plba_fence(); *last = (plba_atomic_fetch_add_i(&P.counters[CNT_TICKET], 1) == PLBA_NB - 1) ? 1 : 0; }
    PHASE_END
    if (!*last) return;
    PHASE_BEGIN
        if (tid == 0) { plba_fence(); P.counters[CNT_WORK_PT] = 0; P.counters[CNT_WORK_LS] = 0; }
        if (flags & KF_FUSE_CONTROL) for (int w = tid; w < P.n_win; w += PLBA_NT) control_window(P, w);
    PHASE_END
    PHASE_BEGIN
        if (tid == 0) { P.counters[CNT_TICKET] = 0; plba_fence(); if (flags & KF_FUSE_CONTROL) round_epilogue(P, flags); }
    PHASE_END
}

}  // namespace plba
