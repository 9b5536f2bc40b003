// Deterministic synthetic stereo scenes for the LBA path: keyframes on a smooth arc, point
// landmarks with contiguous tracks, 3-D line segments observed as endpoint pixels (Plücker mode) or as normalised
// 2-D lines (endpoint mode), pixel noise, gross outliers, perturbed initial state.  Same bits everywhere: the
// scene is consumed unchanged by the CUDA path, the CPU oracle and the Python tests.
//
// Field meanings follow the reference's map: T_kf_w is camera->world
// (src/mapHandler.cpp:299-300), Plücker vector = [n; d] with |d| = 1 at creation (:451-459), line observation =
// left-image endpoint pixels (:460-461) or sp x ep / sqrt(lx^2+ly^2) (src2/stereoFrame.cpp:359).
#include <cstdint>
#include <cmath>
#include <vector>
#include <algorithm>
#include <cstring>
#include "../../include/plba.h"

namespace {

struct Rng {   // SplitMix64-seeded xoshiro256**
    uint64_t s[4];
    static uint64_t splitmix(uint64_t &x) { uint64_t z = (x += 0x9e3779b97f4a7c15ULL); z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL; z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL; return z ^ (z >> 31); }
    explicit Rng(uint64_t seed) { for (int i = 0; i < 4; i++) s[i] = splitmix(seed); }
    static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
    uint64_t next() { uint64_t r = rotl(s[1] * 5, 7) * 9, t = s[1] << 17; s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45); return r; }
    double uni() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
    double uni(double a, double b) { return a + (b - a) * uni(); }
    int below(int n) { return (int)(uni() * n); }
    double normal() { double u1 = 1.0 - uni(), u2 = uni(); return std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2); }
    int poisson(double mean) { double L = std::exp(-mean), p = 1.0; int k = 0; do { k++; p *= uni(); } while (p > L); return k - 1; }
};

struct Pose { double R[9]; double t[3]; };   // camera->world

void rodrigues(const double w[3], double R[9]) {
    double th = std::sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
    double K[9] = {0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0};
    double a = th < 1e-12 ? 1.0 : std::sin(th) / th, b = th < 1e-12 ? 0.5 : (1 - std::cos(th)) / (th * th);
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) {
        double kk = 0; for (int k = 0; k < 3; k++) kk += K[r * 3 + k] * K[k * 3 + c];
        R[r * 3 + c] = (r == c ? 1.0 : 0.0) + a * K[r * 3 + c] + b * kk;
    }
}
// log of SE(3) in the reference's ordering [translation; rotation] (src2/auxiliar.cpp:143-173 semantics)
void log_se3(const Pose &T, double x[6]) {
    double c = (T.R[0] + T.R[4] + T.R[8] - 1.0) / 2.0; c = std::min(1.0, std::max(-1.0, c));
    double s = std::sqrt(1.0 - c * c), th = std::acos(c);
    double w[3] = {0, 0, 0}; double V[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    if (th > 1e-6) {
        double f = th / (2.0 * s);
        w[0] = f * (T.R[7] - T.R[5]); w[1] = f * (T.R[2] - T.R[6]); w[2] = f * (T.R[3] - T.R[1]);
        double K[9] = {0, -w[2] / th, w[1] / th, w[2] / th, 0, -w[0] / th, -w[1] / th, w[0] / th, 0};
        for (int r = 0; r < 3; r++) for (int cc = 0; cc < 3; cc++) {
            double kk = 0; for (int k = 0; k < 3; k++) kk += K[r * 3 + k] * K[k * 3 + cc];
            V[r * 3 + cc] = (r == cc ? 1.0 : 0.0) + K[r * 3 + cc] * (1.0 - c) / th + kk * (th - s) / th;
        }
    }
    // t = V^-1 * T.t  (3x3 adjugate inverse)
    double det = V[0] * (V[4] * V[8] - V[5] * V[7]) - V[1] * (V[3] * V[8] - V[5] * V[6]) + V[2] * (V[3] * V[7] - V[4] * V[6]);
    double inv[9] = {(V[4] * V[8] - V[5] * V[7]) / det, (V[2] * V[7] - V[1] * V[8]) / det, (V[1] * V[5] - V[2] * V[4]) / det,
                     (V[5] * V[6] - V[3] * V[8]) / det, (V[0] * V[8] - V[2] * V[6]) / det, (V[2] * V[3] - V[0] * V[5]) / det,
                     (V[3] * V[7] - V[4] * V[6]) / det, (V[1] * V[6] - V[0] * V[7]) / det, (V[0] * V[4] - V[1] * V[3]) / det};
    for (int r = 0; r < 3; r++) x[r] = inv[r * 3] * T.t[0] + inv[r * 3 + 1] * T.t[1] + inv[r * 3 + 2] * T.t[2];
    x[3] = w[0]; x[4] = w[1]; x[5] = w[2];
}
void to_cam(const Pose &T, const double Xw[3], double Xc[3]) {   // Xc = R^T (Xw - t)
    double d[3] = {Xw[0] - T.t[0], Xw[1] - T.t[1], Xw[2] - T.t[2]};
    for (int c = 0; c < 3; c++) Xc[c] = T.R[c] * d[0] + T.R[3 + c] * d[1] + T.R[6 + c] * d[2];
}
void to_world(const Pose &T, const double Xc[3], double Xw[3]) {
    for (int r = 0; r < 3; r++) Xw[r] = T.R[r * 3] * Xc[0] + T.R[r * 3 + 1] * Xc[1] + T.R[r * 3 + 2] * Xc[2] + T.t[r];
}

}  // namespace

struct plba_scene_s {
    plba_scene_spec spec;
    int n_kf = 0;
    std::vector<double> kf_T, kf_T_true, x_pose, pt, pt_true, plk, plk_true, ends;
    std::vector<int32_t> kf_slot, po_lm, po_kf, lo_lm, lo_kf;
    std::vector<double> po_uv, lo_ab;
};

extern "C" {

void plba_scene_preset(int32_t config_id, plba_scene_spec *s) {
    std::memset(s, 0, sizeof(*s));
    // EuRoC-shaped defaults
    s->width = 752; s->height = 480; s->fx = s->fy = 435.2; s->cx = 367.2; s->cy = 252.2;
    s->kf_spacing = 0.3; s->depth_min = 2; s->depth_max = 12;
    s->pixel_noise = 1.0; s->outlier_frac = 0.02;
    s->pose_rot_noise = 0.01; s->pose_trans_noise = 0.02; s->pt_noise = 0.05; s->ls_noise = 0.01;
    s->seed = 20261018ULL + (uint64_t)config_id; s->line_mode = 0; s->loop_every = 0;
    switch (config_id) {
    default:
    case 1: s->n_kf_free = 10; s->n_kf_fixed = 2; s->n_pt = 2000; s->n_ls = 500; s->mean_track = 4; break;   // also the window shape of config 3
    case 2:
        s->n_kf_free = 20; s->n_kf_fixed = 2; s->n_pt = 8000; s->n_ls = 2000; s->mean_track = 5;
        s->width = 1241; s->height = 376; s->fx = s->fy = 718.856; s->cx = 607.1928; s->cy = 185.2157;
        s->kf_spacing = 1.5; s->depth_min = 5; s->depth_max = 50; break;
    case 3: s->n_kf_free = 10; s->n_kf_fixed = 2; s->n_pt = 2000; s->n_ls = 500; s->mean_track = 4; break;
    case 4: s->n_kf_free = 200; s->n_kf_fixed = 1; s->n_pt = 200000; s->n_ls = 50000; s->mean_track = 5; break;
    case 5:
        s->n_kf_free = 2000; s->n_kf_fixed = 1; s->n_pt = 2000000; s->n_ls = 500000; s->mean_track = 5;
        s->width = 1241; s->height = 376; s->fx = s->fy = 718.856; s->cx = 607.1928; s->cy = 185.2157;
        s->kf_spacing = 1.5; s->depth_min = 5; s->depth_max = 50; break;
    }
}

int plba_scene_create(const plba_scene_spec *spec, plba_scene *out) {
    if (!spec || !out || spec->n_kf_free < 0 || spec->n_kf_fixed < 1 || spec->n_pt < 0 || spec->n_ls < 0) return PLBA_E_ARG;
    plba_scene_s *S = new plba_scene_s();
    S->spec = *spec;
    const plba_scene_spec &sp = S->spec;
    Rng rng(sp.seed);
    const int nk = sp.n_kf_free + sp.n_kf_fixed;
    S->n_kf = nk;
    // ---- trajectory: camera z forward, yaw 2 deg / KF about the y axis ----
    std::vector<Pose> Ttrue(nk), Tinit(nk);
    double px = 0, pz = 0;
    const double yaw_step = 2.0 * 3.14159265358979323846 / 180.0;
    for (int i = 0; i < nk; i++) {
        double yaw = yaw_step * i;
        double w[3] = {0, yaw, 0};
        rodrigues(w, Ttrue[i].R);
        Ttrue[i].t[0] = px; Ttrue[i].t[1] = 0.02 * std::sin(0.7 * i); Ttrue[i].t[2] = pz;
        px += sp.kf_spacing * std::sin(yaw); pz += sp.kf_spacing * std::cos(yaw);
    }
    S->kf_slot.resize(nk);
    for (int i = 0; i < nk; i++) S->kf_slot[i] = (i < sp.n_kf_fixed) ? -1 : i - sp.n_kf_fixed;   // oldest KFs are the fixed observers
    for (int i = 0; i < nk; i++) {
        Tinit[i] = Ttrue[i];
        if (S->kf_slot[i] >= 0) {
            double w[3] = {sp.pose_rot_noise * rng.normal(), sp.pose_rot_noise * rng.normal(), sp.pose_rot_noise * rng.normal()};
            double dR[9]; rodrigues(w, dR);
            for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) { double s = 0; for (int k = 0; k < 3; k++) s += Ttrue[i].R[r * 3 + k] * dR[k * 3 + c]; Tinit[i].R[r * 3 + c] = s; }
            for (int r = 0; r < 3; r++) Tinit[i].t[r] += sp.pose_trans_noise * rng.normal();
        }
    }
    auto rows = [&](const std::vector<Pose> &T, std::vector<double> &o) {
        o.resize((size_t)nk * 12);
        for (int i = 0; i < nk; i++) for (int r = 0; r < 3; r++) { for (int c = 0; c < 3; c++) o[(size_t)i * 12 + r * 4 + c] = T[i].R[r * 3 + c]; o[(size_t)i * 12 + r * 4 + 3] = T[i].t[r]; }
    };
    rows(Tinit, S->kf_T); rows(Ttrue, S->kf_T_true);
    S->x_pose.resize((size_t)sp.n_kf_free * 6);
    for (int i = 0; i < nk; i++) if (S->kf_slot[i] >= 0) log_se3(Tinit[i], &S->x_pose[(size_t)S->kf_slot[i] * 6]);

    auto in_image = [&](const double Xc[3], double uv[2]) {
        if (Xc[2] <= 0.1) return false;
        uv[0] = sp.cx + sp.fx * Xc[0] / Xc[2]; uv[1] = sp.cy + sp.fy * Xc[1] / Xc[2];
        return uv[0] >= 0 && uv[0] < sp.width && uv[1] >= 0 && uv[1] < sp.height;
    };
    auto track = [&](int home, std::vector<int> &kfs) {   // contiguous run around the home KF (+ optional revisit)
        int L = std::min(nk, std::max(2, rng.poisson(std::max(0.0, sp.mean_track - 2.0)) + 2));
        int start = home - rng.below(L);
        start = std::max(0, std::min(start, nk - L));
        kfs.clear();
        for (int k = start; k < start + L; k++) kfs.push_back(k);
        if (sp.loop_every > 0) for (int k = start + sp.loop_every; k < start + sp.loop_every + 2 && k < nk; k++) kfs.push_back(k);
    };
    auto noisy = [&](double v) { v += sp.pixel_noise * rng.normal(); if (rng.uni() < sp.outlier_frac) v += rng.uni(-20, 20); return v; };

    // ---- points ----
    S->pt.reserve((size_t)sp.n_pt * 3); S->pt_true.reserve((size_t)sp.n_pt * 3);
    std::vector<int> kfs; std::vector<int> vis; std::vector<double> uvs;
    for (int l = 0; l < sp.n_pt; l++) {
        double Xw[3];
        for (int tries = 0;; tries++) {
            int home = rng.below(nk);
            double z = rng.uni(sp.depth_min, sp.depth_max), u = rng.uni(0, sp.width), v = rng.uni(0, sp.height);
            double Xc[3] = {(u - sp.cx) / sp.fx * z, (v - sp.cy) / sp.fy * z, z};
            to_world(Ttrue[home], Xc, Xw);
            track(home, kfs);
            vis.clear(); uvs.clear();
            for (int k : kfs) { double Pc[3], uv[2]; to_cam(Ttrue[k], Xw, Pc); if (in_image(Pc, uv)) { vis.push_back(k); uvs.push_back(uv[0]); uvs.push_back(uv[1]); } }
            if (vis.size() >= 2 || tries > 50) break;
        }
        for (size_t j = 0; j < vis.size(); j++) {
            S->po_lm.push_back(l); S->po_kf.push_back(vis[j]);
            S->po_uv.push_back(noisy(uvs[2 * j])); S->po_uv.push_back(noisy(uvs[2 * j + 1]));
        }
        for (int a = 0; a < 3; a++) { S->pt_true.push_back(Xw[a]); S->pt.push_back(Xw[a] + sp.pt_noise * rng.normal()); }
    }
    // ---- lines ----
    for (int l = 0; l < sp.n_ls; l++) {
        double Pw[3], Qw[3];
        for (int tries = 0;; tries++) {
            int home = rng.below(nk);
            double z = rng.uni(sp.depth_min, sp.depth_max), u = rng.uni(0, sp.width), v = rng.uni(0, sp.height);
            double Mc[3] = {(u - sp.cx) / sp.fx * z, (v - sp.cy) / sp.fy * z, z}, Mw[3];
            to_world(Ttrue[home], Mc, Mw);
            double dir[3] = {rng.normal(), rng.normal(), rng.normal()};
            double dn = std::sqrt(dir[0] * dir[0] + dir[1] * dir[1] + dir[2] * dir[2]) + 1e-12;
            double len = rng.uni(0.5, 2.0);
            for (int a = 0; a < 3; a++) { Pw[a] = Mw[a] - 0.5 * len * dir[a] / dn; Qw[a] = Mw[a] + 0.5 * len * dir[a] / dn; }
            track(home, kfs);
            vis.clear(); uvs.clear();
            for (int k : kfs) {
                double Pc[3], Qc[3], Mcc[3], a[2], b[2], m[2];
                to_cam(Ttrue[k], Pw, Pc); to_cam(Ttrue[k], Qw, Qc); to_cam(Ttrue[k], Mw, Mcc);
                if (!in_image(Mcc, m) || Pc[2] <= 0.1 || Qc[2] <= 0.1) continue;
                a[0] = sp.cx + sp.fx * Pc[0] / Pc[2]; a[1] = sp.cy + sp.fy * Pc[1] / Pc[2];
                b[0] = sp.cx + sp.fx * Qc[0] / Qc[2]; b[1] = sp.cy + sp.fy * Qc[1] / Qc[2];
                if (std::hypot(a[0] - b[0], a[1] - b[1]) < 4.0) continue;   // degenerate image segment
                vis.push_back(k); uvs.push_back(a[0]); uvs.push_back(a[1]); uvs.push_back(b[0]); uvs.push_back(b[1]);
            }
            if (vis.size() >= 2 || tries > 50) break;
        }
        for (size_t j = 0; j < vis.size(); j++) {
            double a0 = noisy(uvs[4 * j]), b0 = noisy(uvs[4 * j + 1]), a1 = noisy(uvs[4 * j + 2]), b1 = noisy(uvs[4 * j + 3]);
            S->lo_lm.push_back(l); S->lo_kf.push_back(vis[j]);
            if (sp.line_mode == 0) { S->lo_ab.push_back(a0); S->lo_ab.push_back(b0); S->lo_ab.push_back(a1); S->lo_ab.push_back(b1); }
            else {   // le = sp x ep, normalised by sqrt(lx^2 + ly^2)
                double lx = b0 - b1, ly = a1 - a0, lz = a0 * b1 - a1 * b0, nn = std::sqrt(lx * lx + ly * ly) + 1e-300;
                S->lo_ab.push_back(lx / nn); S->lo_ab.push_back(ly / nn); S->lo_ab.push_back(lz / nn); S->lo_ab.push_back(0.0);
            }
        }
        // truth Plücker (n = p x d, |d| = 1) and the perturbed initial value
        double d[3] = {Qw[0] - Pw[0], Qw[1] - Pw[1], Qw[2] - Pw[2]};
        double dn = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
        for (int a = 0; a < 3; a++) d[a] /= dn;
        double n[3] = {Pw[1] * d[2] - Pw[2] * d[1], Pw[2] * d[0] - Pw[0] * d[2], Pw[0] * d[1] - Pw[1] * d[0]};
        for (int a = 0; a < 3; a++) S->plk_true.push_back(n[a]);
        for (int a = 0; a < 3; a++) S->plk_true.push_back(d[a]);
        double Pi[3], Qi[3];
        for (int a = 0; a < 3; a++) { Pi[a] = Pw[a] + sp.pt_noise * rng.normal() * (sp.line_mode ? 1.0 : sp.ls_noise / std::max(1e-12, sp.pt_noise)); Qi[a] = Qw[a] + sp.pt_noise * rng.normal() * (sp.line_mode ? 1.0 : sp.ls_noise / std::max(1e-12, sp.pt_noise)); }
        for (int a = 0; a < 3; a++) S->ends.push_back(Pi[a]);
        for (int a = 0; a < 3; a++) S->ends.push_back(Qi[a]);
        double di[3] = {Qi[0] - Pi[0], Qi[1] - Pi[1], Qi[2] - Pi[2]};
        double din = std::sqrt(di[0] * di[0] + di[1] * di[1] + di[2] * di[2]);
        for (int a = 0; a < 3; a++) di[a] /= din;
        double ni[3] = {Pi[1] * di[2] - Pi[2] * di[1], Pi[2] * di[0] - Pi[0] * di[2], Pi[0] * di[1] - Pi[1] * di[0]};
        for (int a = 0; a < 3; a++) S->plk.push_back(ni[a]);
        for (int a = 0; a < 3; a++) S->plk.push_back(di[a]);
    }
    *out = S;
    return PLBA_OK;
}

int plba_scene_problem(plba_scene S, plba_problem *p) {
    if (!S || !p) return PLBA_E_ARG;
    std::memset(p, 0, sizeof(*p));
    p->n_kf = S->n_kf; p->n_free = S->spec.n_kf_free; p->n_pt = S->spec.n_pt; p->n_ls = S->spec.n_ls;
    p->n_pobs = (int32_t)S->po_lm.size(); p->n_lobs = (int32_t)S->lo_lm.size();
    p->cam[0] = S->spec.fx; p->cam[1] = S->spec.fy; p->cam[2] = S->spec.cx; p->cam[3] = S->spec.cy;
    p->kf_T_wc = S->kf_T.data(); p->kf_slot = S->kf_slot.data(); p->x_pose = S->x_pose.data();
    p->pt_xyz = S->pt.data(); p->ls_plk = S->plk.data(); p->ls_end = S->ends.data();
    p->po_lm = S->po_lm.data(); p->po_kf = S->po_kf.data(); p->po_uv = S->po_uv.data(); p->po_sig2 = nullptr;
    p->lo_lm = S->lo_lm.data(); p->lo_kf = S->lo_kf.data(); p->lo_ab = S->lo_ab.data(); p->lo_sig2 = nullptr;
    return PLBA_OK;
}

int plba_scene_truth(plba_scene S, const double **kf, const double **pt, const double **plk) {
    if (!S) return PLBA_E_ARG;
    if (kf) *kf = S->kf_T_true.data();
    if (pt) *pt = S->pt_true.data();
    if (plk) *plk = S->plk_true.data();
    return PLBA_OK;
}

void plba_scene_destroy(plba_scene S) { delete S; }

}  // extern "C"
