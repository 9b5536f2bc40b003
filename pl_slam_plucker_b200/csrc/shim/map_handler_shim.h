// Host shim with the reference's own LBA entry points (PLSLAM::MapHandler, include/mapHandler.h:127-134) on top of the
// C ABI (include/plba.h).  Header-only C++; no CUDA, Eigen, OpenCV, g2o or DBoW2 needed.
//
// The real mapHandler.h cannot be compiled without OpenCV / DBoW2 / g2o / Eigen, so the map types below are minimal
// stand-ins that keep the reference's member NAMES and meaning (include/keyFrame.h:47-79, include/mapFeatures.h:41-122):
// swapping in the real headers is a matter of replacing Vec/Mat4 accessors by Eigen's.
//
//   int  localBundleAdjustment()                               src/mapHandler.cpp:1392-1502
//   int  levMarquardtOptimizationLBA(X_aux, kf_list, ...)      src/mapHandler.cpp:2334-3016   -> PLBA_PROFILE_H_END
//   int  localBundleAdjustmentForPluker()                      src/mapHandler.cpp:1505-1615
//   int  levMarquardtOptimizationLBAForPluker(...)             src/mapHandler.cpp:1618-2332   -> PLBA_PROFILE_H_PLK
//   void localBundleAdjustmentForPlukerWithG2O()               src/mapHandler.cpp:5851-6323   -> PLBA_PROFILE_G
//   void globalBundleAdjustment()                              src/mapHandler.cpp:3022-3126
//   void levMarquardtOptimizationGBA(X_aux, kf_list, ...)      src/mapHandler.cpp:3128-3728   -> PLBA_PROFILE_H_END + PLBA_SHELL_GBA
//   void removeBadMapLandmarksForPluker()                      src/mapHandler.cpp:3816-3897   (host bookkeeping right after the LBA)
//
// Only index bookkeeping (flattening, write-back, observation erasure) happens here; every FP64 operation of the path
// runs in the CUDA library.  If the library reports an error the shim throws: there is no CPU fallback.
#pragma once
#include <array>
#include <cmath>
#include <map>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>
#include "../../../include/plba.h"

namespace plba_shim {

typedef std::array<int, 6> Vector6i;          // (lm map id, lm local idx, obs idx, kf map id, kf local idx | -1, inlier), include/mapHandler.h:63
typedef std::array<double, 16> Matrix4d;      // row-major 4x4
typedef std::array<double, 6> Vector6d;
typedef std::array<double, 4> Vector4d;
typedef std::array<double, 3> Vector3d;
typedef std::array<double, 2> Vector2d;

struct KeyFrame {      // include/keyFrame.h:47-79
    bool local = false;
    int kf_idx = 0;
    Matrix4d T_kf_w{};                          // camera -> world
    Vector6d x_kf_w{};                          // logmap_se3(T_kf_w), translation first (never refreshed by LBA: Q11)
    std::vector<int> stereo_pt_idx, stereo_ls_idx;   // stereo_frame->stereo_pt[i]->idx / stereo_ls[i]->idx: map landmark of each stereo feature, -1 = none
};
struct MapPoint {      // include/mapFeatures.h:41-70
    int idx = 0; bool inlier = true, local = true;
    Vector3d point3D{};
    std::vector<Vector2d> obs_list; std::vector<int> kf_obs_list; std::vector<double> sigma_list;
    void addMapPointObservation(int kf_obs, Vector2d obs, double sigma2 = 1.0) { obs_list.push_back(obs); kf_obs_list.push_back(kf_obs); sigma_list.push_back(sigma2); }
};
struct MapLine {       // include/mapFeatures.h:72-122
    int idx = 0; bool inlier = true, local = true;
    Vector6d line3D{};                          // endpoint mode: [P; Q]
    Vector6d NDw{};                             // Plücker mode: [n; d]
    std::vector<Vector3d> obs_list;             // endpoint mode: normalised 2-D line
    std::vector<Vector4d> NDw_obs_list;         // Plücker mode: endpoint pixels (spl.x, spl.y, epl.x, epl.y)
    std::vector<int> kf_obs_list; std::vector<double> sigma_list;
};
struct PinholeStereoCamera { double fx, fy, cx, cy; double getFx() const { return fx; } double getFy() const { return fy; } double getCx() const { return cx; } double getCy() const { return cy; } };
struct SlamConfig {    // src/slamConfig.cpp:65-67, src2/config.cpp:80-85
    double lambda_lba_lm = 1e-5, lambda_lba_k = 10.0; int max_iters_lba = 15;
    double homog_th = 1e-7, min_error = 1e-7, min_error_change = 1e-7;
    int min_lm_obs = 5;                         // src/slamConfig.cpp:48
};
enum { VO_PROCESSING = 0, VO_INSERTING_KF = 1 };

class MapHandlerShim {
public:
    std::vector<KeyFrame *> map_keyframes;      // nullptr slots allowed (culled), as in the reference
    std::vector<MapPoint *> map_points;
    std::vector<MapLine *> map_lines;
    PinholeStereoCamera *cam = nullptr;
    SlamConfig config;
    unsigned int max_kf_idx = 0;                // include/mapHandler.h:156
    std::map<int, std::vector<int>> map_points_kf_idx, map_lines_kf_idx;   // base KF -> landmarks first seen there (include/mapHandler.h:148-149)
    int vo_status = VO_PROCESSING;              // never assigned in the reference (Q17)
    int quirks = PLBA_QUIRKS_FAITHFUL;
    int n_bad_point_obs = 0, n_bad_line_obs = 0;
    std::vector<plba_trace_rec> last_trace;

    explicit MapHandlerShim(int device = 0) { if (plba_create(device, nullptr, &h_) != PLBA_OK) throw std::runtime_error("plba_create failed: a CUDA device is required (no CPU fallback)"); }
    ~MapHandlerShim() { plba_destroy(h_); }
    MapHandlerShim(const MapHandlerShim &) = delete;

    // ---- drivers: flatten the local map exactly as the reference does (src/mapHandler.cpp:1396-1493 / :1509-1607) ----
    int localBundleAdjustment() { Gathered g = gather(false); if (g.pt_obs.size() + g.ls_obs.size() == 0) return -1; return levMarquardtOptimizationLBA(g.X, g.kfs, g.pts, g.lss, g.pt_obs, g.ls_obs); }
    int localBundleAdjustmentForPluker() { Gathered g = gather(true); if (g.pt_obs.size() + g.ls_obs.size() == 0) return -1; return levMarquardtOptimizationLBAForPluker(g.X, g.kfs, g.pts, g.lss, g.pt_obs, g.ls_obs); }

    int levMarquardtOptimizationLBA(std::vector<double> X_aux, std::vector<int> kf_list, std::vector<int> pt_list, std::vector<int> ls_list,
                                    std::vector<Vector6i> pt_obs_list, std::vector<Vector6i> ls_obs_list) {
        return hand_lm(PLBA_PROFILE_H_END, X_aux, kf_list, pt_list, ls_list, pt_obs_list, ls_obs_list);
    }
    int levMarquardtOptimizationLBAForPluker(std::vector<double> X_aux, std::vector<int> kf_list, std::vector<int> pt_list, std::vector<int> ls_list,
                                             std::vector<Vector6i> pt_obs_list, std::vector<Vector6i> ls_obs_list) {
        return hand_lm(PLBA_PROFILE_H_PLK, X_aux, kf_list, pt_list, ls_list, pt_obs_list, ls_obs_list);
    }

    // ---- global BA, run once at shutdown (app/plslam_dataset.cpp:174): every non-NULL KF but KF 0, every non-NULL landmark ----
    void globalBundleAdjustment() { Gathered g = gather(false, true); levMarquardtOptimizationGBA(g.X, g.kfs, g.pts, g.lss, g.pt_obs, g.ls_obs); }   // no emptiness test (:3120)
    void levMarquardtOptimizationGBA(std::vector<double> X_aux, std::vector<int> kf_list, std::vector<int> pt_list, std::vector<int> ls_list,
                                     std::vector<Vector6i> pt_obs_list, std::vector<Vector6i> ls_obs_list) {
        if (pt_obs_list.size() + ls_obs_list.size() == 0) return;      // the reference would factor an empty system; nothing to write back
        hand_lm(PLBA_PROFILE_H_END, X_aux, kf_list, pt_list, ls_list, pt_obs_list, ls_obs_list, PLBA_SHELL_GBA);
    }

    // ---- culling right after the LBA (localMappingThread, src/mapHandler.cpp:1278-1279).  Landmarks are owned by the caller here, so a
    //      culled landmark's slot is set to nullptr and the pointer is handed back instead of being deleted (:3851, :3892) ----
    std::vector<void *> removeBadMapLandmarksForPluker() {
        std::vector<void *> culled;
        cull(map_points, map_points_kf_idx, true, culled);
        cull(map_lines, map_lines_kf_idx, false, culled);
        return culled;
    }

    // ---- the Plücker-mode LBA that actually runs in the reference (src/mapHandler.cpp:5851-6323) ----
    void localBundleAdjustmentForPlukerWithG2O() {
        std::vector<MapPoint *> lpt; std::vector<MapLine *> lls;
        for (MapPoint *p : map_points) if (p && p->local) lpt.push_back(p);
        for (MapLine *l : map_lines) if (l && l->local) lls.push_back(l);
        std::map<int, KeyFrame *> nofix, fix;                                               // :5870-5919
        for (KeyFrame *kf : map_keyframes) if (kf && kf->local) nofix[kf->kf_idx] = kf;
        auto promote = [&](const std::vector<int> &kfs) {
            for (int k : kfs) {
                KeyFrame *kf = map_keyframes.at(k);
                if (!kf || kf->kf_idx != k) throw std::runtime_error("[Wrong index in the map_keyframes and landmark obs.....]");   // exit(0) in the reference (:5891-5894)
                if (!kf->local) { fix[k] = kf; kf->local = true; }
            }
        };
        for (MapPoint *p : lpt) promote(p->kf_obs_list);
        for (MapLine *l : lls) promote(l->kf_obs_list);
        std::vector<int> ids; std::unordered_map<int, int> row_of;
        { std::map<int, KeyFrame *> all = nofix; all.insert(fix.begin(), fix.end()); for (auto &kv : all) { row_of[kv.first] = (int)ids.size(); ids.push_back(kv.first); } }
        Flat f; f.cam = cam;
        int ns = 0;
        for (int k : ids) {
            const bool free_kf = nofix.count(k) && k != 0;                                   // KF 0 fixed (:5943-5945)
            f.kf_slot.push_back(free_kf ? ns++ : -1);
            push_T(f.kf_T, (nofix.count(k) ? nofix[k] : fix[k])->T_kf_w);
        }
        for (size_t loc = 0; loc < lpt.size(); loc++) {
            for (int i = 0; i < 3; i++) f.pt_xyz.push_back(lpt[loc]->point3D[i]);
            for (size_t i = 0; i < lpt[loc]->kf_obs_list.size(); i++) {
                f.po_lm.push_back((int)loc); f.po_kf.push_back(row_of.at(lpt[loc]->kf_obs_list[i]));
                f.po_uv.push_back(lpt[loc]->obs_list[i][0]); f.po_uv.push_back(lpt[loc]->obs_list[i][1]); f.po_sig2.push_back(lpt[loc]->sigma_list[i]);
            }
        }
        for (size_t loc = 0; loc < lls.size(); loc++) {
            for (int i = 0; i < 6; i++) f.ls_plk.push_back(lls[loc]->NDw[i]);
            for (size_t i = 0; i < lls[loc]->kf_obs_list.size(); i++) {
                f.lo_lm.push_back((int)loc); f.lo_kf.push_back(row_of.at(lls[loc]->kf_obs_list[i]));
                for (int k = 0; k < 4; k++) f.lo_ab.push_back(lls[loc]->NDw_obs_list[i][k]);
                f.lo_sig2.push_back(lls[loc]->sigma_list[i]);
            }
        }
        if (f.po_lm.size() + f.lo_lm.size() == 0) return;
        Out o = run(f, PLBA_PROFILE_G, (int)lpt.size(), (int)lls.size());
        // bad observations, newest first so that erase positions stay valid (:6156-6293)
        n_bad_point_obs = erase_bad(lpt, f.po_lm, o.po_flags, true);
        n_bad_line_obs = erase_bad(lls, f.lo_lm, o.lo_flags, false);
        for (size_t j = 0; j < ids.size(); j++) if (nofix.count(ids[j])) pull_T(o.kf_T, j, nofix[ids[j]]->T_kf_w);      // :6297-6303
        for (size_t loc = 0; loc < lpt.size(); loc++) for (int i = 0; i < 3; i++) lpt[loc]->point3D[i] = o.pt_xyz[3 * loc + i];       // :6306-6311
        for (size_t loc = 0; loc < lls.size(); loc++) for (int i = 0; i < 6; i++) lls[loc]->NDw[i] = o.ls_plk[6 * loc + i];           // :6314-6319
    }

private:
    plba_handle h_ = nullptr;
    struct Gathered { std::vector<double> X; std::vector<int> kfs, pts, lss; std::vector<Vector6i> pt_obs, ls_obs; };
    struct Flat {
        PinholeStereoCamera *cam = nullptr;
        std::vector<double> kf_T, x_pose, pt_xyz, ls_plk, ls_end, po_uv, po_sig2, lo_ab, lo_sig2;
        std::vector<int32_t> kf_slot, po_lm, po_kf, lo_lm, lo_kf;
    };
    struct Out { std::vector<double> kf_T, pt_xyz, ls_plk, ls_end; std::vector<uint8_t> pt_inlier, ls_inlier, po_flags, lo_flags; int rc = 0; };

    static void push_T(std::vector<double> &v, const Matrix4d &T) { for (int i = 0; i < 12; i++) v.push_back(T[i]); }
    static void pull_T(const std::vector<double> &v, size_t row, Matrix4d &T) { for (int i = 0; i < 12; i++) T[i] = v[12 * row + i]; T[12] = T[13] = T[14] = 0; T[15] = 1; }

    static Vector4d pluker_to_orth(const Vector6d &pl) {   // MapLine::changePlukerToOrth, src/mapFeatures.cpp:186-201 (fills X_aux only)
        const double nn = std::sqrt(pl[0] * pl[0] + pl[1] * pl[1] + pl[2] * pl[2]), dn = std::sqrt(pl[3] * pl[3] + pl[4] * pl[4] + pl[5] * pl[5]);
        const double u1[3] = {pl[0] / nn, pl[1] / nn, pl[2] / nn}, u2[3] = {pl[3] / dn, pl[4] / dn, pl[5] / dn};
        const double c[3] = {pl[1] * pl[5] - pl[2] * pl[4], pl[2] * pl[3] - pl[0] * pl[5], pl[0] * pl[4] - pl[1] * pl[3]};
        const double cn = std::sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]);
        return {std::atan2(u2[2], c[2] / cn), std::asin(-u1[2]), std::atan2(u1[1], u1[0]), std::asin(dn / std::sqrt(nn * nn + dn * dn))};
    }

    Gathered gather(bool pluker, bool global = false) {
        Gathered g;
        for (KeyFrame *kf : map_keyframes) if (kf && (global || kf->local) && kf->kf_idx != 0) { for (double v : kf->x_kf_w) g.X.push_back(v); g.kfs.push_back(kf->kf_idx); }
        std::unordered_map<int, int> pos;                                                    // replaces the O(Nobs*Nkf) search (:1437-1444)
        for (size_t j = 0; j < g.kfs.size(); j++) pos[g.kfs[j]] = (int)j;
        int loc = 0;
        for (MapPoint *p : map_points) if (p && (global || p->local)) {
            for (double v : p->point3D) g.X.push_back(v);
            for (size_t i = 0; i < p->obs_list.size(); i++) { const int k = p->kf_obs_list[i]; g.pt_obs.push_back({p->idx, loc, (int)i, k, pos.count(k) ? pos[k] : -1, 1}); }
            g.pts.push_back(p->idx); loc++;
        }
        loc = 0;
        for (MapLine *l : map_lines) if (l && (global || l->local)) {
            size_t n_obs;
            if (pluker) { const Vector4d o = pluker_to_orth(l->NDw); for (double v : o) g.X.push_back(v); n_obs = (quirks == PLBA_QUIRKS_FAITHFUL) ? l->obs_list.size() : l->NDw_obs_list.size(); }   // Q10 (:1582)
            else { for (double v : l->line3D) g.X.push_back(v); n_obs = l->obs_list.size(); }
            for (size_t i = 0; i < n_obs; i++) { const int k = l->kf_obs_list[i]; g.ls_obs.push_back({l->idx, loc, (int)i, k, pos.count(k) ? pos[k] : -1, 1}); }
            g.lss.push_back(l->idx); loc++;
        }
        return g;
    }

    Out run(Flat &f, int profile, int n_pt, int n_ls, int shell = PLBA_SHELL_LBA) {
        plba_problem P{};
        P.n_kf = (int)f.kf_slot.size(); P.n_free = 0; for (int s : f.kf_slot) if (s >= 0) P.n_free++;
        P.n_pt = n_pt; P.n_ls = n_ls; P.n_pobs = (int)f.po_lm.size(); P.n_lobs = (int)f.lo_lm.size();
        P.cam[0] = f.cam->getFx(); P.cam[1] = f.cam->getFy(); P.cam[2] = f.cam->getCx(); P.cam[3] = f.cam->getCy();
        P.kf_T_wc = f.kf_T.data(); P.kf_slot = f.kf_slot.data(); P.x_pose = f.x_pose.empty() ? nullptr : f.x_pose.data();
        P.pt_xyz = f.pt_xyz.data(); P.ls_plk = f.ls_plk.empty() ? nullptr : f.ls_plk.data(); P.ls_end = f.ls_end.empty() ? nullptr : f.ls_end.data();
        P.po_lm = f.po_lm.data(); P.po_kf = f.po_kf.data(); P.po_uv = f.po_uv.data(); P.po_sig2 = f.po_sig2.empty() ? nullptr : f.po_sig2.data();
        P.lo_lm = f.lo_lm.data(); P.lo_kf = f.lo_kf.data(); P.lo_ab = f.lo_ab.data(); P.lo_sig2 = f.lo_sig2.empty() ? nullptr : f.lo_sig2.data();
        plba_options opt; plba_default_options(profile, &opt);
        opt.quirks = quirks; opt.shell = shell; opt.lambda_lba_lm = config.lambda_lba_lm; opt.lambda_lba_k = config.lambda_lba_k; opt.max_iters_lba = config.max_iters_lba;
        opt.homog_th = config.homog_th; opt.min_error = config.min_error; opt.min_error_change = config.min_error_change;
        Out o;
        o.kf_T.resize(12 * (size_t)P.n_kf); o.pt_xyz.resize(3 * (size_t)n_pt); o.ls_plk.resize(6 * (size_t)n_ls); o.ls_end.resize(6 * (size_t)n_ls);
        o.pt_inlier.resize(n_pt); o.ls_inlier.resize(n_ls); o.po_flags.resize(P.n_pobs); o.lo_flags.resize(P.n_lobs);
        last_trace.assign(256, plba_trace_rec{});
        plba_result R{};
        R.kf_T_wc = o.kf_T.data(); R.pt_xyz = o.pt_xyz.data(); R.ls_plk = o.ls_plk.data(); R.ls_end = o.ls_end.data();
        R.pt_inlier = o.pt_inlier.data(); R.ls_inlier = o.ls_inlier.data(); R.po_flags = o.po_flags.data(); R.lo_flags = o.lo_flags.data();
        R.trace = last_trace.data(); R.trace_cap = (int)last_trace.size();
        o.rc = plba_solve(h_, &P, &opt, &R);
        if (o.rc < PLBA_DISCARDED) throw std::runtime_error(std::string("plba_solve: ") + plba_last_error(h_));
        last_trace.resize(R.n_trace < R.trace_cap ? R.n_trace : R.trace_cap);
        return o;
    }

    int hand_lm(int profile, const std::vector<double> &X, const std::vector<int> &kf_list, const std::vector<int> &pt_list, const std::vector<int> &ls_list,
                const std::vector<Vector6i> &pt_obs_list, const std::vector<Vector6i> &ls_obs_list, int shell = PLBA_SHELL_LBA) {
        const int Nkf = (int)kf_list.size();
        const int Npt = pt_obs_list.empty() ? 0 : pt_obs_list.back()[1] + 1;                 // :2359-2360
        const int Nls = ls_obs_list.empty() ? 0 : ls_obs_list.back()[1] + 1;                 // :2441-2442
        const int dl = profile == PLBA_PROFILE_H_END ? 6 : 4;
        Flat f; f.cam = cam;
        std::vector<int> rows(kf_list); std::unordered_map<int, int> row_of;
        for (int j = 0; j < Nkf; j++) { row_of[kf_list[j]] = j; f.kf_slot.push_back(j); }
        auto add_fixed = [&](const std::vector<Vector6i> &obs) { for (const Vector6i &ob : obs) if (ob[4] == -1 && !row_of.count(ob[3])) { row_of[ob[3]] = (int)rows.size(); rows.push_back(ob[3]); f.kf_slot.push_back(-1); } };
        add_fixed(pt_obs_list); add_fixed(ls_obs_list);
        for (int k : rows) push_T(f.kf_T, map_keyframes.at(k)->T_kf_w);
        f.x_pose.assign(X.begin(), X.begin() + 6 * Nkf);
        f.pt_xyz.assign(X.begin() + 6 * Nkf, X.begin() + 6 * Nkf + 3 * Npt);
        const double *lx = X.data() + 6 * Nkf + 3 * Npt;
        for (const Vector6i &ob : pt_obs_list) {
            if (!map_points.at(ob[0]) || !map_keyframes.at(ob[3])) continue;                 // :2368
            f.po_lm.push_back(ob[1]); f.po_kf.push_back(row_of.at(ob[3]));
            const Vector2d &uv = map_points[ob[0]]->obs_list.at(ob[2]); f.po_uv.push_back(uv[0]); f.po_uv.push_back(uv[1]);
        }
        for (const Vector6i &ob : ls_obs_list) {
            if (!map_lines.at(ob[0]) || !map_keyframes.at(ob[3])) continue;                  // :2450
            f.lo_lm.push_back(ob[1]); f.lo_kf.push_back(row_of.at(ob[3]));
            if (profile == PLBA_PROFILE_H_END) { const Vector3d &l = map_lines[ob[0]]->obs_list.at(ob[2]); f.lo_ab.insert(f.lo_ab.end(), {l[0], l[1], l[2], 0.0}); }
            else { const Vector4d &l = map_lines[ob[0]]->NDw_obs_list.at(ob[2]); f.lo_ab.insert(f.lo_ab.end(), l.begin(), l.end()); }
        }
        if (profile == PLBA_PROFILE_H_END) f.ls_end.assign(lx, lx + 6 * Nls);
        else for (int i = 0; i < Nls; i++) for (double v : map_lines.at(ls_list[i])->NDw) f.ls_plk.push_back(v);   // pass 0 reads the MAP Plücker vector (:1744)
        (void)dl;
        Out o = run(f, profile, Npt, Nls, shell);
        if (shell == PLBA_SHELL_LBA && vo_status == VO_INSERTING_KF) return -1;                                         // computed but discarded (:2841, :3011-3012)
        for (int i = 0; i < Nkf; i++) pull_T(o.kf_T, i, map_keyframes[kf_list[i]]->T_kf_w);   // write-back under m_insert_kf (:2844-2882)
        for (int i = 0; i < Npt; i++) { MapPoint *p = map_points[pt_list[i]]; if (!o.pt_inlier[i]) p->inlier = false; for (int k = 0; k < 3; k++) p->point3D[k] = o.pt_xyz[3 * i + k]; }
        for (int i = 0; i < Nls; i++) {
            MapLine *l = map_lines[ls_list[i]]; if (!o.ls_inlier[i]) l->inlier = false;
            for (int k = 0; k < 6; k++) (profile == PLBA_PROFILE_H_END ? l->line3D : l->NDw)[k] = (profile == PLBA_PROFILE_H_END ? o.ls_end : o.ls_plk)[6 * i + k];
        }
        return 0;
    }

    static size_t n_obs_for_culling(const MapPoint *p) { return p->obs_list.size(); }           // :3826
    static size_t n_obs_for_culling(const MapLine *l) { return l->NDw_obs_list.size(); }        // :3865
    template <typename LM>
    void cull(std::vector<LM *> &lms, std::map<int, std::vector<int>> &by_kf, bool point, std::vector<void *> &culled) {
        for (LM *&lm : lms) {
            if (!lm || lm->local || !((long long)max_kf_idx - lm->kf_obs_list.at(0) > 10)) continue;                       // :3824 / :3863
            if (lm->inlier && n_obs_for_culling(lm) >= (size_t)config.min_lm_obs) continue;                                  // :3826 / :3865
            const int kf_obs = lm->kf_obs_list[0], lm_idx = lm->idx;
            std::vector<int> &ids = point ? map_keyframes.at(kf_obs)->stereo_pt_idx : map_keyframes.at(kf_obs)->stereo_ls_idx;
            for (int &v : ids) if (v == lm_idx) { v = -1; break; }                                                           // :3831-3838
            auto it = by_kf.find(kf_obs);
            if (it != by_kf.end()) for (size_t j = 0; j < it->second.size(); j++) if (it->second[j] == lm_idx) { it->second.erase(it->second.begin() + j); break; }   // :3841-3848
            culled.push_back(lm); lm = nullptr;
        }
    }

    template <typename LM>
    int erase_bad(std::vector<LM *> &lms, const std::vector<int32_t> &ob_lm, const std::vector<uint8_t> &flags, bool point) {
        int n_bad = 0;
        std::vector<int> first(lms.size(), -1);
        for (size_t i = 0; i < ob_lm.size(); i++) if (first[ob_lm[i]] < 0) first[ob_lm[i]] = (int)i;
        for (int i = (int)ob_lm.size() - 1; i >= 0; i--) {
            if (!(flags[i] & PLBA_OBS_BAD)) continue;
            n_bad++;
            LM *lm = lms[ob_lm[i]];
            const size_t n_obs = lm->kf_obs_list.size();
            if (n_obs > 1) { const int j = i - first[ob_lm[i]]; if ((size_t)j < n_obs) erase_obs(lm, j); }
            else lm->inlier = false;
        }
        (void)point;
        return n_bad;
    }
    static void erase_obs(MapPoint *p, int j) { p->obs_list.erase(p->obs_list.begin() + j); p->kf_obs_list.erase(p->kf_obs_list.begin() + j); p->sigma_list.erase(p->sigma_list.begin() + j); }
    static void erase_obs(MapLine *l, int j) { l->NDw_obs_list.erase(l->NDw_obs_list.begin() + j); l->kf_obs_list.erase(l->kf_obs_list.begin() + j); l->sigma_list.erase(l->sigma_list.begin() + j); }
};

}  // namespace plba_shim
