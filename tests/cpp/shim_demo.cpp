// Builds a pointer-style map from a synthetic scene and runs the reference-named LBA entry points of the C++ host shim.
// Usage: shim_demo [n_pt n_ls]   (prints one line of key=value pairs; exit code 0 on success)
#include <cstdio>
#include <cstdlib>
#include <memory>
#include "../../pl_slam_plucker_b200/csrc/shim/map_handler_shim.h"
#include "../../pl_slam_plucker_b200/csrc/shim/stereo_frame_handler_shim.h"
using namespace plba_shim;

int main(int argc, char **argv) {
    plba_scene_spec sp; plba_scene_preset(1, &sp);
    sp.n_kf_free = 5; sp.n_kf_fixed = 2; sp.n_pt = argc > 1 ? std::atoi(argv[1]) : 120; sp.n_ls = argc > 2 ? std::atoi(argv[2]) : 30; sp.seed = 13;
    plba_scene sc; if (plba_scene_create(&sp, &sc) != 0) return 2;
    plba_problem P; plba_scene_problem(sc, &P);
    PinholeStereoCamera cam{P.cam[0], P.cam[1], P.cam[2], P.cam[3]};
    std::vector<std::unique_ptr<KeyFrame>> kfs; std::vector<std::unique_ptr<MapPoint>> pts; std::vector<std::unique_ptr<MapLine>> lns;
    try {
        MapHandlerShim mh(0);
        mh.cam = &cam;
        for (int k = 0; k < P.n_kf; k++) {
            kfs.emplace_back(new KeyFrame()); KeyFrame *kf = kfs.back().get();
            kf->kf_idx = k; kf->local = P.kf_slot[k] >= 0;
            for (int i = 0; i < 12; i++) kf->T_kf_w[i] = P.kf_T_wc[12 * k + i];
            kf->T_kf_w[15] = 1.0;
            if (P.kf_slot[k] >= 0 && P.x_pose) for (int i = 0; i < 6; i++) kf->x_kf_w[i] = P.x_pose[6 * P.kf_slot[k] + i];
            mh.map_keyframes.push_back(kf);
        }
        for (int l = 0; l < P.n_pt; l++) { pts.emplace_back(new MapPoint()); pts.back()->idx = l; for (int i = 0; i < 3; i++) pts.back()->point3D[i] = P.pt_xyz[3 * l + i]; mh.map_points.push_back(pts.back().get()); }
        for (int i = 0; i < P.n_pobs; i++) mh.map_points[P.po_lm[i]]->addMapPointObservation(P.po_kf[i], {P.po_uv[2 * i], P.po_uv[2 * i + 1]});
        for (int l = 0; l < P.n_ls; l++) { lns.emplace_back(new MapLine()); lns.back()->idx = l; for (int i = 0; i < 6; i++) lns.back()->NDw[i] = P.ls_plk[6 * l + i]; mh.map_lines.push_back(lns.back().get()); }
        for (int i = 0; i < P.n_lobs; i++) { MapLine *l = mh.map_lines[P.lo_lm[i]]; l->NDw_obs_list.push_back({P.lo_ab[4 * i], P.lo_ab[4 * i + 1], P.lo_ab[4 * i + 2], P.lo_ab[4 * i + 3]}); l->kf_obs_list.push_back(P.lo_kf[i]); l->sigma_list.push_back(1.0); }
        mh.localBundleAdjustmentForPlukerWithG2O();
        const double chi0 = mh.last_trace.empty() ? -1 : mh.last_trace.front().chi, chi1 = mh.last_trace.empty() ? -1 : mh.last_trace.back().chi_new;
        std::printf("g2o_path trials=%zu chi_first=%.9e chi_last=%.9e bad_pt_obs=%d bad_ls_obs=%d T2=%.12f,%.12f,%.12f\n", mh.last_trace.size(), chi0, chi1,
                    mh.n_bad_point_obs, mh.n_bad_line_obs, mh.map_keyframes[2]->T_kf_w[3], mh.map_keyframes[2]->T_kf_w[7], mh.map_keyframes[2]->T_kf_w[11]);
        const int rc = mh.localBundleAdjustmentForPluker();          // hand LM on the (now optimised) map: faithful Q10 => lines carry no observations
        std::printf("hand_lm rc=%d iters=%zu\n", rc, mh.last_trace.size());
        mh.vo_status = VO_INSERTING_KF;
        std::printf("discarded rc=%d\n", mh.localBundleAdjustmentForPluker());
        {   // culling right after the LBA (src/mapHandler.cpp:1278-1279): every landmark is old and non-local now; the ones the g2o path
            // left with fewer than minLMObs observations (or flagged as outliers) go
            mh.max_kf_idx = 100;
            size_t expect = 0;
            for (MapPoint *p : mh.map_points) { p->local = false; mh.map_points_kf_idx[p->kf_obs_list[0]].push_back(p->idx); if (!p->inlier || p->obs_list.size() < 5) expect++; }
            for (MapLine *l : mh.map_lines) { l->local = false; if (!l->inlier || l->NDw_obs_list.size() < 5) expect++; }
            const size_t culled = mh.removeBadMapLandmarksForPluker().size();
            size_t left = 0; for (MapPoint *p : mh.map_points) if (p) left++;
            std::printf("culled=%zu expected=%zu points_left=%zu\n", culled, expect, left);
            // (the unique_ptr vectors above still own the culled objects; restore the slots so that the GBA below sees the whole map)
            for (size_t i = 0; i < pts.size(); i++) { mh.map_points[i] = pts[i].get(); pts[i]->local = true; }
            for (size_t i = 0; i < lns.size(); i++) { mh.map_lines[i] = lns[i].get(); lns[i]->local = true; }
        }
        mh.globalBundleAdjustment();                                  // shutdown-time global BA: ignores vo_status and the `local` flags, returns nothing
        std::printf("gba iters=%zu\n", mh.last_trace.size());
        {   // the steps either side of the path under the reference's names: pose tracking on 60 synthetic point matches, one new map line
            StereoFrameHandlerShim sfh(0);
            sfh.fx = cam.fx; sfh.fy = cam.fy; sfh.cx = cam.cx; sfh.cy = cam.cy; sfh.b = 0.110078;
            std::vector<std::unique_ptr<PointFeature>> feats;
            const double tz = 0.05, tx = 0.02;                                  // true motion: pure translation previous -> current frame
            for (int i = 0; i < 60; i++) {
                feats.emplace_back(new PointFeature());
                PointFeature *f = feats.back().get();
                f->P = {-2.0 + 0.07 * i, -1.0 + 0.035 * ((i * 7) % 60), 4.0 + 0.1 * ((i * 13) % 40)};
                f->pl_obs = {cam.cx + cam.fx * (f->P[0] + tx) / (f->P[2] + tz), cam.cy + cam.fy * f->P[1] / (f->P[2] + tz)};
                sfh.matched_pt.push_back(f);
            }
            Matrix4d_rm DT{}; DT[0] = DT[5] = DT[10] = DT[15] = 1.0;
            Matrix6d_rm cov{}; double err = 0;
            sfh.gaussNewtonOptimizationforPluker(DT, cov, err, sfh.config.max_iters_ref);
            std::printf("track good=%d iters=%d t=%.6f,%.6f,%.6f err=%.3e\n", sfh.last_good ? 1 : 0, sfh.last_iters, DT[3], DT[7], DT[11], err);
            // a horizontal-ish 3-D segment seen by a rectified pair at the identity pose and by a second keyframe 0.3 m to the right
            const double P1[3] = {-0.5, 0.2, 5.0}, P2[3] = {0.6, -0.3, 6.0};
            auto px = [&](const double *P, double off) { return std::array<double, 2>{cam.cx + cam.fx * (P[0] - off) / P[2], cam.cy + cam.fy * P[1] / P[2]}; };
            const auto a = px(P1, 0), c2 = px(P2, 0), ar = px(P1, sfh.b), cr = px(P2, sfh.b), ac = px(P1, 0.3), cc = px(P2, 0.3);
            Matrix4d_rm T0{}; T0[0] = T0[5] = T0[10] = T0[15] = 1.0; Matrix4d_rm T1 = T0; T1[3] = 0.3;
            const auto nl = sfh.createMapLine({a[0], a[1], c2[0], c2[1]}, {ar[0], ar[1], cr[0], cr[1]}, {ac[0], ac[1], cc[0], cc[1]}, T0, T1);
            std::printf("newline accepted=%d error=%.3e error2=%.3e d=%.6f,%.6f,%.6f\n", nl.accepted ? 1 : 0, nl.error, nl.error2, nl.NDw[3], nl.NDw[4], nl.NDw[5]);
        }
    } catch (const std::exception &e) { std::fprintf(stderr, "error: %s\n", e.what()); plba_scene_destroy(sc); return 1; }
    plba_scene_destroy(sc);
    return 0;
}
