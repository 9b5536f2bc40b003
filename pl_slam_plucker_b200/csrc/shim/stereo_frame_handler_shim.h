// Host shim with the reference's names for the two steps either side of the LBA path, on top of the
// C ABI (include/plba.h).  Header-only C++; stand-in feature types with the reference's member names
// (include2/stereoFeatures.h: PointFeature :33-58, LineFeature :60-150).
//
//   void StereoFrameHandler::gaussNewtonOptimizationforPluker(DT, DT_cov, err_, max_iters)   src2/stereoFrameHandler.cpp:803-853
//   Vector6d line_pluker = pipi_plk(pi_from_ppp(..), pi_from_ppp(..))                          src2/stereoFrame.cpp:381-397
//   NDw, re-projection gate of a new MapLine                                                   src/mapHandler.cpp:449-492
//
// Only packing happens here; the arithmetic runs in the CUDA library.  Errors of the library are thrown: no CPU fallback.
#pragma once
#include <array>
#include <cmath>
#include <list>
#include <stdexcept>
#include <string>
#include <vector>
#include "../../../include/plba.h"

namespace plba_shim {

typedef std::array<double, 16> Matrix4d_rm;    // row-major 4x4
typedef std::array<double, 36> Matrix6d_rm;    // row-major 6x6

struct PointFeature { std::array<double, 3> P{}; std::array<double, 2> pl{}, pl_obs{}; bool inlier = true; double sigma2 = 1.0; };
struct LineFeature {
    std::array<double, 3> sP{}, eP{}; std::array<double, 6> NDc{};
    std::array<double, 2> spl{}, epl{}, spl_obs{}, epl_obs{};
    bool inlier = true; double sigma2 = 1.0;
};
struct TrackConfig { double homog_th = 1e-7, min_error = 1e-7, min_error_change = 1e-7; int max_iters = 5, max_iters_ref = 10; };   // src2/config.cpp:80-85

class StereoFrameHandlerShim {
public:
    std::list<PointFeature *> matched_pt;       // include2/stereoFrameHandler.h: list<PointFeature*> matched_pt
    std::list<LineFeature *> matched_ls;
    double fx = 0, fy = 0, cx = 0, cy = 0, b = 0;
    TrackConfig config;

    explicit StereoFrameHandlerShim(int device = 0) { if (plba_create(device, nullptr, &h_) != PLBA_OK) throw std::runtime_error("plba_create failed: a CUDA device is required (no CPU fallback)"); }
    ~StereoFrameHandlerShim() { plba_destroy(h_); }
    StereoFrameHandlerShim(const StereoFrameHandlerShim &) = delete;

    // src2/stereoFrameHandler.cpp:803-853: DT in/out, DT_cov and err_ out (err_ = -1 and DT restored when the solve failed)
    void gaussNewtonOptimizationforPluker(Matrix4d_rm &DT, Matrix6d_rm &DT_cov, double &err_, int max_iters) {
        std::vector<double> pP, pO, sP, eP, nd, ob, sg, s2; std::vector<uint8_t> pin, lin;
        for (const PointFeature *p : matched_pt) { pP.insert(pP.end(), p->P.begin(), p->P.end()); pO.insert(pO.end(), p->pl_obs.begin(), p->pl_obs.end()); pin.push_back(p->inlier ? 1 : 0); }
        for (const LineFeature *l : matched_ls) {
            sP.insert(sP.end(), l->sP.begin(), l->sP.end()); eP.insert(eP.end(), l->eP.begin(), l->eP.end()); nd.insert(nd.end(), l->NDc.begin(), l->NDc.end());
            ob.insert(ob.end(), {l->spl_obs[0], l->spl_obs[1], l->epl_obs[0], l->epl_obs[1]}); sg.insert(sg.end(), {l->spl[0], l->spl[1], l->epl[0], l->epl[1]});
            s2.push_back(l->sigma2); lin.push_back(l->inlier ? 1 : 0);
        }
        plba_track_frame F{};
        F.n_pt = (int32_t)pin.size(); F.n_ls = (int32_t)lin.size();
        for (int i = 0; i < 12; i++) F.DT[i] = DT[i];
        F.pt_P = pP.data(); F.pt_obs = pO.data(); F.pt_inlier = pin.data();
        F.ls_sP = sP.data(); F.ls_eP = eP.data(); F.ls_NDc = nd.data(); F.ls_obs = ob.data(); F.ls_seg = sg.data(); F.ls_sigma2 = s2.data(); F.ls_inlier = lin.data();
        plba_track_options o; plba_track_default_options(&o);
        o.cam[0] = fx; o.cam[1] = fy; o.cam[2] = cx; o.cam[3] = cy;
        o.homog_th = config.homog_th; o.min_error = config.min_error; o.min_error_change = config.min_error_change; o.max_iters = max_iters;
        plba_track_result R{};
        if (plba_track_solve(h_, 1, &F, &o, &R) != PLBA_OK) throw std::runtime_error(std::string("plba_track_solve: ") + plba_last_error(h_));
        for (int i = 0; i < 12; i++) DT[i] = R.DT[i];
        DT[12] = DT[13] = DT[14] = 0; DT[15] = 1;
        for (int i = 0; i < 36; i++) DT_cov[i] = R.DT_cov[i];
        err_ = R.err;
        last_iters = R.iters; last_good = R.good != 0;
    }
    int last_iters = 0; bool last_good = false;

    // one new MapLine candidate: stereo triangulation in the creating keyframe (src2/stereoFrame.cpp:381-397), NDw and the gate (src/mapHandler.cpp:449-492)
    struct NewLine { std::array<double, 6> NDc, NDw; double error, error2; bool accepted; };
    NewLine createMapLine(const std::array<double, 4> &seg_l, const std::array<double, 4> &seg_r, const std::array<double, 4> &seg_curr,
                          const Matrix4d_rm &T_prev_w, const Matrix4d_rm &T_curr_w) {
        double T[24]; for (int i = 0; i < 12; i++) { T[i] = T_prev_w[i]; T[12 + i] = T_curr_w[i]; }
        const int32_t kp = 0, kc = 1;
        plba_newline_batch B{}; B.n = 1; B.n_kf = 2; B.cam[0] = fx; B.cam[1] = fy; B.cam[2] = cx; B.cam[3] = cy; B.cam[4] = b;
        B.seg_l = seg_l.data(); B.seg_r = seg_r.data(); B.seg_curr = seg_curr.data(); B.kf_prev = &kp; B.kf_curr = &kc; B.kf_T_wc = T;
        NewLine out{}; uint8_t acc = 0;
        if (plba_create_lines(h_, &B, out.NDc.data(), out.NDw.data(), &out.error, &out.error2, &acc) != PLBA_OK) throw std::runtime_error(std::string("plba_create_lines: ") + plba_last_error(h_));
        out.accepted = acc != 0;
        return out;
    }

private:
    plba_handle h_ = nullptr;
};

}  // namespace plba_shim
