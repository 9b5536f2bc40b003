"""Host-side mirror of the reference's LBA interface (PLSLAM::MapHandler, include/mapHandler.h:128-134).

Same names, argument meaning and return / error behaviour as the reference so that the parity tests read like calls into
the reference:

    MapHandler.localBundleAdjustment()                       src/mapHandler.cpp:1392-1502
    MapHandler.levMarquardtOptimizationLBA(...)              src/mapHandler.cpp:2334-3016   -> profile H_END
    MapHandler.localBundleAdjustmentForPluker()              src/mapHandler.cpp:1505-1615
    MapHandler.levMarquardtOptimizationLBAForPluker(...)     src/mapHandler.cpp:1618-2332   -> profile H_PLK
    MapHandler.localBundleAdjustmentForPlukerWithG2O()       src/mapHandler.cpp:5851-6323   -> profile G

Only the flattening (index bookkeeping) and the write-back live here; all arithmetic runs in the CUDA library behind
include/plba.h.  The C++ twin of this file is csrc/shim/map_handler_shim.h.
"""
import numpy as np

from . import abi

VO_PROCESSING, VO_INSERTING_KF = 0, 1


class SlamConfig:
    """The LBA-relevant keys of SlamConfig / Config (src/slamConfig.cpp:61-67, src2/config.cpp:80-85)."""
    min_lm_obs = 5            # src/slamConfig.cpp:48
    lambda_lba_lm = 1e-5
    lambda_lba_k = 10.0
    max_iters_lba = 15
    homog_th = 1e-7
    min_error = 1e-7
    min_error_change = 1e-7


class PinholeStereoCamera:
    def __init__(self, fx, fy, cx, cy):
        self.fx, self.fy, self.cx, self.cy = float(fx), float(fy), float(cx), float(cy)

    def getFx(self): return self.fx
    def getFy(self): return self.fy
    def getCx(self): return self.cx
    def getCy(self): return self.cy


class KeyFrame:   # include/keyFrame.h:47-79
    def __init__(self, kf_idx, T_kf_w, x_kf_w=None, local=False):
        self.kf_idx = int(kf_idx)
        self.T_kf_w = np.array(T_kf_w, dtype=np.float64).reshape(4, 4)
        self.x_kf_w = None if x_kf_w is None else np.array(x_kf_w, dtype=np.float64)
        self.local = bool(local)
        # stereo_frame->stereo_pt[i]->idx / stereo_ls[i]->idx (include/stereoFeatures.h): map landmark id of each stereo feature, -1 = none
        self.stereo_pt_idx, self.stereo_ls_idx = [], []


class MapPoint:   # include/mapFeatures.h:41-70
    def __init__(self, idx, point3D, local=True):
        self.idx = int(idx)
        self.point3D = np.array(point3D, dtype=np.float64)
        self.obs_list, self.kf_obs_list, self.sigma_list = [], [], []
        self.inlier, self.local = True, bool(local)

    def addMapPointObservation(self, kf_obs, obs, sigma2=1.0):
        self.obs_list.append(np.array(obs, dtype=np.float64)); self.kf_obs_list.append(int(kf_obs)); self.sigma_list.append(float(sigma2))


class MapLine:    # include/mapFeatures.h:72-122
    def __init__(self, idx, line3D=None, NDw=None, local=True):
        self.idx = int(idx)
        self.line3D = None if line3D is None else np.array(line3D, dtype=np.float64)
        self.NDw = None if NDw is None else np.array(NDw, dtype=np.float64)
        self.obs_list, self.NDw_obs_list, self.kf_obs_list, self.sigma_list = [], [], [], []
        self.inlier, self.local = True, bool(local)

    def addMapLineObservation(self, kf_obs, obs, sigma2=1.0):
        """obs of length 3 = normalised 2-D line (endpoint mode); length 4 = endpoint pixels (Plücker mode, src/mapFeatures.cpp:132-138)."""
        obs = np.array(obs, dtype=np.float64)
        (self.obs_list if obs.size == 3 else self.NDw_obs_list).append(obs)
        self.kf_obs_list.append(int(kf_obs)); self.sigma_list.append(float(sigma2))


class MapHandler:
    """Map container + the reference's LBA entry points, with the numeric core replaced by the CUDA library."""

    def __init__(self, cam, solver, config=None, quirks=abi.QUIRKS_FAITHFUL):
        self.cam, self.solver, self.cfg, self.quirks = cam, solver, config or SlamConfig(), quirks
        self.map_keyframes, self.map_points, self.map_lines = [], [], []
        self.full_graph = None
        self.max_kf_idx = 0                   # include/mapHandler.h:156
        self.map_points_kf_idx, self.map_lines_kf_idx = {}, {}      # base KF -> landmark ids first seen there (include/mapHandler.h:148-149)
        self.vo_status = VO_PROCESSING        # never assigned in the reference (Q17): treated as VO_PROCESSING
        self.last_result = None

    # ------------------------------------------------------------------ drivers (flattening) -----------------
    def _gather(self, pluker, global_ba=False):
        """src/mapHandler.cpp:1396-1493 (and the Plücker twin :1509-1607): X_aux, id lists and Vector6i observation tuples.
        global_ba: the flattening of globalBundleAdjustment (:3022-3117), which ignores the `local` flags."""
        X_aux, kf_list = [], []
        for kf in self.map_keyframes:
            if kf is not None and (global_ba or kf.local) and kf.kf_idx != 0:
                X_aux.extend(kf.x_kf_w.tolist()); kf_list.append(kf.kf_idx)
        kf_pos = {k: j for j, k in enumerate(kf_list)}     # replaces the O(Nobs*Nkf) linear search (:1437-1444)
        pt_obs_list, pt_list = [], []
        for loc, pt in enumerate(p for p in self.map_points if p is not None and (global_ba or p.local)):
            X_aux.extend(pt.point3D.tolist())
            for i in range(len(pt.obs_list)):
                kfo = pt.kf_obs_list[i]
                pt_obs_list.append((pt.idx, loc, i, kfo, kf_pos.get(kfo, -1), 1))
            pt_list.append(pt.idx)
        ls_obs_list, ls_list = [], []
        for loc, ls in enumerate(l for l in self.map_lines if l is not None and (global_ba or l.local)):
            if pluker:
                X_aux.extend(_pluker_to_orth(ls.NDw).tolist())       # :1577
                n_obs = len(ls.obs_list) if self.quirks == abi.QUIRKS_FAITHFUL else len(ls.NDw_obs_list)   # Q10 (:1582)
            else:
                X_aux.extend(ls.line3D.tolist())
                n_obs = len(ls.obs_list)
            for i in range(n_obs):
                kfo = ls.kf_obs_list[i]
                ls_obs_list.append((ls.idx, loc, i, kfo, kf_pos.get(kfo, -1), 1))
            ls_list.append(ls.idx)
        return X_aux, kf_list, pt_list, ls_list, pt_obs_list, ls_obs_list

    def localBundleAdjustment(self):
        a = self._gather(False)
        if len(a[4]) + len(a[5]) != 0:
            return self.levMarquardtOptimizationLBA(*a)
        return -1                                            # :1499-1500

    def localBundleAdjustmentForPluker(self):
        a = self._gather(True)
        if len(a[4]) + len(a[5]) != 0:
            return self.levMarquardtOptimizationLBAForPluker(*a)
        return -1

    def globalBundleAdjustment(self):
        """src/mapHandler.cpp:3022-3126: every non-NULL KF but KF 0 and every non-NULL landmark; void, no emptiness test."""
        self.levMarquardtOptimizationGBA(*self._gather(False, global_ba=True))

    # ------------------------------------------------------------------ hand-LM entry points -----------------
    def levMarquardtOptimizationGBA(self, X_aux, kf_list, pt_list, ls_list, pt_obs_list, ls_obs_list):
        """src/mapHandler.cpp:3128-3728 (include/mapHandler.h:137): H_END arithmetic inside the GBA shell; returns nothing."""
        if len(pt_obs_list) + len(ls_obs_list) != 0:
            self._hand_lm(abi.PROFILE_H_END, X_aux, kf_list, pt_list, ls_list, pt_obs_list, ls_obs_list, shell=abi.SHELL_GBA)

    def levMarquardtOptimizationLBA(self, X_aux, kf_list, pt_list, ls_list, pt_obs_list, ls_obs_list):
        return self._hand_lm(abi.PROFILE_H_END, X_aux, kf_list, pt_list, ls_list, pt_obs_list, ls_obs_list)

    def levMarquardtOptimizationLBAForPluker(self, X_aux, kf_list, pt_list, ls_list, pt_obs_list, ls_obs_list):
        return self._hand_lm(abi.PROFILE_H_PLK, X_aux, kf_list, pt_list, ls_list, pt_obs_list, ls_obs_list)

    def _hand_lm(self, profile, X_aux, kf_list, pt_list, ls_list, pt_obs_list, ls_obs_list, shell=abi.SHELL_LBA):
        X = np.asarray(X_aux, dtype=np.float64)
        Nkf = len(kf_list)
        Npt = (pt_obs_list[-1][1] + 1) if len(pt_obs_list) else 0          # :2359-2360
        Nls = (ls_obs_list[-1][1] + 1) if len(ls_obs_list) else 0          # :2441-2442
        dl = 6 if profile == abi.PROFILE_H_END else 4
        # KF table: the free KFs of kf_list, then every fixed observer (kf_idx_loc == -1) in order of first appearance
        rows, slot = list(kf_list), list(range(Nkf))
        row_of = {k: j for j, k in enumerate(rows)}
        for ob in list(pt_obs_list) + list(ls_obs_list):
            if ob[4] == -1 and ob[3] not in row_of:
                row_of[ob[3]] = len(rows); rows.append(ob[3]); slot.append(-1)
        T = np.stack([self.map_keyframes[k].T_kf_w[:3, :].reshape(12) for k in rows]) if rows else np.zeros((0, 12))
        pts = X[6 * Nkf:6 * Nkf + 3 * Npt].reshape(-1, 3)
        lns = X[6 * Nkf + 3 * Npt:6 * Nkf + 3 * Npt + dl * Nls].reshape(-1, dl)
        po = [ob for ob in pt_obs_list if self.map_points[ob[0]] is not None and self.map_keyframes[ob[3]] is not None]   # :2368
        lo = [ob for ob in ls_obs_list if self.map_lines[ob[0]] is not None and self.map_keyframes[ob[3]] is not None]    # :2450
        po_uv = np.array([self.map_points[ob[0]].obs_list[ob[2]] for ob in po]).reshape(-1, 2)
        if profile == abi.PROFILE_H_END:
            lo_ab = np.zeros((len(lo), 4))
            for i, ob in enumerate(lo):
                lo_ab[i, :3] = self.map_lines[ob[0]].obs_list[ob[2]]
            ls_end, ls_plk = lns, None
        else:
            lo_ab = np.array([self.map_lines[ob[0]].NDw_obs_list[ob[2]] for ob in lo]).reshape(-1, 4)
            ls_end = None
            # pass 0 reads the MAP Plücker vector (:1744), later iterations the orth part of X
            ls_plk = np.stack([self.map_lines[i].NDw for i in ls_list[:Nls]]) if Nls else np.zeros((0, 6))
        prob = abi.Problem([self.cam.getFx(), self.cam.getFy(), self.cam.getCx(), self.cam.getCy()], T, np.array(slot, np.int32), pts,
                           np.array([ob[1] for ob in po], np.int32), np.array([row_of[ob[3]] for ob in po], np.int32), po_uv,
                           ls_plk=ls_plk, ls_end=ls_end, lo_lm=np.array([ob[1] for ob in lo], np.int32),
                           lo_kf=np.array([row_of[ob[3]] for ob in lo], np.int32), lo_ab=lo_ab, x_pose=X[:6 * Nkf].reshape(-1, 6))
        opt = abi.Options(profile, self.quirks, shell=shell, lambda_lba_lm=self.cfg.lambda_lba_lm, lambda_lba_k=self.cfg.lambda_lba_k,
                          max_iters_lba=self.cfg.max_iters_lba, homog_th=self.cfg.homog_th, min_error=self.cfg.min_error,
                          min_error_change=self.cfg.min_error_change)
        res = self.solver.solve(prob, opt)
        self.last_result = res
        if shell == abi.SHELL_LBA and self.vo_status == VO_INSERTING_KF:               # :2841 / :3011-3012 : computed but discarded
            return -1
        # write-back under m_insert_kf (:2844-2882)
        for i in range(Nkf):
            Tk = np.eye(4); Tk[:3, :] = res.kf_T_wc[i].reshape(3, 4)
            self.map_keyframes[kf_list[i]].T_kf_w = Tk
        for i in range(Npt):
            mp = self.map_points[pt_list[i]]
            if not res.pt_inlier[i]:
                mp.inlier = False                           # moved by more than 0.01 (:2858-2860)
            mp.point3D = res.pt_xyz[i].copy()
        for i in range(Nls):
            ml = self.map_lines[ls_list[i]]
            if not res.ls_inlier[i]:
                ml.inlier = False
            if profile == abi.PROFILE_H_END:
                ml.line3D = res.ls_end[i].copy()
            else:
                ml.NDw = res.ls_plk[i].copy()
        return 0

    # ------------------------------------------------------------------ g2o path -----------------------------
    def localBundleAdjustmentForPlukerWithG2O(self):
        """src/mapHandler.cpp:5851-6323.  Returns nothing, like the reference; statistics are left in self.last_result."""
        local_pt = [p for p in self.map_points if p is not None and p.local]
        local_ls = [l for l in self.map_lines if l is not None and l.local]
        nofix = {kf.kf_idx: kf for kf in self.map_keyframes if kf is not None and kf.local}        # :5870-5875
        fix = {}
        for lm in local_pt + local_ls:                                                            # :5888-5919
            for k in lm.kf_obs_list:
                kf = self.map_keyframes[k]
                if kf.kf_idx != k:
                    raise SystemExit("[Wrong index in the map_keyframes and landmark obs.....]")  # exit(0) in the reference
                if not kf.local:
                    fix[k] = kf; kf.local = True
        all_ids = sorted(set(nofix) | set(fix))
        row_of = {k: j for j, k in enumerate(all_ids)}
        slot, ns = [], 0
        for k in all_ids:
            free = (k in nofix) and k != 0                                                        # KF 0 fixed (:5943-5945)
            slot.append(ns if free else -1); ns += int(free)
        kfs = [nofix.get(k, fix.get(k)) for k in all_ids]
        T = np.stack([kf.T_kf_w[:3, :].reshape(12) for kf in kfs]) if kfs else np.zeros((0, 12))
        po_lm, po_kf, po_uv, po_s, lo_lm, lo_kf, lo_ab, lo_s = [], [], [], [], [], [], [], []
        for loc, p in enumerate(local_pt):
            for i, k in enumerate(p.kf_obs_list):
                po_lm.append(loc); po_kf.append(row_of[k]); po_uv.append(p.obs_list[i]); po_s.append(p.sigma_list[i])
        for loc, l in enumerate(local_ls):
            for i, k in enumerate(l.kf_obs_list):
                lo_lm.append(loc); lo_kf.append(row_of[k]); lo_ab.append(l.NDw_obs_list[i]); lo_s.append(l.sigma_list[i])
        prob = abi.Problem([self.cam.getFx(), self.cam.getFy(), self.cam.getCx(), self.cam.getCy()], T, np.array(slot, np.int32),
                           np.array([p.point3D for p in local_pt]).reshape(-1, 3), np.array(po_lm, np.int32), np.array(po_kf, np.int32),
                           np.array(po_uv).reshape(-1, 2), ls_plk=np.array([l.NDw for l in local_ls]).reshape(-1, 6),
                           lo_lm=np.array(lo_lm, np.int32), lo_kf=np.array(lo_kf, np.int32), lo_ab=np.array(lo_ab).reshape(-1, 4),
                           po_sig2=np.array(po_s), lo_sig2=np.array(lo_s))
        if prob.n_obs == 0:
            return
        res = self.solver.solve(prob, abi.Options(abi.PROFILE_G, self.quirks))
        self.last_result = res
        # bad observations (:6156-6293), reverse order as the reference so that erase indices stay valid
        self.bad_point_obs = self._erase_bad(local_pt, po_lm, res.po_flags, all_ids, po_kf, point=True)
        self.bad_line_obs = self._erase_bad(local_ls, lo_lm, res.lo_flags, all_ids, lo_kf, point=False)
        for j, k in enumerate(all_ids):                                                           # :6297-6303
            if k in nofix:
                Tk = np.eye(4); Tk[:3, :] = res.kf_T_wc[j].reshape(3, 4)
                nofix[k].T_kf_w = Tk
        for loc, p in enumerate(local_pt):
            p.point3D = res.pt_xyz[loc].copy()                                                    # :6306-6311
        for loc, l in enumerate(local_ls):
            l.NDw = res.ls_plk[loc].copy()                                                        # :6314-6319

    def localMappingStep(self):
        """The LBA part of one pass of localMappingThread in Plücker mode (src/mapHandler.cpp:1277-1279): LBA, then culling."""
        self.localBundleAdjustmentForPlukerWithG2O()
        self.removeBadMapLandmarksForPluker()

    def removeBadMapLandmarksForPluker(self):
        """src/mapHandler.cpp:3816-3897: the step right after the LBA.  A landmark that is no longer local, was first
        seen more than 10 keyframes ago and is an outlier (the `inlier` flag the LBA write-back cleared) or has fewer than minLMObs
        observations is deleted: its id is cleared in the stereo features of its base keyframe and in map_*_kf_idx."""
        removed = [0, 0]
        for cls, (lms, by_kf, feat) in enumerate(((self.map_points, self.map_points_kf_idx, "stereo_pt_idx"), (self.map_lines, self.map_lines_kf_idx, "stereo_ls_idx"))):
            for i, lm in enumerate(lms):
                if lm is None or lm.local or not (self.max_kf_idx - lm.kf_obs_list[0] > 10):
                    continue
                n_obs = len(lm.obs_list) if cls == 0 else len(lm.NDw_obs_list)          # :3826 / :3865
                if lm.inlier and n_obs >= self.cfg.min_lm_obs:
                    continue
                kf_obs = lm.kf_obs_list[0]
                ids = getattr(self.map_keyframes[kf_obs], feat)
                for j, v in enumerate(ids):                                              # first match only (:3831-3838)
                    if v == lm.idx:
                        ids[j] = -1
                        break
                lst = by_kf.get(kf_obs, [])
                if lm.idx in lst:
                    lst.remove(lm.idx)                                                   # first occurrence (:3841-3848)
                lms[i] = None
                removed[cls] += 1
        return tuple(removed)

    def _erase_bad(self, lms, ob_lm, flags, all_ids, ob_kf, point):
        n_bad = 0
        first = {}
        for i, l in enumerate(ob_lm):
            first.setdefault(l, i)
        for i in range(len(ob_lm) - 1, -1, -1):
            if not (flags[i] & abi.OBS_BAD):
                continue
            n_bad += 1
            lm = lms[ob_lm[i]]
            obs = lm.obs_list if point else lm.NDw_obs_list
            if len(obs) > 1:
                j = i - first[ob_lm[i]]                      # vpLmObsIdx: position inside the landmark at graph-build time
                kf_obs = all_ids[ob_kf[i]]
                if j < len(obs):
                    del obs[j]; del lm.kf_obs_list[j]; del lm.sigma_list[j]
                if self.full_graph is not None:
                    for k in lm.kf_obs_list:
                        if k != kf_obs:
                            self.full_graph[kf_obs][k] -= 1; self.full_graph[k][kf_obs] -= 1
            else:
                lm.inlier = False
        return n_bad


def _pluker_to_orth(pl):
    """MapLine::changePlukerToOrth (src/mapFeatures.cpp:186-201) — index bookkeeping only needs it to fill X_aux."""
    n, d = pl[:3], pl[3:]
    nn, dn = np.linalg.norm(n), np.linalg.norm(d)
    u1, u2 = n / nn, d / dn
    c = np.cross(n, d); u3 = c / np.linalg.norm(c)
    return np.array([np.arctan2(u2[2], u3[2]), np.arcsin(-u1[2]), np.arctan2(u1[1], u1[0]), np.arcsin(dn / np.sqrt(nn * nn + dn * dn))])


def map_from_problem(prob, solver, quirks=abi.QUIRKS_FAITHFUL, endpoint_lines=False):
    """Builds a MapHandler whose map reproduces a flattened `Problem` (test helper: the inverse of the drivers)."""
    cam = PinholeStereoCamera(*prob.cam)
    mh = MapHandler(cam, solver, quirks=quirks)
    nfix = int((prob.kf_slot < 0).sum())
    for k in range(prob.n_kf):
        T = np.eye(4); T[:3, :] = prob.kf_T_wc[k].reshape(3, 4)
        s = prob.kf_slot[k]
        mh.map_keyframes.append(KeyFrame(k, T, x_kf_w=None if s < 0 or prob.x_pose is None else prob.x_pose[s], local=(s >= 0)))
    for l in range(prob.n_pt):
        mh.map_points.append(MapPoint(l, prob.pt_xyz[l]))
    for i in range(prob.n_pobs):
        mh.map_points[prob.po_lm[i]].addMapPointObservation(prob.po_kf[i], prob.po_uv[i])
    for l in range(prob.n_ls):
        mh.map_lines.append(MapLine(l, line3D=None if prob.ls_end is None else prob.ls_end[l], NDw=None if prob.ls_plk is None else prob.ls_plk[l]))
    for i in range(prob.n_lobs):
        ob = prob.lo_ab[i][:3] if endpoint_lines else prob.lo_ab[i]
        mh.map_lines[prob.lo_lm[i]].addMapLineObservation(prob.lo_kf[i], ob)
    assert nfix >= 1
    return mh
