"""Shared comparison helpers for the parity tests (tolerances from BASELINE.json: cost 1e-9 rel, state 1e-8)."""
import numpy as np

from pl_slam_plucker_b200 import abi

COST_RTOL = 1e-9      # per-iteration cost, relative
STATE_ATOL = 1e-8     # final poses / landmarks
# accept/reject decisions whose |rho| is below this are rounding-dominated in the reference itself (stalled LM):
RHO_MARGIN = 1e-6


def robust_prefix(tr_ref):
    """Number of leading trace records whose accept decision is not rounding-dominated."""
    n = 0
    for t in tr_ref:
        if t["stage"] >= 0 and abs(t["rho"]) < RHO_MARGIN and t["chi_new"] != 0.0:
            break
        n += 1
    return n


def assert_trace_close(ref, got, profile, cost_rtol=COST_RTOL):
    n = min(robust_prefix(ref) if profile == abi.PROFILE_G else len(ref), len(got))
    assert n > 0
    for k in ("stage", "iter", "trial", "accepted", "stop"):
        assert (ref[k][:n] == got[k][:n]).all(), k
    for k in ("chi", "chi_new", "lambda", "dx_norm", "err_pt", "err_ls"):
        a, b = ref[k][:n], got[k][:n]
        fin = np.isfinite(a)
        assert (np.isfinite(b) == fin).all(), k
        np.testing.assert_allclose(b[fin], a[fin], rtol=cost_rtol, atol=1e-300, err_msg=k)
    return n


# Per-observation chi2 (profile G, returned for the host's observation-removal bookkeeping) is a function of the final state: with the
# state agreeing to delta <= 1e-8, d(chi2) = 2 sqrt(chi2) |J| delta with |J| up to ~1e3 px per unit, i.e. a few 1e-6 relative on a small
# residual.  It is therefore compared at rtol 1e-6 / atol 1e-8 (measured: one element in 3 671 at 2.2e-7 relative / 2.4e-9 absolute);
# the FLAGS derived from it are compared exactly except within 1e-6 of the 5.991 gate.
CHI2_RTOL, CHI2_ATOL = 1e-6, 1e-8


def assert_state_close(ref, got, prob, profile, atol=STATE_ATOL, chi2_rtol=CHI2_RTOL):
    np.testing.assert_allclose(got.kf_T_wc, ref.kf_T_wc, rtol=0, atol=atol)
    np.testing.assert_allclose(got.pt_xyz, ref.pt_xyz, rtol=0, atol=atol)
    if profile == abi.PROFILE_H_END:
        np.testing.assert_allclose(got.ls_end, ref.ls_end, rtol=0, atol=atol)
    else:
        np.testing.assert_allclose(got.ls_orth, ref.ls_orth, rtol=0, atol=atol)
        np.testing.assert_allclose(got.ls_plk, ref.ls_plk, rtol=0, atol=atol)
    if profile == abi.PROFILE_G:
        # gating is index work: bit-exact, except for edges whose chi2 sits within 1e-9 of the 5.991 gate
        near = np.abs(ref.po_chi2 - 5.991) < 1e-6
        assert ((ref.po_flags == got.po_flags) | near).all()
        nearl = np.abs(ref.lo_chi2 - 5.991) < 1e-6
        assert ((ref.lo_flags == got.lo_flags) | nearl).all()
        np.testing.assert_allclose(got.po_chi2, ref.po_chi2, rtol=chi2_rtol, atol=CHI2_ATOL)
        np.testing.assert_allclose(got.lo_chi2, ref.lo_chi2, rtol=chi2_rtol, atol=CHI2_ATOL)
    else:
        assert (ref.pt_inlier == got.pt_inlier).all() and (ref.ls_inlier == got.ls_inlier).all()
