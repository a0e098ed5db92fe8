"""The boundary value is computed in double / fixed point here and in float by OpenMM's CUDA CustomCentroidBondForce
(unpinned: OpenMM is absent).  tools/cv_float_contract.py restates the float evaluation and counts the steps of a thermal,
constantly bouncing MMVT run on which the two would decide differently; profiles/r02_cv_float_contract.json holds the long
run.  Here: a short run, and the property that makes the rate tiny -- a difference needs the advanced configuration to land
within a few float ulps of r^2 of the surface."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import cv_float_contract as F  # noqa: E402


def test_float_and_fixed_point_boundary_decisions_agree_on_a_thermal_run():
    for r in F.main(steps=3000)["runs"]:
        assert r["bounce_steps"] > 50, r                      # the run really keeps hitting the surfaces
        for order in ("sequential", "pairwise"):
            assert r["rate_per_step"][order] <= 2e-3, r
            # whenever they differ, the float surface function is within a few ulps of r^2 of zero
            assert r["largest_abs_u_float_at_a_difference_nm2"][order] <= 16 * r["ulp_of_r2_float_nm2"], r
