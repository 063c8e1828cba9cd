"""Parity of the EXPERIMENTAL warp-specialised slab kernel (jots_b200/csrc/apply_slab5.cuh, JOTS_SLAB_VARIANT=6).  The kernel
was written after the GPU budget of round 1 was spent and has not run on hardware yet, so these tests are skipped unless
JOTS_TEST_EXPERIMENTAL=1 (the variant is read once per process: run this file in its own pytest process)."""
import os

import numpy as np
import pytest

from test_gpu_parity import APPLY_TOL, FIELD_TOL, build, rel, rnd, smooth_state

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("JOTS_TEST_EXPERIMENTAL") != "1", reason="experimental kernel: set JOTS_TEST_EXPERIMENTAL=1")]


@pytest.fixture(scope="module")
def J():
    os.environ["JOTS_SLAB_VARIANT"] = "6"
    import jots_b200
    if jots_b200.load_library().jots_device_count() < 1:
        pytest.fail("no CUDA device")
    return jots_b200


@pytest.mark.parametrize("props", ["const", "k", "k2", "kc"])
@pytest.mark.parametrize("sim", ["Linearized_Unsteady", "Nonlinear_Unsteady"])
@pytest.mark.parametrize("n", [(5, 3, 4), (16, 16, 9)])
def test_warp_specialised_kernel_forms(J, tmp_path, sim, props, n):
    """Same checks as test_axis_aligned_fast_path on Q2 bricks: a few tiles, and enough tiles that every block loops."""
    dev, orc = build(J, tmp_path, sim, 3, 2, props, grade=0.6, shear=0.0, n=n)
    nd = dev.n
    state = smooth_state(orc)
    x = rnd(nd, 11)
    it = orc.it
    if sim == "Linearized_Unsteady":
        orc.u = state.copy()
        orc.update_mat_props(True)
        for form, mat in ((J.FORM_MASS, it.M_mat), (J.FORM_STIFFNESS, it.K_mat), (J.FORM_SYSTEM, it.A_mat)):
            y = dev.op.form_apply(form, x, np.empty(nd), state=state)
            assert rel(y, mat @ x) < APPLY_TOL, (form, rel(y, mat @ x))
    else:
        k = 50.0 * rnd(nd, 12)
        it.u_n = state
        it.new_state(k)
        r = dev.op.form_apply(J.FORM_RESIDUAL, k, np.empty(nd), state=state)
        assert rel(r, it.mult(k)) < APPLY_TOL
        y = dev.op.form_apply(J.FORM_SYSTEM, x, np.empty(nd), state=state + it.dt * k)
        assert rel(y, it.gradient(k) @ x) < APPLY_TOL


def test_warp_specialised_kernel_time_steps(J, tmp_path):
    dev, orc = build(J, tmp_path, "Nonlinear_Unsteady", 3, 2, "k", grade=0.0, shear=0.0, n=(12, 5, 4), solver="CG", prec="Jacobi", steps=2)
    assert dev.run() == 2
    assert rel(dev.get_u(), orc.run()) < FIELD_TOL
