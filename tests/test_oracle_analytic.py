"""Pins the CPU oracle on every regression the reference ships for the conduction solve
(/root/reference/tests/{linearized_unsteady_heat,nonlinear_unsteady_heat,steady_heat}/*.cpp:5)
with the reference's own error functional (tests/test_helper.hpp.in:134)."""
import os

import numpy as np
import pytest

from oracle import analytic, jots as OJ
from cases import CASES, REF, make_case


def _run(name, tmp_path, use_ref):
    ini = make_case(name, tmp_path, use_reference_mesh=use_ref)
    cfg = OJ.Config(ini)
    drv = OJ.JOTSDriver(cfg)
    u = drv.run()
    t_end = cfg.max_timesteps * cfg.dt if cfg.using_time_integration else 0.0
    err = analytic.reference_error(drv.space, u, getattr(analytic, CASES[name]["exact"]), t_end)
    return err, drv


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_matches_reference_analytic_pin(name, tmp_path):
    err, drv = _run(name, tmp_path, use_ref=False)
    assert np.isfinite(err) and err <= CASES[name]["eps"], (name, err)


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present (GPU box)")
@pytest.mark.parametrize("name", ["Reinert_B2", "NL_Reinert_B2", "Danish_Model1_Neumann"])
def test_oracle_on_shipped_mesh_files(name, tmp_path):
    """Same pins on the reference's own mesh files (Pointwise vertex order, un-structured numbering)."""
    err, drv = _run(name, tmp_path, use_ref=True)
    assert np.isfinite(err) and err <= CASES[name]["eps"], (name, err)


@pytest.mark.parametrize("name", ["Reinert_B1", "Danish_Model1_Neumann"])
def test_committed_golden_fields_are_the_oracle_output(name, tmp_path):
    """tests/golden/regressions.npz (the fixture the GPU tests compare against) is reproducible."""
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "regressions.npz"))
    ini = make_case(name, tmp_path)
    u = OJ.JOTSDriver(OJ.Config(ini)).run()
    assert np.linalg.norm(u - gold[name]) <= 1e-12 * np.linalg.norm(gold[name])
