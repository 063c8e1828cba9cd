"""Generates tests/golden/regressions.npz: the CPU oracle's final temperature fields for the eight
regression configurations of the reference (tests/cases.py restates their .ini files on regenerated
99x1x1 brick meshes).  The reference ships no golden vectors of its own (its tests compare against
analytic solutions) and cannot be run here, so these are oracle outputs, pinned separately on those
analytic solutions by tests/test_oracle_analytic.py.

    python tests/golden/make_golden.py
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from cases import CASES, make_case  # noqa: E402
from oracle import jots as OJ  # noqa: E402


def main():
    out = {}
    with tempfile.TemporaryDirectory() as d:
        for name in sorted(CASES):
            ini = make_case(name, d)
            drv = OJ.JOTSDriver(OJ.Config(ini))
            out[name] = drv.run()
            st = drv.it.stats
            out[name + "__stats"] = np.array([st.get("steps", 0), st.get("newton_iters", 0), st.get("krylov_iters", 0)])
            print(name, out[name].shape, st)
    np.savez_compressed(os.path.join(HERE, "regressions.npz"), **out)


if __name__ == "__main__":
    main()
