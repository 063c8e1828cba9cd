#!/usr/bin/env python
"""The shipped (tiny) configurations of BASELINE.json configs[0..2] -- the reference's regression cases on 99-hex meshes,
1 791 DOFs at Q2 -- on the GPU next to the CPU oracle (NumPy/SciPy restatement: sparse assembly + SpMV Krylov, one thread;
the reference itself cannot be built here).  These cases are launch-latency bound on a GPU: the numbers are reported so
that the regime is visible, not because it is where a B200 belongs.  One JSON line per configuration."""
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import jots_b200 as J  # noqa: E402
from cases import make_case  # noqa: E402
from oracle import jots as OJ  # noqa: E402

names = sys.argv[1:] or ["Danish_Model1", "Reinert_B1", "Reinert_B3", "NL_Reinert_B2", "NL_Reinert_B3"]
d = tempfile.mkdtemp()
for name in names:
    ini = make_case(name, d)
    J.Driver(ini).run(2)                      # warm-up: module load, first launches
    drv = J.Driver(ini)
    t0 = time.perf_counter()
    n = drv.run()
    dt = time.perf_counter() - t0
    st = drv.ctx.stats()
    orc = OJ.JOTSDriver(OJ.Config(ini))
    t0 = time.perf_counter()
    orc.run()
    dt_cpu = time.perf_counter() - t0
    so = orc.it.stats
    print(json.dumps({"config": name, "n_dof": drv.n, "steps": n, "gpu_s": dt, "gpu_steps_per_s": n / dt,
                      "krylov_iters": st["krylov_iters"], "newton_iters": st["newton_iters"], "gpu_launches": st["kernel_launches"],
                      "us_per_launch": 1e6 * dt / max(st["kernel_launches"], 1),
                      "cpu_oracle_s": dt_cpu, "cpu_oracle_steps_per_s": so["steps"] / dt_cpu, "cpu_oracle_krylov_iters": so["krylov_iters"],
                      "cpu_kind": "port (NumPy/SciPy oracle, 1 thread)"}), flush=True)
