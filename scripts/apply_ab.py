#!/usr/bin/env python
"""A/B timing of the operator applies on the BASELINE configs[3] brick (128^3 Q2 hexes) for several kernel
configurations, one subprocess each (the kernel variant is read from the environment when a context is built).
Prints one JSON line per configuration: CUDA-event time of the Jacobian / residual / linear applies and, for the
first configuration, nothing else -- every later line also carries the difference of its results to the first."""
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CONFIGS = {
    "patch2_barrier": {"JOTS_SLAB_VARIANT": "7", "JOTS_PATCH2_LAG": "0"},          # two-phase patch kernel, one block barrier per tile
    "patch2_barrier_dot": {"JOTS_SLAB_VARIANT": "7", "JOTS_PATCH2_LAG": "0", "JOTS_FORM_TIME_DOT": "1"},
    "slab3_gather_E16": {"JOTS_SLAB_VARIANT": "3"},                   # gather-table kernel (round-1 default)
    # gather-table kernel + (x, A x) in phase 3 / + the CG loop's side stream (two staging buffers: Q1, Q3 -- AB_ORDER=3 AB_N=84)
    "slab3_dot": {"JOTS_SLAB_VARIANT": "3", "JOTS_FORM_TIME_DOT": "1"},
    "slab3_dot_side": {"JOTS_SLAB_VARIANT": "3", "JOTS_FORM_TIME_DOT": "2"},
    "patch3_4x4": {"JOTS_SLAB_VARIANT": "7", "JOTS_PATCH2": "0"},     # three-phase unique-pencil patch kernel
    "patch2_4x4": {"JOTS_SLAB_VARIANT": "7"},                         # two-phase patch kernel, warp-decoupled pipeline (default)
    "patch2_4x4_dot": {"JOTS_SLAB_VARIANT": "7", "JOTS_FORM_TIME_DOT": "1"},   # the same + (x, A x) formed in phase 3 (CG loop)
    # the same + the CG loop's side stream (x_sol += alpha d_old, zeroing of the next output) instead of the zeroing pass
    "patch2_4x4_dot_side": {"JOTS_SLAB_VARIANT": "7", "JOTS_FORM_TIME_DOT": "2"},
}


def child(name, n, order, reps, variant):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import numpy as np
    import torch
    import jots_b200 as J
    import bench
    bench._variant = variant
    with tempfile.TemporaryDirectory() as d:
        ini = os.path.join(d, "c.ini")
        bench.write_ini(ini, order, "CG", "Jacobi", 1)
        drv = J.Driver(ini, 0, mesh=J.Mesh.brick(n, n, n, 0.01, 0.01, 0.01))
    ndof = drv.n
    g = torch.Generator(device="cuda").manual_seed(12345)
    x = torch.empty(ndof, dtype=torch.float64, device="cuda").uniform_(-1.0, 1.0, generator=g)
    X = torch.tensor(drv.space().node_coordinates(), device="cuda")
    z = 300.0 + 100.0 * torch.prod(torch.sin(torch.pi * X / 0.01), dim=1)
    y = torch.empty_like(x)
    torch.cuda.synchronize()
    out = {"config": name, "n_dof": ndof, "order": order, "variant": variant}
    for key, form in (("jac", J.FORM_SYSTEM), ("res", J.FORM_RESIDUAL)):
        drv.op.form_time(form, x, y, state_dev=z, reps=10)
        out[key + "_ms"] = drv.op.form_time(form, x, y, state_dev=z, reps=reps)
        drv.op.form_apply(form, x, y, state=z)
        torch.cuda.synchronize()
        out[key + "_norm"] = float(torch.linalg.vector_norm(y))
        if os.environ.get("AB_DIR"):      # absent when the child is run on its own (ncu captures)
            np.save(os.path.join(os.environ["AB_DIR"], "%s_%s.npy" % (name, key)), y[::97].cpu().numpy())
    out["jac_gdofs"] = ndof / out["jac_ms"] / 1e6
    out["kernel"] = drv.ctx.last_kernel()
    out["patch_applies"] = drv.ctx.stats()["patch_applies"]
    print(json.dumps(out), flush=True)


def main():
    import numpy as np
    n = int(os.environ.get("AB_N", "128"))
    order = int(os.environ.get("AB_ORDER", "2"))
    reps = int(os.environ.get("AB_REPS", "50"))
    variant = os.environ.get("AB_VARIANT", "A")
    names = sys.argv[1:] or list(CONFIGS)
    with tempfile.TemporaryDirectory() as d:
        first = None
        for name in names:
            env = dict(os.environ, AB_DIR=d, **CONFIGS[name])
            r = subprocess.run([sys.executable, __file__, "--child", name, str(n), str(order), str(reps), variant], env=env,
                               capture_output=True, text=True, timeout=300)
            sys.stdout.write(r.stdout)
            if r.returncode != 0:
                sys.stdout.write(json.dumps({"config": name, "failed": r.stderr[-800:]}) + "\n")
                continue
            if first is None:
                first = name
            else:
                for key in ("jac", "res"):
                    a = np.load(os.path.join(d, "%s_%s.npy" % (first, key)))
                    b = np.load(os.path.join(d, "%s_%s.npy" % (name, key)))
                    print(json.dumps({"config": name, "vs": first, key + "_rel_diff": float(np.linalg.norm(a - b) / np.linalg.norm(a))}), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--child":
        child(sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), sys.argv[6])
    else:
        main()
