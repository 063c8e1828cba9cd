"""Test inputs: the reference's regression configurations restated as .ini text.

The key set / values follow /root/reference/tests/*/*.ini.in (Reinert_B1-B3 linearized
and nonlinear, Danish_Model1 (+_Neumann)).  /root/reference does not exist on the GPU box,
so the meshes are re-generated: Reinert_1D_block.mesh and Danish_1D_bar.mesh are 99x1x1
bricks; when the reference tree is present (CPU container) the oracle
test additionally runs on the shipped mesh files.
"""
import os

REF = "/root/reference"

_COMMON_TAIL = """
[Output]
Restart_Freq=0
Visualization_Freq=0
"""


def ini_text(sim, mesh, t0, props, bcs, time=True, newton=None, solver="FGMRES", prec="Chebyshev",
         steps=200, dt="1e-2", order=2, scheme="Euler_Implicit", refine=0, ls_print="Errors", newton_print="Errors", print_freq=1000000):
    """ls_print / newton_print / print_freq: the reference's defaults ("Errors, Warnings", "Errors, Warnings, Iterations", 1)
    would print a line per Newton iteration and per step; the test inputs keep the logs short."""
    s = "[FiniteElementSetup]\nSimulation_Type=%s\nFE_Order=%d\nUse_Restart=No\nMesh_File=%s\n" % (sim, order, mesh)
    s += "Serial_Refine=%d\nParallel_Refine=0\nInitial_Temperature=%s\n\n[MaterialProperties]\n" % (refine, t0)
    for k, v in props:
        s += "%s=%s\n" % (k, v)
    s += "\n[BoundaryConditions]\n"
    for i, v in enumerate(bcs):
        s += "Boundary_Attr_%d=%s\n" % (i + 1, v)
    if time:
        s += "\n[TimeIntegration]\nTime_Scheme=%s\nDelta_Time=%s\nMax_Timesteps=%d\nPrint_Freq=%d\n" % (scheme, dt, steps, print_freq)
    s += "\n[LinearSolverSettings]\nSolver=%s\nPreconditioner=%s\nMax_Iterations=100\n" % (solver, prec)
    s += "Absolute_Tolerance=1e-16\nRelative_Tolerance=1e-10\nPrint_Level=%s\n" % ls_print
    if newton is not None:
        s += "\n[NewtonSolverSettings]\nMax_Iterations=%d\nAbsolute_Tolerance=1e-16\nRelative_Tolerance=1e-10\nPrint_Level=%s\n" % (newton, newton_print)
    return s + _COMMON_TAIL


_CONST = [("Density_rho", "Uniform, 8000"), ("Specific_Heat_C", "Uniform, 500"), ("Thermal_Conductivity_k", "Uniform, 10")]
_B3 = [("Density_rho", "Uniform, 8000"), ("Specific_Heat_C", "Polynomial, 4.5, -850"),
       ("Thermal_Conductivity_k", "Polynomial, 0.09, -17")]
_HF0 = "HeatFlux, 0"
_BC_B1 = [_HF0, _HF0, _HF0, "Isothermal, 500", _HF0, _HF0]
_BC_B2 = ["HeatFlux, 7.5e5", _HF0, _HF0, _HF0, _HF0, _HF0]

# name -> dict(sim, mesh key, analytic name, eps, ini builder)
CASES = {
    "Reinert_B1": dict(mesh="reinert", exact="reinert_b1", eps=1e-8,
                       ini=lambda m, **kw: ini_text("Linearized_Unsteady", m, 300, _CONST, _BC_B1, **kw)),
    "Reinert_B2": dict(mesh="reinert", exact="reinert_b2", eps=1e-8,
                       ini=lambda m, **kw: ini_text("Linearized_Unsteady", m, 300, _CONST, _BC_B2, **kw)),
    "Reinert_B3": dict(mesh="reinert", exact="reinert_b3", eps=1e-8,
                       ini=lambda m, **kw: ini_text("Linearized_Unsteady", m, 300, _B3, _BC_B2, **kw)),
    "NL_Reinert_B1": dict(mesh="reinert", exact="reinert_b1", eps=1e-8,
                          ini=lambda m, **kw: ini_text("Nonlinear_Unsteady", m, 300, _CONST, _BC_B1, newton=100, **kw)),
    "NL_Reinert_B2": dict(mesh="reinert", exact="reinert_b2", eps=1e-8,
                          ini=lambda m, **kw: ini_text("Nonlinear_Unsteady", m, 300, _CONST, _BC_B2, newton=100, **kw)),
    "NL_Reinert_B3": dict(mesh="reinert", exact="reinert_b3", eps=1e-8,
                          ini=lambda m, **kw: ini_text("Nonlinear_Unsteady", m, 300, _B3, _BC_B2, newton=100, **kw)),
    "Danish_Model1": dict(mesh="danish", exact="danish_model1", eps=3e-5,
                          ini=lambda m, **kw: ini_text("Steady", m, 0.5, [("Thermal_Conductivity_k", "Polynomial, 50, 1")],
                                                   ["Isothermal, 0", _HF0, _HF0, "Isothermal, 1", _HF0, _HF0],
                                                   time=False, newton=5000, solver="CG", **kw)),
    "Danish_Model1_Neumann": dict(mesh="danish", exact="danish_model1", eps=3e-5,
                                  ini=lambda m, **kw: ini_text("Steady", m, 0.5, [("Thermal_Conductivity_k", "Polynomial, 50, 1")],
                                                           ["Isothermal, 0", _HF0, _HF0, "HeatFlux, 26", _HF0, _HF0],
                                                           time=False, newton=5000, solver="CG", **kw)),
}

# brick equivalents of the shipped meshes: (nx, ny, nz, lx, ly, lz)
BRICKS = {"reinert": (99, 1, 1, 0.01, 0.01, 0.01), "danish": (99, 1, 1, 1.0, 0.1, 0.1)}
REF_MESH = {"reinert": "examples/meshes/Reinert_1D_block.mesh", "danish": "examples/meshes/Danish_1D_bar.mesh"}


def _grade(t, g):
    """monotone map of [0,1] onto itself; g = 0 is the identity"""
    return t + g * t * (1.0 - t)


def _jit(i, j, k, comp, amp):
    """deterministic pseudo-random offset in [-amp, amp]"""
    h = (i * 73856093) ^ (j * 19349663) ^ (k * 83492791) ^ (comp * 2654435761)
    h = (h * 0x9E3779B1) & 0xFFFFFFFF
    return amp * (2.0 * (h / 4294967295.0) - 1.0)


def write_brick_mesh(path, nx, ny, nz, lx, ly, lz, grade=0.0, shear=0.0, jitter=0.0):
    """Write a brick in 'MFEM mesh v1.0' format with the attribute convention of
    examples/meshes/cube_25_L0p01.mesh (1=x0 2=y0 3=z0 4=x1 5=y1 6=z1).  grade != 0 makes the
    element sizes vary (still affine per element); shear != 0 applies a global affine shear."""
    vx, vy = nx + 1, ny + 1
    vid = lambda i, j, k: i + vx * (j + vy * k)
    out = ["MFEM mesh v1.0", "", "dimension", "3", "", "elements", str(nx * ny * nz)]
    for k in range(nz):
        for j in range(ny):
            for i in range(nx):
                v = [vid(i, j, k), vid(i + 1, j, k), vid(i + 1, j + 1, k), vid(i, j + 1, k),
                     vid(i, j, k + 1), vid(i + 1, j, k + 1), vid(i + 1, j + 1, k + 1), vid(i, j + 1, k + 1)]
                out.append("1 5 " + " ".join(map(str, v)))
    b = []
    for k in range(nz):
        for j in range(ny):
            b.append((1, [vid(0, j, k), vid(0, j, k + 1), vid(0, j + 1, k + 1), vid(0, j + 1, k)]))
            b.append((4, [vid(nx, j, k), vid(nx, j + 1, k), vid(nx, j + 1, k + 1), vid(nx, j, k + 1)]))
    for k in range(nz):
        for i in range(nx):
            b.append((2, [vid(i, 0, k), vid(i + 1, 0, k), vid(i + 1, 0, k + 1), vid(i, 0, k + 1)]))
            b.append((5, [vid(i, ny, k), vid(i, ny, k + 1), vid(i + 1, ny, k + 1), vid(i + 1, ny, k)]))
    for j in range(ny):
        for i in range(nx):
            b.append((3, [vid(i, j, 0), vid(i, j + 1, 0), vid(i + 1, j + 1, 0), vid(i + 1, j, 0)]))
            b.append((6, [vid(i, j, nz), vid(i + 1, j, nz), vid(i + 1, j + 1, nz), vid(i, j + 1, nz)]))
    out += ["", "boundary", str(len(b))]
    out += ["%d 3 %s" % (a, " ".join(map(str, v))) for a, v in b]
    out += ["", "vertices", str(vx * vy * (nz + 1)), "3"]
    for k in range(nz + 1):
        for j in range(ny + 1):
            for i in range(nx + 1):
                x, y, z = lx * _grade(i / nx, grade), ly * _grade(j / ny, -grade), lz * _grade(k / nz, 0.5 * grade)
                if jitter and 0 < i < nx and 0 < j < ny and 0 < k < nz:   # interior vertices: general trilinear hexes
                    x += _jit(i, j, k, 0, jitter * lx / nx); y += _jit(i, j, k, 1, jitter * ly / ny); z += _jit(i, j, k, 2, jitter * lz / nz)
                out.append("%.17g %.17g %.17g" % (x + shear * y + 0.5 * shear * z, y + 0.25 * shear * z, z))
    with open(path, "w") as f:
        f.write("\n".join(out) + "\n")


def write_quad_mesh(path, nx, ny, lx, ly, y0=0.0, grade=0.0, shear=0.0, attrs=(1, 2, 3, 4), jitter=0.0):
    """2-D rectangle; attrs = attributes of (bottom, right, top, left): (1, 2, 3, 4) is the plate.mesh
    convention, (1, 3, 1, 2) the bar.mesh one (1 = long sides, 2 = x0, 3 = x1)."""
    vx = nx + 1
    vid = lambda i, j: i + vx * j
    out = ["MFEM mesh v1.0", "", "dimension", "2", "", "elements", str(nx * ny)]
    for j in range(ny):
        for i in range(nx):
            out.append("1 3 %d %d %d %d" % (vid(i, j), vid(i + 1, j), vid(i + 1, j + 1), vid(i, j + 1)))
    b = []
    for i in range(nx):
        b.append((attrs[0], [vid(i, 0), vid(i + 1, 0)]))
        b.append((attrs[2], [vid(i + 1, ny), vid(i, ny)]))
    for j in range(ny):
        b.append((attrs[1], [vid(nx, j), vid(nx, j + 1)]))
        b.append((attrs[3], [vid(0, j + 1), vid(0, j)]))
    out += ["", "boundary", str(len(b))]
    out += ["%d 1 %d %d" % (a, v[0], v[1]) for a, v in b]
    out += ["", "vertices", str(vx * (ny + 1)), "2"]
    for j in range(ny + 1):
        for i in range(nx + 1):
            x, y = lx * _grade(i / nx, grade), ly * _grade(j / ny, -grade)
            if jitter and 0 < i < nx and 0 < j < ny:
                x += _jit(i, j, 0, 0, jitter * lx / nx); y += _jit(i, j, 0, 1, jitter * ly / ny)
            out.append("%.17g %.17g" % (x + shear * y, y0 + y))
    with open(path, "w") as f:
        f.write("\n".join(out) + "\n")


def make_case(name, tmpdir, use_reference_mesh=False, **kw):
    """Write mesh + ini for a named case into tmpdir; returns the ini path."""
    c = CASES[name]
    if use_reference_mesh:
        mesh = os.path.join(REF, REF_MESH[c["mesh"]])
    else:
        mesh = os.path.join(str(tmpdir), c["mesh"] + ".mesh")
        if not os.path.exists(mesh):
            write_brick_mesh(mesh, *BRICKS[c["mesh"]])
    ini = os.path.join(str(tmpdir), name + ".ini")
    with open(ini, "w") as f:
        f.write(c["ini"](mesh, **kw))
    return ini
