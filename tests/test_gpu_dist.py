"""Multi-GPU parity (needs >= 2 GPUs; skipped otherwise): the partitioned device solve -- one
subdomain per GPU, NCCL interface exchange -- against the single-domain oracle."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _ngpu():
    import jots_b200
    return jots_b200.load_library().jots_device_count()


def _guard(fn):
    """A worker that dies must not leave the parent waiting for the queue time-out (GPU minutes): report the error."""
    import functools

    @functools.wraps(fn)
    def run(*args):
        q = args[-1]
        try:
            fn(*args)
        except BaseException as e:      # noqa: BLE001
            import traceback
            q.put(("error", traceback.format_exc() + repr(e)))
    return run


def _collect(q, world):
    res = []
    for _ in range(world):
        r = q.get(timeout=300)
        if r[0] == "error":
            pytest.fail("worker failed:\n" + r[1])
        res.append(r)
    return res


def _worker_impl(rank, world, ini, shape, uid, q):
    import jots_b200 as J
    comm = J.Comm(uid, rank, world, rank)
    mesh = J.Mesh.brick(*shape) if shape is not None else None   # None: every rank reads Mesh_File
    drv = J.Driver(ini, rank, mesh=mesh, comm=comm)
    steps = drv.run()
    g, n_owned = drv.global_ids()
    q.put((rank, steps, g, n_owned, drv.get_u(), drv.ctx.stats()))
    del drv, comm


def _worker(*args):
    _guard(_worker_impl)(*args)


CASES = [
    ("Nonlinear_Unsteady", "k", "FGMRES", "Chebyshev", 2),
    ("Nonlinear_Unsteady", "full", "GMRES", "Jacobi", 2),
    ("Linearized_Unsteady", "full", "CG", "Chebyshev", 2),
    ("Steady", "k", "CG", "Jacobi", 2),
]


@pytest.mark.parametrize("sim,props,solver,prec,p", CASES)
def test_partitioned_solve_matches_oracle(tmp_path, sim, props, solver, prec, p):
    world = min(_ngpu(), 4)
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    import torch.multiprocessing as mp
    import jots_b200 as J
    from oracle import jots as OJ, mesh as OM
    from cases import ini_text
    from test_gpu_parity import PROPS, HF0
    shape = (6, 4, 4, 0.012, 0.01, 0.008)
    bcs = ["HeatFlux, 7.5e5", "Isothermal, 350", HF0, "Sinusoidal_Isothermal, 50, 40, 0.3, 400", HF0, "HeatFlux, -1e5"]
    if sim == "Steady":
        bcs = ["Isothermal, 300", HF0, HF0, "HeatFlux, 2e4", HF0, HF0]
    txt = ini_text(sim, "unused", 300, PROPS[props] if sim != "Steady" else [PROPS[props][2]], bcs, time=(sim != "Steady"),
                   newton=None if sim == "Linearized_Unsteady" else 50, solver=solver, prec=prec, steps=3, order=p)
    ini = str(tmp_path / "c.ini")
    with open(ini, "w") as f:
        f.write(txt)
    uid = J.Comm.unique_id()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, ini, shape, uid, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    res = _collect(q, world)
    for pr in procs:
        pr.join(timeout=60)
    orc = OJ.JOTSDriver(OJ.Config(ini), mesh=OM.make_brick(*shape))
    u_ref = orc.run()
    u = np.full(u_ref.size, np.nan)
    owned = np.zeros(u_ref.size, dtype=int)
    for rank, steps, g, n_owned, ul, st in res:
        owned[g[:n_owned]] += 1
        # shared copies agree with the owner's value
        u[g[:n_owned]] = ul[:n_owned]
    assert (owned == 1).all()
    for rank, steps, g, n_owned, ul, st in res:
        assert np.abs(ul - u[g]).max() <= 1e-9 * np.abs(u).max()
    err = np.linalg.norm(u - u_ref) / np.linalg.norm(u_ref)
    assert err < 1e-8, err


def test_partitioned_solve_on_a_mesh_file_rcb(tmp_path, monkeypatch):
    """Unstructured path on several GPUs: mesh file -> first-appearance numbering -> recursive coordinate
    bisection -> NCCL interface exchange, against the single-domain oracle on the same file.  The run also writes ParaView
    output every step (rank 0 collects the owned values): the last cycle holds the global field and the owner of each node."""
    world = min(_ngpu(), 4)
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    import torch.multiprocessing as mp
    import jots_b200 as J
    from oracle import jots as OJ
    from cases import ini_text, write_brick_mesh
    from test_gpu_parity import PROPS, HF0
    mesh = str(tmp_path / "m.mesh")
    write_brick_mesh(mesh, 6, 4, 3, 0.012, 0.01, 0.008, grade=0.4, shear=0.2, jitter=0.15)
    bcs = ["HeatFlux, 7.5e5", "Isothermal, 350", HF0, "Isothermal, 320", HF0, "HeatFlux, -1e5"]
    ini = str(tmp_path / "c.ini")
    with open(ini, "w") as f:
        f.write(ini_text("Nonlinear_Unsteady", mesh, 300, PROPS["kc"], bcs, newton=50, solver="FGMRES", prec="Chebyshev", steps=2, order=2)
                .replace("Visualization_Freq=0", "Visualization_Freq=1"))
    monkeypatch.setenv("JOTS_PARAVIEW_DIR", str(tmp_path / "ParaView"))       # inherited by the spawned ranks
    uid = J.Comm.unique_id()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, ini, None, uid, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    res = _collect(q, world)
    for pr in procs:
        pr.join(timeout=60)
    u_ref = OJ.JOTSDriver(OJ.Config(ini)).run()
    u = np.full(u_ref.size, np.nan)
    for rank, steps, g, n_owned, ul, st in res:
        u[g[:n_owned]] = ul[:n_owned]
    assert not np.isnan(u).any()
    assert np.linalg.norm(u - u_ref) / np.linalg.norm(u_ref) < 1e-8
    from oracle import vtkfmt
    out = str(tmp_path / "ParaView")
    assert [f for _, f in vtkfmt.read_pvd(out + "/ParaView.pvd")] == ["Cycle%06d/data.pvtu" % c for c in (0, 1, 2)]
    v = vtkfmt.read_vtu(out + "/Cycle000002/proc000000.vtu")
    g = J.Space(J.Mesh.read(mesh), 2).gather()
    assert np.array_equal(v["point_data"]["Temperature"].reshape(-1, 27), u[g])
    owner = np.full(u_ref.size, -1.0)
    for rank, steps, gl, n_owned, ul, st in res:
        owner[gl[:n_owned]] = rank
    assert np.array_equal(v["point_data"]["Rank"].reshape(-1, 27), owner[g])


def _worker_tables(*args):
    _guard(_worker_tables_impl)(*args)


def _worker_tables_impl(rank, world, shape, uid, q):
    """One rank of the binding a JOTS maintainer would write: the subdomain comes from the CALLER's decomposition tables
    (scrambled L-vector numbering, true-DOF vectors), the step loop is JOTSDriver::Run's order of calls on the C ABI."""
    import jots_b200 as J
    from test_host_cpu import _caller_tables
    comm = J.Comm(uid, rank, world, rank)
    mesh = J.Mesh.brick(*shape)
    ref = J.Partition(mesh, J.Space(mesh, 2), world, rank)
    part = J.Partition.from_tables(rank=rank, nranks=world, attributes=[1, 2, 3, 4, 5, 6], **_caller_tables(ref, 7 + rank))
    ctx = J.Context.from_partition(part, rank, comm)
    ctx.set_material(0, ("Uniform", 8000))
    ctx.set_material(1, ("Polynomial", 4.5, -850))
    ctx.set_material(2, ("Polynomial", 0.09, -17))
    ctx.set_bcs([("HeatFlux", 7.5e5), ("Isothermal", 350), ("HeatFlux", 0), ("Sinusoidal_Isothermal", 50, 40, 0.3, 400), ("HeatFlux", 0), ("HeatFlux", -1e5)])
    u = ctx.expand_true(np.full(part.n_owned, 300.0))           # true-DOF vector -> local vector
    ctx.update_bcs(0.0)
    ctx.apply_dirichlet(u)
    ctx.set_material_state(u)
    op = J.Operator(ctx, J.NONLINEAR_UNSTEADY, dt=1e-2, solver=J.FGMRES, prec=J.CHEBYSHEV, newton_max_iter=50)
    t = 0.0
    for step in range(3):
        ctx.update_bcs(t)
        ctx.apply_dirichlet(u, only_time_dependent=(step > 0))
        ctx.set_material_state(u)
        for which in (1, 2):
            op.process_mat_prop_update(which)
        t = op.iterate(u, t)
    g, _ = part.global_ids()
    q.put((rank, g[:part.n_owned], u[:part.n_owned], ctx.essential_dofs().size, ctx.stats()))
    del op, ctx, comm


def test_partitioned_solve_from_caller_supplied_tables(tmp_path):
    """jots_partition_create_from_tables + jots_ctx_create_part + the operator entry points, one rank per GPU: the field
    equals the single-domain oracle's; boundary DOF sets are completed by the marker exchange (an x1 / y0 edge DOF whose
    boundary faces belong to another rank's elements must still be essential)."""
    world = min(_ngpu(), 4)
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    import torch.multiprocessing as mp
    import jots_b200 as J
    from oracle import jots as OJ, mesh as OM
    from cases import ini_text
    from test_gpu_parity import PROPS, HF0
    shape = (6, 4, 4, 0.012, 0.01, 0.008)
    bcs = ["HeatFlux, 7.5e5", "Isothermal, 350", HF0, "Sinusoidal_Isothermal, 50, 40, 0.3, 400", HF0, "HeatFlux, -1e5"]
    props = [("Density_rho", "Uniform, 8000"), ("Specific_Heat_C", "Polynomial, 4.5, -850"), ("Thermal_Conductivity_k", "Polynomial, 0.09, -17")]
    ini = str(tmp_path / "c.ini")
    with open(ini, "w") as f:
        f.write(ini_text("Nonlinear_Unsteady", "unused", 300, props, bcs, newton=50, solver="FGMRES", prec="Chebyshev", steps=3, order=2))
    uid = J.Comm.unique_id()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_tables, args=(r, world, shape, uid, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    res = _collect(q, world)
    for pr in procs:
        pr.join(timeout=60)
    orc = OJ.JOTSDriver(OJ.Config(ini), mesh=OM.make_brick(*shape))
    u_ref = orc.run()
    u = np.full(u_ref.size, np.nan)
    for rank, g, ul, ness, st in res:
        u[g] = ul
    assert not np.isnan(u).any()
    assert np.linalg.norm(u - u_ref) / np.linalg.norm(u_ref) < 1e-8
    assert sum(st["newton_iters"] for _, _, _, _, st in res) == world * orc.it.stats["newton_iters"]
