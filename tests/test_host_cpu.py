"""CPU-side tests (no GPU): the C-ABI library loads and exports every declared symbol, the host-side
mesh / space / partition index sets are bit-exact against the oracle, and errors are reported through
the status/last_error convention instead of aborting."""
import os
import re

import numpy as np
import pytest

import jots_b200 as J
from cases import REF, write_brick_mesh, write_quad_mesh

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    lib = J.load_library()
    hdr = open(os.path.join(ROOT, "include", "jots_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(jots_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) > 50
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, missing
    assert lib.jots_version().startswith(b"jots_b200")


def test_no_cpu_fallback_when_no_device():
    lib = J.load_library()
    if lib.jots_device_count() > 0:
        pytest.skip("a CUDA device is present")
    sp = J.Space(J.Mesh.brick(2, 2, 2, 1, 1, 1), 2)
    with pytest.raises(J.JotsError, match="no CUDA device"):
        J.Context(sp, 0)


@pytest.mark.parametrize("p", [1, 2, 3, 4])
def test_brick_space_index_tables_bit_exact(p):
    from oracle import fem, mesh as OM
    m = J.Mesh.brick(4, 3, 2, 1.0, 0.5, 0.25)
    sp = J.Space(m, p)
    osp = fem.H1Space(OM.make_brick(4, 3, 2, 1.0, 0.5, 0.25), p)
    assert sp.n_dof == osp.ndof
    assert (sp.gather() == osp.gather).all()
    bg, ba, be, bf = sp.boundary()
    assert (bg == osp.bdr_gather).all() and (ba == osp.mesh.bdr_attr).all()
    assert (be == osp.bdr_elem).all() and (bf == osp.bdr_face).all()
    assert np.abs(sp.node_coordinates() - osp.node_coordinates()).max() < 1e-15
    for attrs in ([1], [4], [1, 4], [2, 3, 6], list(range(1, 7))):
        assert (sp.essential_dofs(attrs) == osp.essential_dofs(attrs)).all()


@pytest.mark.parametrize("dim,p,refine", [(3, 2, 0), (3, 3, 1), (2, 1, 0), (2, 4, 1), (2, 2, 2)])
def test_file_mesh_numbering_and_refinement_bit_exact(tmp_path, dim, p, refine):
    """Unstructured path: mesh file -> (refine) -> first-appearance lattice numbering."""
    from oracle import fem, mesh as OM
    path = str(tmp_path / "m.mesh")
    if dim == 3:
        write_brick_mesh(path, 3, 2, 2, 1.0, 0.5, 0.25, grade=0.3)
    else:
        write_quad_mesh(path, 4, 3, 1.0, 0.25, y0=-0.25)
    m, om = J.Mesh.read(path), OM.read_mfem_mesh(path)
    for _ in range(refine):
        m, om = m.refine(), OM.uniform_refine(om)
    v, e, b, ba = m.arrays()
    assert (e == om.elems).all() and (b == om.bdr).all() and (ba == om.bdr_attr).all()
    assert np.abs(v - om.verts).max() < 1e-15
    sp, osp = J.Space(m, p), fem.H1Space(om, p)
    assert (sp.gather() == osp.gather).all() and (sp.boundary()[0] == osp.bdr_gather).all()


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present (GPU box)")
@pytest.mark.parametrize("name", ["Reinert_1D_block", "Danish_1D_bar", "bar", "plate", "cube_25_L0p01"])
def test_shipped_meshes(name):
    """Every mesh the reference ships (Pointwise vertex order; plate.mesh needs the orientation fix)."""
    from oracle import fem, mesh as OM
    path = os.path.join(REF, "examples", "meshes", name + ".mesh")
    m, om = J.Mesh.read(path), OM.read_mfem_mesh(path)
    v, e, b, ba = m.arrays()
    assert (e == om.elems).all() and (b == om.bdr).all()
    if name != "cube_25_L0p01":
        sp, osp = J.Space(m, 2), fem.H1Space(om, 2)
        assert (sp.gather() == osp.gather).all() and (sp.boundary()[0] == osp.bdr_gather).all()
        assert (sp.essential_dofs([1]) == osp.essential_dofs([1])).all()


def test_space_from_caller_tables_round_trip():
    """jots_space_create_from_tables: the drop-in entry for a caller that already holds MFEM's tables."""
    sp = J.Space(J.Mesh.brick(3, 2, 2, 1.0, 0.5, 0.25), 2)
    ec, bc = sp.corners()
    bg, ba, _, _ = sp.boundary()
    sp2 = J.Space.from_tables(3, 2, sp.n_dof, sp.gather(), ec, bg, bc, ba)
    assert (sp2.n_dof, sp2.n_elem, sp2.n_bdr) == (sp.n_dof, sp.n_elem, sp.n_bdr)
    assert (sp2.gather() == sp.gather()).all()
    assert (sp2.essential_dofs([1, 4]) == sp.essential_dofs([1, 4])).all()
    assert np.abs(sp2.node_coordinates() - sp.node_coordinates()).max() < 1e-15
    with pytest.raises(J.JotsError, match="out of range"):
        J.Space.from_tables(3, 2, 5, sp.gather(), ec, bg, bc, ba)


@pytest.mark.parametrize("dim,p", [(3, 1), (3, 2), (3, 3), (2, 2), (2, 3)])
def test_boundary_adjacency_recovered_from_the_dof_ids(tmp_path, dim, p):
    """jots_space_boundary_adjacency (what the wall heat flux of the preCICE write path needs): a space built from the
    caller's tables carries no boundary-element adjacency and may list the DOFs of a boundary face in an order of its own
    (here: transposed and reversed).  The adjacent element and its face must be the mesh's own, and every node must be
    placed where the mesh-built space places it -- integer compare."""
    mesh = str(tmp_path / "m.mesh")
    if dim == 3:
        write_brick_mesh(mesh, 4, 3, 2, 0.01, 0.02, 0.015, grade=0.4)
    else:
        write_quad_mesh(mesh, 6, 4, 0.02, 0.01, grade=0.4)
    sp = J.Space(J.Mesh.read(mesh), p)
    ec, bc = sp.corners()
    bg, ba, be, bf = sp.boundary()
    D = p + 1
    nfl = D ** (dim - 1)
    if dim == 3:
        perm = np.arange(nfl).reshape(D, D).T[::-1, ::-1].ravel()
        cperm = np.arange(4).reshape(2, 2).T[::-1, ::-1].ravel()
    else:
        perm = np.arange(nfl)[::-1]
        cperm = np.arange(2)[::-1]
    bg2 = np.ascontiguousarray(bg.reshape(-1, nfl)[:, perm])
    bc2 = np.ascontiguousarray(bc.reshape(bg2.shape[0], 2 ** (dim - 1), dim)[:, cperm, :])
    sp2 = J.Space.from_tables(dim, p, sp.n_dof, sp.gather(), ec, bg2, bc2, ba)
    e1, f1, nl1 = sp.boundary_adjacency()
    e2, f2, nl2 = sp2.boundary_adjacency()
    assert (e1 == be).all() and (f1 == bf).all()
    assert (e2 == be).all() and (f2 == bf).all()
    assert (nl2 == nl1[:, perm]).all()
    g = sp.gather()
    assert (g[e2[:, None], nl2] == bg2).all()          # the place really holds the DOF the caller listed
    # a boundary table that does not belong to the element table is an error, not a wrong flux
    bad = bg2.copy()
    bad[0, 0] = bad[-1, -1]
    with pytest.raises(J.JotsError, match="matches no element face|not on the matched element face"):
        J.Space.from_tables(dim, p, sp.n_dof, sp.gather(), ec, bad, bc2, ba).boundary_adjacency()


def _emulated_exchange(parts, vals):
    """pairwise exchange + canonical-order reduction, as halo_sum_exchange does on the device"""
    recv = []
    for p in parts:
        nb = p.neighbors()
        buf = []
        for r, idx in nb:
            q = parts[r]
            back = [i for (rr, i) in q.neighbors() if rr == p.rank][0]
            assert back.size == idx.size
            buf.append(vals[r][back])
        recv.append(np.concatenate(buf) if buf else np.zeros(0))
    out = []
    for p, v, rb in zip(parts, vals, recv):
        v = v.copy()
        dof, ptr, src = p.reduction()
        for s in range(dof.size):
            acc = 0.0
            for k in range(ptr[s], ptr[s + 1]):
                acc += v[dof[s]] if src[k] < 0 else rb[src[k]]
            v[dof[s]] = acc
        out.append(v)
    return out


@pytest.mark.parametrize("nranks,shape,p", [(2, (4, 2, 2), 2), (4, (4, 4, 2), 1), (8, (4, 4, 4), 2), (3, (5, 3, 2), 3)])
def test_partition_interface_sum_is_global_assembly(nranks, shape, p):
    """Sum-exchange of local element multiplicities reproduces the global assembly on every rank;
    ownership, neighbour lists and global ids are consistent (bit-exact integer contract)."""
    m = J.Mesh.brick(*shape, 1.0, 1.0, 1.0)
    sp = J.Space(m, p)
    gmult = np.bincount(sp.gather().ravel(), minlength=sp.n_dof).astype(float)
    parts = [J.Partition(m, sp, nranks, r) for r in range(nranks)]
    owned = np.zeros(sp.n_dof, dtype=int)
    vals, gids = [], []
    nelem = 0
    for prt in parts:
        g, el = prt.global_ids()
        gids.append(g)
        lsp = prt.space()
        assert lsp.n_dof == prt.n_local
        # local gather maps to the global gather of the same elements
        assert (g[lsp.gather()] == sp.gather()[el]).all()
        nelem += prt.n_elem
        owned[g[:prt.n_owned]] += 1
        assert (np.diff(g[:prt.n_owned]) > 0).all()
        vals.append(np.bincount(lsp.gather().ravel(), minlength=prt.n_local).astype(float) * (1.0 + 0.25 * prt.rank))
        # essential DOFs of the subdomain = global essential DOFs restricted to it
        ge = sp.essential_dofs([1, 5])
        le = lsp.essential_dofs([1, 5])
        assert (np.sort(g[le]) == np.intersect1d(ge, g)).all()
    assert nelem == sp.n_elem and (owned == 1).all()
    # reference: weighted global assembly
    ref = np.zeros(sp.n_dof)
    for prt, g in zip(parts, gids):
        _, el = prt.global_ids()
        np.add.at(ref, sp.gather()[el].ravel(), 1.0 + 0.25 * prt.rank)
    out = _emulated_exchange(parts, vals)
    for prt, g, v in zip(parts, gids, out):
        assert np.array_equal(v, ref[g]) or np.allclose(v, ref[g], rtol=1e-15, atol=0)


def test_partition_of_unstructured_mesh_rcb(tmp_path):
    path = str(tmp_path / "m.mesh")
    write_quad_mesh(path, 6, 5, 1.0, 0.5)
    m = J.Mesh.read(path).refine()
    sp = J.Space(m, 2)
    parts = [J.Partition(m, sp, 4, r) for r in range(4)]
    assert sum(p.n_elem for p in parts) == sp.n_elem
    assert sum(p.n_owned for p in parts) == sp.n_dof
    assert max(p.n_elem for p in parts) - min(p.n_elem for p in parts) <= 1


def test_errors_are_status_codes_not_aborts(tmp_path):
    with pytest.raises(J.JotsError, match="cannot open mesh"):
        J.Mesh.read(str(tmp_path / "missing.mesh"))
    with pytest.raises(J.JotsError, match="FE_Order"):
        J.Space(J.Mesh.brick(1, 1, 1, 1, 1, 1), 7)
    bad = tmp_path / "bad.mesh"
    bad.write_text("not a mesh\n")
    with pytest.raises(J.JotsError, match="MFEM mesh v1.0"):
        J.Mesh.read(str(bad))


@pytest.mark.parametrize("p,E", [(1, 32), (2, 16), (2, 14), (3, 12)])
def _config_pairs(path):
    """(library Config, oracle Config) of one .ini as comparable dicts."""
    import jots_b200 as J
    from oracle import jots as OJ
    d = J.config_dump(path)
    o = OJ.Config(path)
    lib = {k: d[k] for k in ("sim_type", "mesh_file", "time_scheme", "solver", "prec")}
    ora = {k: getattr(o, k) for k in lib}
    for k in ("use_restart", "with_precice", "using_time_integration", "using_newton"):
        lib[k], ora[k] = bool(int(d[k])), bool(getattr(o, k))
    for k in ("fe_order", "serial_refine", "parallel_refine", "max_timesteps", "max_iter", "newton_max_iter"):
        lib[k], ora[k] = int(d[k]), int(getattr(o, k))
    for k in ("initial_temp", "dt", "abs_tol", "rel_tol", "newton_abs_tol", "newton_rel_tol"):
        lib[k], ora[k] = float(d[k]), float(getattr(o, k))
    lib["mat"] = {k[4:]: v for k, v in d.items() if k.startswith("mat:")}
    ora["mat"] = {k: list(v) for k, v in o.mat_props.items()}
    lib["bcs"] = {int(k[3:]): v for k, v in d.items() if k.startswith("bc:")}
    ora["bcs"] = {int(k): list(v) for k, v in o.bcs.items()}
    return lib, ora


def test_config_defaults_and_quirks_match_the_oracle(tmp_path):
    """Config (config_file.cpp:7-172): missing sections take the defaults, a commented-out [NewtonSolverSettings] means
    zero Newton iterations (Heated_Plate as shipped), non-numeric values fall back to the default."""
    from cases import ini_text
    variants = {
        "full": ini_text("Nonlinear_Unsteady", "m.mesh", 300, [("Density_rho", "Uniform, 8000"), ("Specific_Heat_C", "Polynomial, 4.5, -850"),
                         ("Thermal_Conductivity_k", "Polynomial, 1e-4, 0.09, -17")], ["HeatFlux, 7.5e5", "Isothermal, 300", "Sinusoidal_Isothermal, 1, 2, 3, 4"],
                         newton=50, solver="CG", prec="Jacobi", steps=7, dt="2.5e-3", order=3, scheme="RK4", refine=2),
        "steady_no_newton": ini_text("Steady", "m.mesh", 0.5, [("Thermal_Conductivity_k", "Uniform, 1")], ["Isothermal, 0", "Isothermal, 1"], time=False),
        "minimal": "[FiniteElementSetup]\nMesh_File=x.mesh\n\n[MaterialProperties]\nThermal_Conductivity_k=Uniform, 2\n\n[BoundaryConditions]\nBoundary_Attr_1=HeatFlux, 0\n",
        "bad_numbers": "[FiniteElementSetup]\nFE_Order=two\nInitial_Temperature=hot\nMesh_File=x.mesh\n\n[MaterialProperties]\nDensity_rho=Uniform, 1\n\n"
                       "[BoundaryConditions]\nBoundary_Attr_2=Isothermal, 5\n\n[TimeIntegration]\nDelta_Time=fast\nMax_Timesteps=3.7\n\n[LinearSolverSettings]\nMax_Iterations=lots\n",
    }
    for name, txt in variants.items():
        p = str(tmp_path / (name + ".ini"))
        with open(p, "w") as f:
            f.write(txt)
        lib, ora = _config_pairs(p)
        assert lib == ora, (name, {k: (lib[k], ora[k]) for k in lib if lib[k] != ora[k]})


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present (GPU box)")
def test_every_shipped_config_parses_like_the_oracle():
    """All .ini files the reference ships (examples/ and tests/*.ini.in): same keys, values and defaults on both sides."""
    import glob
    paths = sorted(glob.glob(os.path.join(REF, "examples", "**", "*.ini"), recursive=True))
    assert len(paths) >= 10
    for p in paths:
        lib, ora = _config_pairs(p)
        assert lib == ora, (p, {k: (lib[k], ora[k]) for k in lib if lib[k] != ora[k]})


@pytest.mark.parametrize("source", ["brick", "ragged_brick", "file_mesh", "subdomain"])
def test_patches_reproduce_the_gather_table_bit_exact(tmp_path, source):
    """Element patches of the unique-pencil kernel (apply_patch.cuh, patches.cpp): every element sits in exactly one patch
    slot, its 27 nodes are found at [k][2 ly + j][2 lx + i] with the essential complement, slots that no element touches
    are marked unused, and the kernel's phase-3 items visit every (element slot, local ij) exactly once.  The lattice is
    recovered from the gather table alone: lexicographic bricks, first-appearance numbering of a mesh file and the
    owned-first numbering of a partitioned subdomain all give full patches."""
    import jots_b200 as J
    from cases import write_brick_mesh
    p, D = 2, 3
    px, py, items = J.patch_shape(p)
    assert (px, py) == (4, 4) and len(items) == 90
    if source == "brick":
        s = J.Space(J.Mesh.brick(8, 12, 3, 0.01, 0.02, 0.03), p)
    elif source == "ragged_brick":
        s = J.Space(J.Mesh.brick(6, 5, 2, 0.01, 0.02, 0.03), p)
    elif source == "file_mesh":
        path = str(tmp_path / "m.mesh")
        write_brick_mesh(path, 8, 4, 3, 0.01, 0.02, 0.015, grade=0.4)
        s = J.Space(J.Mesh.read(path), p)
    else:
        mesh = J.Mesh.brick(16, 8, 4, 0.01, 0.02, 0.03)
        part = J.Partition(mesh, J.Space(mesh, p), 2, 1)
        s = part.space()
    g = s.gather()
    ne = g.shape[0]
    ess = s.essential_dofs([1, 5]) if source != "subdomain" else np.unique(g[: ne // 3, :9].ravel()).astype(np.int32)
    em = np.zeros(s.n_dof, bool)
    em[ess] = True
    tab, fill = s.patches(ess, px, py)
    nxn, nyn = 2 * px + 1, 2 * py + 1
    npen, nn = nxn * nyn, nxn * nyn * D
    assert tab.shape[1] == (nn + px * py + 3) // 4 * 4
    assert abs(fill - ne / (tab.shape[0] * px * py)) < 1e-15
    if source != "ragged_brick":
        assert fill == 1.0
    INVALID = -2 ** 31
    seen = np.zeros(ne, int)
    for rec in tab:
        nodes = rec[:nn].reshape(D, nyn, nxn)
        elems = rec[nn:nn + px * py]
        touched = np.zeros((D, nyn, nxn), bool)
        for slot, e in enumerate(elems):
            if e < 0:
                continue
            seen[e] += 1
            ly, lx = divmod(slot, px)
            ge = g[e].reshape(D, D, D)                       # [k][j][i]
            want = np.where(em[ge], ~ge, ge)
            assert (nodes[:, 2 * ly:2 * ly + 3, 2 * lx:2 * lx + 3] == want).all()
            touched[:, 2 * ly:2 * ly + 3, 2 * lx:2 * lx + 3] = True
        assert (nodes[~touched] == INVALID).all() and (nodes[touched] != INVALID).all()
    assert (seen == 1).all()
    # phase-3 items: every (element slot, local ij) of a full patch exactly once, on the right pencil
    visits = np.zeros((px * py, D * D), int)
    for it in items:
        pencil, nv = int(it) & 0xff, (int(it) >> 8) & 3
        J_, I_ = divmod(pencil, nxn)
        assert 1 <= nv <= 2
        for v in range(nv):
            ev = (int(it) >> (10 + 10 * v)) & 0x3ff
            el, ij = ev >> 5, ev & 31
            ly, lx = divmod(el, px)
            j, i = divmod(ij, D)
            assert (2 * ly + j, 2 * lx + i) == (J_, I_)
            visits[el, ij] += 1
    assert (visits == 1).all()


def test_patches_absent_when_the_mesh_is_not_a_lattice(tmp_path):
    """Q1 has no face-interior node to recover the lattice from (and no patch instantiation); a 99 x 1 x 1 row keeps the
    gather-table kernel through the fill threshold of the context (fill 0.25 here)."""
    import jots_b200 as J
    s1 = J.Space(J.Mesh.brick(4, 4, 4, 1.0, 1.0, 1.0), 1)
    assert s1.patches(np.zeros(0, np.int32))[0].shape[0] == 0
    s = J.Space(J.Mesh.brick(99, 1, 1, 0.01, 0.01, 0.01), 2)
    tab, fill = s.patches(np.zeros(0, np.int32))
    assert tab.shape[0] == 25 and abs(fill - 99 / 400) < 1e-12


def _caller_tables(part, seed):
    """Emulate what ParMesh + ParFiniteElementSpace hold on one rank: the subdomain of `part` with its L-vector DOFs in a
    scrambled (caller) numbering, true DOFs in ascending global id, shared lists per neighbour in scrambled order."""
    rng = np.random.default_rng(seed)
    s = part.space()
    gid, _ = part.global_ids()
    n = part.n_local
    sigma = rng.permutation(n)                 # caller index j <-> library local index sigma[j]
    inv = np.empty(n, dtype=np.int64)
    inv[sigma] = np.arange(n)
    g = inv[s.gather()]
    bg, ba, _, _ = s.boundary()
    ec, bc = s.corners()
    nbrs = [(r, rng.permutation(inv[idx])) for r, idx in part.neighbors()]
    rng.shuffle(nbrs)
    return dict(dim=s.dim, order=s.order, gather=g, elem_corners=ec, bdr_gather=inv[bg], bdr_corners=bc, bdr_attr=ba,
                ldof_global_id=gid[sigma], ldof_of_true=inv[:part.n_owned], neighbors=nbrs)


@pytest.mark.parametrize("shape,p,nranks", [((6, 4, 4), 2, 4), ((5, 3, 2), 3, 2), ((4, 4, 4), 1, 8)])
def test_partition_from_caller_tables_matches_the_library_partition_bit_exact(shape, p, nranks):
    """jots_partition_create_from_tables (the caller's ParMesh / ParFiniteElementSpace decomposition with an arbitrary L-vector
    numbering) must give, index for index, the subdomain jots_partition_create builds: local numbering (owned first in true-DOF
    order, ghosts by global id), gather and boundary tables, neighbours with their shared lists, canonical reduction lists."""
    import jots_b200 as J
    mesh = J.Mesh.brick(*shape, 0.01, 0.02, 0.03)
    space = J.Space(mesh, p)
    for rank in range(nranks):
        ref = J.Partition(mesh, space, nranks, rank)
        t = _caller_tables(ref, 100 + rank)
        new = J.Partition.from_tables(rank=rank, nranks=nranks, **t)
        assert (new.n_local, new.n_owned, new.n_elem, new.n_neighbors, new.n_shared) == (ref.n_local, ref.n_owned, ref.n_elem, ref.n_neighbors, ref.n_shared)
        assert (new.global_ids()[0] == ref.global_ids()[0]).all()
        sn, sr = new.space(), ref.space()
        assert (sn.gather() == sr.gather()).all()
        assert (sn.boundary()[0] == sr.boundary()[0]).all() and (sn.boundary()[1] == sr.boundary()[1]).all()
        assert np.array_equal(sn.corners()[0], sr.corners()[0]) and np.array_equal(sn.corners()[1], sr.corners()[1])
        for (ra, ia), (rb, ib) in zip(new.neighbors(), ref.neighbors()):
            assert ra == rb and (ia == ib).all()
        for a, b in zip(new.reduction(), ref.reduction()):
            assert (a == b).all()


def test_partition_from_caller_tables_rejects_bad_input():
    import jots_b200 as J
    mesh = J.Mesh.brick(4, 2, 2, 1.0, 1.0, 1.0)
    ref = J.Partition(mesh, J.Space(mesh, 2), 2, 0)
    t = _caller_tables(ref, 3)
    bad = dict(t, ldof_of_true=np.concatenate([t["ldof_of_true"][:-1], t["ldof_of_true"][:1]]))   # a true DOF listed twice
    with pytest.raises(J.JotsError, match="injective"):
        J.Partition.from_tables(rank=0, nranks=2, **bad)
    bad = dict(t, neighbors=[(0, t["neighbors"][0][1])])                                            # itself as a neighbour
    with pytest.raises(J.JotsError, match="neighbour"):
        J.Partition.from_tables(rank=0, nranks=2, **bad)


def test_print_level_masks():
    """Factory::CreatePrintLevel (jots_common.inl:57-91): flags accumulate, None clears, All sets all, unknown labels fail."""
    import jots_b200 as J
    m = J.load_library().jots_print_level_mask
    assert m(b"Errors, Warnings") == 3 and m(b"Errors, Warnings, Iterations") == 7
    assert m(b"Summary") == 8 and m(b"FirstAndLast") == 16 and m(b"All") == 31
    assert m(b"All, None") == 0 and m(b"None, Warnings") == 2 and m(b"") == 0
    assert m(b"Verbose") == -1


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("p", [1, 2, 3, 4])
def test_h1_native_dof_map(dim, p):
    """MFEM-native local order: a bijection with the vertices first, then edge interiors (p-1 per edge,
    collinear, hex edges in increasing coordinate), face interiors ((p-1)^2 per face, coplanar on a boundary face), interior;
    and jots_gather_from_native brings a natively ordered element table back to the lexicographic one bit-exactly."""
    import jots_b200 as J
    D = p + 1
    m = J.h1_dof_map(dim, p)
    assert sorted(m) == list(range(D ** dim))
    coords = np.array([[l % D, (l // D) % D, l // (D * D)][:dim] for l in range(D ** dim)])
    nat2coord = np.empty_like(coords)
    nat2coord[m] = coords
    nv, ne = 2 ** dim, (4 if dim == 2 else 12) * (p - 1)
    assert all(set(c) <= {0, p} for c in nat2coord[:nv])                        # vertices first
    expected_v = [(0, 0), (p, 0), (p, p), (0, p)] if dim == 2 else [(0, 0, 0), (p, 0, 0), (p, p, 0), (0, p, 0), (0, 0, p), (p, 0, p), (p, p, p), (0, p, p)]
    assert [tuple(c) for c in nat2coord[:nv]] == expected_v
    edges = nat2coord[nv:nv + ne].reshape(-1, max(p - 1, 1), dim) if p > 1 else np.zeros((0, 1, dim), int)
    for e in edges:                                                             # one free coordinate, the others on the boundary
        free = [a for a in range(dim) if len(set(e[:, a])) > 1 or e[0, a] not in (0, p)]
        assert len(free) == 1 and all(e[0, a] in (0, p) and len(set(e[:, a])) == 1 for a in range(dim) if a != free[0])
        step = np.diff(e[:, free[0]])
        assert (step == 1).all() or (dim == 2 and (step == -1).all())
    if dim == 3 and p > 1:
        nf = 6 * (p - 1) ** 2
        faces = nat2coord[nv + ne:nv + ne + nf].reshape(6, (p - 1) ** 2, 3)
        for f in faces:
            fixed = [a for a in range(3) if len(set(f[:, a])) == 1 and f[0, a] in (0, p)]
            assert len(fixed) == 1
        assert (np.all((nat2coord[nv + ne + nf:] > 0) & (nat2coord[nv + ne + nf:] < p)))
    s = J.Space(J.Mesh.brick(3, 2, 2, 1.0, 1.0, 1.0) if dim == 3 else J.Mesh.rect(3, 2, 1.0, 1.0), p)
    g = s.gather()
    native = np.empty_like(g)
    native[:, m] = g                                                           # native[e][map[l]] = lexicographic[e][l]
    assert (J.gather_from_native(dim, p, native) == g).all()
    assert (J.gather_from_native(dim, p, -1 - native) == g).all()              # MFEM's sign encoding is undone


@pytest.mark.parametrize("dim,p", [(3, 1), (3, 2), (3, 3), (2, 2), (2, 4)])
def test_restart_files_round_trip_and_mfem_numbering(tmp_path, dim, p):
    """Restart data collection (output_manager.cpp:15-21,89-95; jots_driver.cpp:214-257): root JSON + mesh + H1 grid function
    in MFEM's DOF numbering.  The numbering is a bijection with the vertex DOFs first (= vertex ids), then p-1 DOFs per edge
    running from the lower to the higher vertex id, then face and element interiors; write -> read gives back mesh, order,
    cycle, time and the field to the 16 printed digits (the reference's restart test allows 1e-12)."""
    import jots_b200 as J
    mesh = J.Mesh.brick(3, 2, 2, 0.3, 0.2, 0.1) if dim == 3 else J.Mesh.rect(4, 3, 1.0, 0.5)
    s = J.Space(mesh, p)
    num = J.mfem_dof_numbering(mesh, s)
    assert sorted(num) == list(range(s.n_dof))
    verts, elems, _, _ = mesh.arrays()
    nv = verts.shape[0]
    X = s.node_coordinates()
    # vertex DOFs: MFEM number = vertex id, at the vertex coordinates
    for n in np.nonzero(num < nv)[0]:
        assert np.allclose(X[n], verts[num[n]], atol=1e-14)
    if p > 1:
        # the p-1 DOFs of the first edge (element 0, local vertices 0 -> 1) lie on it, ordered from the lower vertex id
        v0, v1 = elems[0][0], elems[0][1]
        lo, hi = (v0, v1) if v0 < v1 else (v1, v0)
        pos = {int(num[n]): X[n] for n in range(s.n_dof) if nv <= num[n] < nv + (p - 1)}
        ts = [np.dot(pos[nv + j] - verts[lo], verts[hi] - verts[lo]) / np.dot(verts[hi] - verts[lo], verts[hi] - verts[lo]) for j in range(p - 1)]
        assert all(0 < t < 1 for t in ts) and ts == sorted(ts)
    rng = np.random.default_rng(3)
    u = 300.0 + 50.0 * rng.standard_normal(s.n_dof)
    prefix = str(tmp_path / "restart")
    J.restart_write(prefix, 1000, 0.7251, mesh, s, u)
    import json, os
    root = json.load(open(prefix + "_001000.mfem_root"))["dsets"]["main"]
    assert root["cycle"] == 1000 and root["domains"] == 1 and root["fields"]["Temperature"]["path"] == "restart_001000/Temperature.%06d"
    assert open(prefix + "_001000/Temperature.000000").read().startswith("FiniteElementSpace\nFiniteElementCollection: H1_%dD_P%d\nVDim: 1\nOrdering: 0\n" % (dim, p))
    m2, order, cyc, t, vals = J.restart_read(prefix, 1000)
    assert (order, cyc) == (p, 1000) and abs(t - 0.7251) < 1e-15
    v2, e2, b2, a2 = m2.arrays()
    assert np.allclose(v2, verts, rtol=1e-15, atol=0) and (e2 == elems).all()
    s2 = J.Space(m2, order)
    back = vals[J.mfem_dof_numbering(m2, s2)]
    # the re-read mesh is a FILE mesh (first-appearance node numbering): compare at the nodes' coordinates
    X2 = s2.node_coordinates()
    key = lambda A: np.lexsort(np.round(A / 1e-9).T[::-1])
    assert np.abs(back[key(X2)] - u[key(X)]).max() <= 1e-13 * np.abs(u).max()


@pytest.mark.parametrize("dim,p", [(2, 1), (2, 2), (2, 3), (2, 4), (3, 1), (3, 2), (3, 3), (3, 4)])
def test_paraview_collection_against_an_independent_reader(tmp_path, dim, p):
    """ParaView output (OutputManager::WriteVizOutput, output_manager.cpp:23-28,81-87: high-order, binary): the .vtu piece read
    back by oracle/vtkfmt.py -- one Lagrange cell of order p per element, its nodes at the positions VTK's node order
    prescribes (constructive restatement, independent of the index formula in paraview.cpp), point data equal to the
    polynomial the nodal field samples (degree p per variable: the interpolation from the Gauss-Lobatto nodes to VTK's uniform
    lattice is exact), cell attributes, and the collection file in the three modes (new / append / restart)."""
    import jots_b200 as J
    from oracle import vtkfmt
    mesh = J.Mesh.brick(3, 2, 2, 0.3, 0.2, 0.1) if dim == 3 else J.Mesh.rect(4, 3, 1.0, 0.5)
    s = J.Space(mesh, p)
    nodes = vtkfmt.lagrange_nodes(dim, p)
    D = p + 1
    assert len(nodes) == D ** dim and len(set(nodes)) == D ** dim
    lex = [c[0] + D * (c[1] + D * (c[2] if dim == 3 else 0)) for c in nodes]
    assert list(J.vtk_lagrange_order(dim, p)) == lex

    def poly(X):
        f = (1.0 + 2.0 * X[:, 0]) ** p * (0.5 - X[:, 1]) ** p
        return f * (2.0 + 3.0 * X[:, 2]) ** p if dim == 3 else f
    X = s.node_coordinates()
    u = 300.0 + 40.0 * poly(X)
    rank = np.full(s.n_dof, 3.0)
    d = str(tmp_path / "ParaView")
    J.paraview_write(d, 0, 0.0, mesh, s, {"Rank": rank, "Temperature": u})
    J.paraview_write(d, 10, 0.1, mesh, s, {"Rank": rank, "Temperature": u}, mode=1)
    J.paraview_write(d, 20, 0.2, mesh, s, {"Rank": rank, "Temperature": u}, mode=1)
    assert vtkfmt.read_pvd(d + "/ParaView.pvd") == [(0.0, "Cycle000000/data.pvtu"), (0.1, "Cycle000010/data.pvtu"), (0.2, "Cycle000020/data.pvtu")]
    J.paraview_write(d, 15, 0.15, mesh, s, {"Rank": rank, "Temperature": u}, mode=2)      # restarted from t = 0.15: later entries dropped
    assert vtkfmt.read_pvd(d + "/ParaView.pvd") == [(0.0, "Cycle000000/data.pvtu"), (0.1, "Cycle000010/data.pvtu"), (0.15, "Cycle000015/data.pvtu")]
    J.paraview_write(d, 0, 0.0, mesh, s, {"Rank": rank, "Temperature": u})                   # a new run starts a new collection
    assert vtkfmt.read_pvd(d + "/ParaView.pvd") == [(0.0, "Cycle000000/data.pvtu")]
    hdr = vtkfmt.read_pvtu(d + "/Cycle000010/data.pvtu")
    assert hdr == {"pieces": ["proc000000.vtu"], "point_arrays": ["Rank", "Temperature"], "cell_arrays": ["attribute"]}
    v = vtkfmt.read_vtu(d + "/Cycle000010/proc000000.vtu")
    ne, nloc = s.n_elem, D ** dim
    assert v["n_cells"] == ne and v["n_points"] == ne * nloc and v["version"] == "2.2"
    assert (v["types"] == (72 if dim == 3 else 70)).all() and list(v["offsets"]) == [nloc * (e + 1) for e in range(ne)]
    assert sorted(v["connectivity"]) == list(range(ne * nloc))
    assert (v["cell_data"]["attribute"] == 1).all()          # the generated meshes carry element attribute 1
    g = s.gather()
    P = v["points"]
    assert P.shape == (ne * nloc, 3) and (dim == 3 or (P[:, 2] == 0).all())
    for e in range(ne):
        corner = X[g[e][0]]                                 # lexicographic node 0 = lattice origin of the element
        far = X[g[e][nloc - 1]]
        for j, c in enumerate(nodes):
            pt = P[v["connectivity"][e * nloc + j]]
            want = [corner[a] + (far[a] - corner[a]) * c[a] / p for a in range(dim)]
            assert np.allclose(pt[:dim], want, atol=1e-14), (e, j, pt, want)
    assert np.abs(v["point_data"]["Temperature"] - (300.0 + 40.0 * poly(P))).max() < 1e-10
    assert np.abs(v["point_data"]["Rank"] - 3.0).max() < 1e-14
