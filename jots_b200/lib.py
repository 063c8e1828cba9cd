"""ctypes binding of libjots_b200.so (declarations: include/jots_b200.h)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HOST, DEVICE = 0, 1
UNIFORM, POLYNOMIAL = 0, 1
BC_ISOTHERMAL, BC_HEATFLUX, BC_SIN_ISOTHERMAL, BC_SIN_HEATFLUX = 0, 1, 2, 3
CG, GMRES, FGMRES = 0, 1, 2
JACOBI, CHEBYSHEV = 0, 1
EULER_IMPLICIT, EULER_EXPLICIT, RK4 = 0, 1, 2
LINEARIZED_UNSTEADY, NONLINEAR_UNSTEADY, STEADY = 0, 1, 2
DENSITY, SPECIFIC_HEAT, THERMAL_CONDUCTIVITY = 0, 1, 2
FORM_MASS, FORM_STIFFNESS, FORM_SYSTEM, FORM_RESIDUAL = 0, 1, 2, 3


class JotsError(RuntimeError):
    pass


_LIB = None
_NOHANDLE = object()
_P = C.c_void_p
_D = C.POINTER(C.c_double)
_I32 = C.POINTER(C.c_int32)
_I64 = C.POINTER(C.c_int64)


def library_path():
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "libjots_b200.so")


def load_library():
    """Load the CUDA library; fails loudly when it has not been built (no fallback)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise JotsError("libjots_b200.so is not built (run `python -c 'import __graft_entry__ as g; g.build()'` "
                        "or `make -C jots_b200/csrc`); jots_b200 has no CPU fallback")
    lib = C.CDLL(path)
    sig = {
        "jots_last_error": (C.c_char_p, []),
        "jots_version": (C.c_char_p, []),
        "jots_device_count": (C.c_int, []),
        "jots_mesh_read": (_P, [C.c_char_p]),
        "jots_mesh_brick": (_P, [C.c_int64] * 3 + [C.c_double] * 3),
        "jots_mesh_rect": (_P, [C.c_int64] * 2 + [C.c_double] * 4),
        "jots_mesh_refine": (_P, [_P]),
        "jots_mesh_destroy": (None, [_P]),
        "jots_mesh_info": (C.c_int, [_P, C.POINTER(C.c_int), _I64, _I64, _I64]),
        "jots_mesh_copy": (C.c_int, [_P, _P, _P, _P, _P]),
        "jots_mesh_set_vertices": (C.c_int, [_P, _P]),
        "jots_space_create": (_P, [_P, C.c_int]),
        "jots_space_destroy": (None, [_P]),
        "jots_space_boundary_adjacency": (C.c_int, [_P, _P, _P, _P]),
        "jots_space_create_from_tables": (_P, [C.c_int, C.c_int, C.c_int64, C.c_int64, _P, _P, C.c_int64, _P, _P, _P]),
        "jots_space_copy_corners": (C.c_int, [_P, _P, _P]),
        "jots_space_info": (C.c_int, [_P, C.POINTER(C.c_int), C.POINTER(C.c_int), _I64, _I64, _I64, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
        "jots_space_copy_gather": (C.c_int, [_P, _P]),
        "jots_space_copy_bdr": (C.c_int, [_P, _P, _P, _P, _P]),
        "jots_space_copy_nodes": (C.c_int, [_P, _P]),
        "jots_space_essential_dofs": (C.c_int64, [_P, _P, C.c_int, _P]),
        "jots_space_patches": (C.c_int64, [_P, _P, C.c_int64, C.c_int, C.c_int, _P, _P, _P]),
        "jots_patch_shape": (C.c_int, [C.c_int, _P, _P, _P]),
        "jots_h1_dof_map": (C.c_int, [C.c_int, C.c_int, _P]),
        "jots_gather_from_native": (C.c_int, [C.c_int, C.c_int, C.c_int64, _P, _P]),
        "jots_mfem_dof_numbering": (C.c_int, [_P, _P, _P]),
        "jots_restart_write": (C.c_int, [C.c_char_p, C.c_int, C.c_double, _P, _P, _P]),
        "jots_restart_read": (_P, [C.c_char_p, C.c_int, _P, _P, _P, _P, _P]),
        "jots_paraview_write": (C.c_int, [C.c_char_p, C.c_int, C.c_double, _P, _P, C.c_int, _P, _P, C.c_int]),
        "jots_vtk_lagrange_order": (C.c_int, [C.c_int, C.c_int, _P]),
        "jots_config_dump": (C.c_int64, [C.c_char_p, _P, C.c_int64]),
        "jots_ctx_create": (_P, [_P, C.c_int]),
        "jots_ctx_destroy": (None, [_P]),
        "jots_ctx_size": (C.c_int64, [_P]),
        "jots_ctx_set_stream": (C.c_int, [_P, _P]),
        "jots_ctx_sync": (C.c_int, [_P]),
        "jots_ctx_set_material": (C.c_int, [_P, C.c_int, C.c_int, _P, C.c_int]),
        "jots_ctx_set_bcs": (C.c_int, [_P, C.c_int, _P, _P]),
        "jots_ctx_essential_dofs": (C.c_int64, [_P, _P]),
        "jots_ctx_update_bcs": (C.c_int, [_P, C.c_double]),
        "jots_ctx_apply_dirichlet": (C.c_int, [_P, _P, C.c_int, C.c_int]),
        "jots_ctx_set_material_state": (C.c_int, [_P, _P, C.c_int]),
        "jots_ctx_precice_bc_size": (C.c_int64, [_P, C.c_int]),
        "jots_ctx_precice_bc_nodes": (C.c_int, [_P, C.c_int, _P, _P]),
        "jots_ctx_precice_bc_set_read_data": (C.c_int, [_P, C.c_int, _P]),
        "jots_ctx_precice_bc_write_data": (C.c_int, [_P, C.c_int, _P, C.c_int, _P]),
        "jots_ctx_material_field": (C.c_int, [_P, C.c_int, C.c_int, _P, C.c_int]),
        "jots_ctx_stats": (C.c_int, [_P, _P, _P]),
        "jots_ctx_expand_true": (C.c_int, [_P, _P, _P, C.c_int]),
        "jots_partition_create_from_tables": (_P, [C.c_int, C.c_int, C.c_int64, C.c_int64, _P, _P, C.c_int64, _P, _P, _P, _P, C.c_int64, _P, C.c_int, _P, _P, _P, C.c_int, _P, C.c_int, C.c_int]),
        "jots_ctx_patch_info": (C.c_int, [_P, _P, _P, _P]),
        "jots_ctx_fused_dots": (C.c_int, [_P, _P]),
        "jots_ctx_side_applies": (C.c_int, [_P, _P]),
        "jots_ctx_last_kernel": (C.c_char_p, [_P]),
        "jots_ctx_reset_stats": (C.c_int, [_P]),
        "jots_op_create": (_P, [_P, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int,
                                C.c_double, C.c_double, C.c_int]),
        "jots_op_destroy": (None, [_P]),
        "jots_op_set_print_level": (C.c_int, [_P, C.c_int, C.c_int]),
        "jots_print_level_mask": (C.c_int, [C.c_char_p]),
        "jots_op_iterate": (C.c_int, [_P, _P, C.c_int, _D]),
        "jots_op_implicit_solve": (C.c_int, [_P, C.c_double, _P, _P, C.c_int]),
        "jots_op_mult": (C.c_int, [_P, _P, _P, C.c_int]),
        "jots_op_process_mat_prop_update": (C.c_int, [_P, C.c_int]),
        "jots_op_update_neumann": (C.c_int, [_P]),
        "jots_op_get_neumann": (C.c_int, [_P, _P, C.c_int]),
        "jots_op_form_apply": (C.c_int, [_P, C.c_int, _P, _P, _P, C.c_int]),
        "jots_op_form_diagonal": (C.c_int, [_P, C.c_int, _P, _P, C.c_int]),
        "jots_op_form_apply_dot": (C.c_int, [_P, C.c_int, _P, _P, _P, C.c_int, _P, _P]),
        "jots_op_form_apply_side": (C.c_int, [_P, C.c_int, _P, _P, _P, C.c_int, _P, C.c_double, _P, _P, _P, _P]),
        "jots_op_form_time": (C.c_int, [_P, C.c_int, _P, _P, _P, C.c_int, _D]),
        "jots_partition_create": (_P, [_P, _P, C.c_int, C.c_int, _P]),
        "jots_partition_destroy": (None, [_P]),
        "jots_partition_info": (C.c_int, [_P, _I64, _I64, _I64, _I64, C.POINTER(C.c_int), _I64]),
        "jots_partition_space": (_P, [_P]),
        "jots_partition_copy": (C.c_int, [_P, _P, _P]),
        "jots_partition_neighbor": (C.c_int64, [_P, C.c_int, C.POINTER(C.c_int), _P]),
        "jots_partition_reduction": (C.c_int64, [_P, _P, _P, _P]),
        "jots_comm_unique_id": (C.c_int, [C.c_char_p]),
        "jots_comm_create": (_P, [C.c_char_p, C.c_int, C.c_int, C.c_int]),
        "jots_comm_destroy": (None, [_P]),
        "jots_ctx_create_part": (_P, [_P, C.c_int, _P]),
        "jots_driver_create_dist": (_P, [C.c_char_p, _P, C.c_int, _P]),
        "jots_driver_copy_global_ids": (C.c_int, [_P, _P, _I64]),
        "jots_driver_create": (_P, [C.c_char_p, C.c_int]),
        "jots_driver_create_on_mesh": (_P, [C.c_char_p, _P, C.c_int]),
        "jots_driver_destroy": (None, [_P]),
        "jots_driver_run": (C.c_int, [_P, C.c_int, C.POINTER(C.c_int)]),
        "jots_driver_size": (C.c_int64, [_P]),
        "jots_driver_get_u": (C.c_int, [_P, _P, C.c_int]),
        "jots_driver_set_u": (C.c_int, [_P, _P, C.c_int]),
        "jots_driver_time": (C.c_double, [_P]),
        "jots_driver_ctx": (_P, [_P]),
        "jots_driver_op": (_P, [_P]),
        "jots_driver_step_count": (C.c_int, [_P]),
        "jots_driver_space": (_P, [_P]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    lib._jots_symbols = sorted(sig)
    _LIB = lib
    return lib


def _err():
    return load_library().jots_last_error().decode()


def _chk(rc):
    if rc != 0:
        raise JotsError(_err())


def _handle(h):
    if not h:
        raise JotsError(_err())
    return h


def _ptr(a):
    """Pointer for a numpy array (host) / int (device address) / torch tensor (device) / None."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(C.c_void_p)
    if isinstance(a, int):
        return C.c_void_p(a)
    if hasattr(a, "data_ptr"):
        return C.c_void_p(a.data_ptr())
    raise TypeError(type(a))


def _loc(*arrs):
    loc = None
    for a in arrs:
        if a is None:
            continue
        l = HOST if isinstance(a, np.ndarray) else DEVICE
        if loc is not None and l != loc:
            raise JotsError("all vector arguments of one call must live in the same memory space")
        loc = l
    return HOST if loc is None else loc


def _f64(a):
    if isinstance(a, np.ndarray):
        if a.dtype != np.float64 or not a.flags.c_contiguous:
            raise JotsError("host vectors must be contiguous float64 arrays")
    return a


def patch_shape(order):
    """(px, py, items) of the unique-pencil patch kernel for an FE order; px == 0: none."""
    px, py = C.c_int(0), C.c_int(0)
    items = np.zeros(128, dtype=np.uint32)
    n = load_library().jots_patch_shape(order, C.byref(px), C.byref(py), _ptr(items))
    return px.value, py.value, items[:n]


def h1_dof_map(dim, order):
    """map[lexicographic l] = MFEM-native local DOF index (GetDofMap of the H1 quad / hex element)."""
    m = np.empty((order + 1) ** dim, dtype=np.int32)
    _chk(load_library().jots_h1_dof_map(dim, order, _ptr(m)))
    return m


def gather_from_native(dim, order, native):
    """Element-DOF table in MFEM's native local order -> lexicographic local order."""
    nat = np.ascontiguousarray(native, dtype=np.int32)
    out = np.empty_like(nat)
    _chk(load_library().jots_gather_from_native(dim, order, nat.shape[0], _ptr(nat), _ptr(out)))
    return out


def mfem_dof_numbering(mesh, space):
    """out[node of `space`] = MFEM's DOF number of that node in the H1 space on `mesh` (restart files)."""
    out = np.empty(space.n_dof, dtype=np.int64)
    _chk(load_library().jots_mfem_dof_numbering(mesh._h, space._h, _ptr(out)))
    return out


def restart_write(prefix, cycle, time, mesh, space, u_nodes):
    u = np.ascontiguousarray(u_nodes, dtype=np.float64)
    _chk(load_library().jots_restart_write(str(prefix).encode(), int(cycle), float(time), mesh._h, space._h, _ptr(u)))


def paraview_write(directory, cycle, time, mesh, space, fields, mode=0):
    """One cycle of a ParaView collection (<directory>/ParaView.pvd, Cycle%06d/data.pvtu, proc000000.vtu); fields: {name: nodal
    values in the numbering of `space`}; mode 0 new collection, 1 append, 2 restart mode."""
    names = [str(k).encode() for k in fields]
    arrs = [np.ascontiguousarray(v, dtype=np.float64) for v in fields.values()]
    cn = (C.c_char_p * len(names))(*names)
    cv = (C.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])
    _chk(load_library().jots_paraview_write(str(directory).encode(), int(cycle), float(time), mesh._h if mesh is not None else None,
                                            space._h, len(names), cn, cv, int(mode)))


def vtk_lagrange_order(dim, p):
    out = np.empty((p + 1) ** dim, dtype=np.int32)
    _chk(load_library().jots_vtk_lagrange_order(int(dim), int(p), _ptr(out)))
    return out


def restart_read(prefix, cycle):
    """(mesh, order, cycle, time, values in MFEM numbering) of the data collection <prefix>_<cycle>."""
    lib = load_library()
    order, cyc, n = C.c_int(), C.c_int(), C.c_int64()
    t = C.c_double()
    h = _handle(lib.jots_restart_read(str(prefix).encode(), int(cycle), C.byref(order), C.byref(cyc), C.byref(t), C.byref(n), None))
    lib.jots_mesh_destroy(h)
    vals = np.empty(n.value)
    h = _handle(lib.jots_restart_read(str(prefix).encode(), int(cycle), None, None, None, None, _ptr(vals)))
    return Mesh(h), order.value, cyc.value, t.value, vals


def config_dump(ini_path):
    """The host-side Config of an .ini file as a dict (scalars as strings, 'mat:<name>' / 'bc:<attr>' as lists)."""
    lib = load_library()
    n = lib.jots_config_dump(os.fsencode(ini_path), None, 0)
    if n < 0:
        raise JotsError(_err())
    buf = C.create_string_buffer(n)
    lib.jots_config_dump(os.fsencode(ini_path), C.cast(buf, C.c_void_p), n)
    out = {}
    for line in buf.value.decode().splitlines():
        k, v = line.split("=", 1)
        out[k] = v.split("|") if k.startswith(("mat:", "bc:")) else v
    return out


class Mesh:
    def __init__(self, handle):
        self._h = _handle(handle)

    @classmethod
    def read(cls, path):
        return cls(load_library().jots_mesh_read(os.fsencode(path)))

    @classmethod
    def brick(cls, nx, ny, nz, lx, ly, lz):
        return cls(load_library().jots_mesh_brick(nx, ny, nz, lx, ly, lz))

    @classmethod
    def rect(cls, nx, ny, lx, ly, x0=0.0, y0=0.0):
        return cls(load_library().jots_mesh_rect(nx, ny, lx, ly, x0, y0))

    def refine(self):
        return Mesh(load_library().jots_mesh_refine(self._h))

    def info(self):
        dim = C.c_int()
        ne, nb, nv = C.c_int64(), C.c_int64(), C.c_int64()
        _chk(load_library().jots_mesh_info(self._h, dim, ne, nb, nv))
        return dict(dim=dim.value, n_elem=ne.value, n_bdr=nb.value, n_verts=nv.value)

    def arrays(self):
        i = self.info()
        dim = i["dim"]
        verts = np.empty((i["n_verts"], dim))
        elems = np.empty((i["n_elem"], 2 ** dim), dtype=np.int64)
        bdr = np.empty((i["n_bdr"], 2 ** (dim - 1)), dtype=np.int64)
        battr = np.empty(i["n_bdr"], dtype=np.int32)
        _chk(load_library().jots_mesh_copy(self._h, _ptr(verts), _ptr(elems), _ptr(bdr), _ptr(battr)))
        return verts, elems, bdr, battr

    def set_vertices(self, verts):
        """New vertex coordinates [n_verts, dim] on the same topology (mfem::Mesh::SetVertices)."""
        _chk(load_library().jots_mesh_set_vertices(self._h, _ptr(np.ascontiguousarray(verts, dtype=np.float64))))

    def __del__(self):
        if getattr(self, "_h", None) and _LIB is not None:
            _LIB.jots_mesh_destroy(self._h)
            self._h = None


class Space:
    def __init__(self, mesh=None, order=1, handle=_NOHANDLE):
        self._h = _handle(load_library().jots_space_create(mesh._h, order) if handle is _NOHANDLE else handle)
        dim, p, dpe, dpf = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        ne, nd, nb = C.c_int64(), C.c_int64(), C.c_int64()
        _chk(load_library().jots_space_info(self._h, dim, p, ne, nd, nb, dpe, dpf))
        self.dim, self.order, self.n_elem, self.n_dof, self.n_bdr = dim.value, p.value, ne.value, nd.value, nb.value
        self.dofs_per_elem, self.dofs_per_face = dpe.value, dpf.value

    @classmethod
    def from_tables(cls, dim, order, n_dof, gather, elem_corners, bdr_gather, bdr_corners, bdr_attr):
        gather = np.ascontiguousarray(gather, dtype=np.int32)
        ec = np.ascontiguousarray(elem_corners, dtype=np.float64)
        bg = np.ascontiguousarray(bdr_gather, dtype=np.int32)
        bc = np.ascontiguousarray(bdr_corners, dtype=np.float64)
        ba = np.ascontiguousarray(bdr_attr, dtype=np.int32)
        h = load_library().jots_space_create_from_tables(dim, order, gather.shape[0], n_dof, _ptr(gather), _ptr(ec), ba.size,
                                                         _ptr(bg), _ptr(bc), _ptr(ba))
        return cls(handle=h)

    def corners(self):
        ec = np.empty((self.n_elem, 2 ** self.dim, self.dim))
        bc = np.empty((self.n_bdr, 2 ** (self.dim - 1), self.dim))
        _chk(load_library().jots_space_copy_corners(self._h, _ptr(ec), _ptr(bc)))
        return ec, bc

    def gather(self):
        g = np.empty((self.n_elem, self.dofs_per_elem), dtype=np.int32)
        _chk(load_library().jots_space_copy_gather(self._h, _ptr(g)))
        return g

    def boundary(self):
        bg = np.empty((self.n_bdr, self.dofs_per_face), dtype=np.int32)
        ba = np.empty(self.n_bdr, dtype=np.int32)
        be = np.empty(self.n_bdr, dtype=np.int64)
        bf = np.empty(self.n_bdr, dtype=np.int32)
        _chk(load_library().jots_space_copy_bdr(self._h, _ptr(bg), _ptr(ba), _ptr(be), _ptr(bf)))
        return bg, ba, be, bf

    def boundary_adjacency(self):
        """(adjacent element, its local face, place of every boundary node in that element's lattice) per boundary element;
        recovered from the DOF ids when the space came from tables (host only)."""
        be = np.empty(self.n_bdr, dtype=np.int32)
        bf = np.empty(self.n_bdr, dtype=np.int32)
        nl = np.empty((self.n_bdr, self.dofs_per_face), dtype=np.int32)
        _chk(load_library().jots_space_boundary_adjacency(self._h, _ptr(be), _ptr(bf), _ptr(nl)))
        return be, bf, nl

    def node_coordinates(self):
        x = np.empty((self.n_dof, self.dim))
        _chk(load_library().jots_space_copy_nodes(self._h, _ptr(x)))
        return x

    def essential_dofs(self, attrs):
        a = np.asarray(list(attrs), dtype=np.int32)
        n = load_library().jots_space_essential_dofs(self._h, _ptr(a), a.size, None)
        if n < 0:
            raise JotsError(_err())
        out = np.empty(n, dtype=np.int32)
        load_library().jots_space_essential_dofs(self._h, _ptr(a), a.size, _ptr(out))
        return out

    def patches(self, ess, px=4, py=4):
        """Element patches of the unique-pencil kernel: (records [n_patches, record_ints] int32, fill); record = nodes
        [k][J][I] (essential as ~index, 0x80000000 unused), then px*py element ids (-1: hole).  Empty when not a lattice."""
        ess = np.ascontiguousarray(ess, dtype=np.int32)
        rec, fill = C.c_int(0), C.c_double(0.0)
        lib = load_library()
        n = lib.jots_space_patches(self._h, _ptr(ess), ess.size, px, py, C.byref(rec), C.byref(fill), None)
        if n < 0:
            raise JotsError(_err())
        if n == 0:
            return np.empty((0, 0), dtype=np.int32), 0.0
        buf = np.empty((n, rec.value), dtype=np.int32)
        lib.jots_space_patches(self._h, _ptr(ess), ess.size, px, py, C.byref(rec), C.byref(fill), _ptr(buf))
        return buf, float(fill.value)

    def __del__(self):
        if getattr(self, "_h", None) and _LIB is not None:
            _LIB.jots_space_destroy(self._h)
            self._h = None


_BC_KIND = {"Isothermal": BC_ISOTHERMAL, "HeatFlux": BC_HEATFLUX, "Sinusoidal_Isothermal": BC_SIN_ISOTHERMAL,
            "Sinusoidal_HeatFlux": BC_SIN_HEATFLUX, "preCICE_Isothermal": 4, "preCICE_HeatFlux": 5}


class Context:
    """Device context of one subdomain."""

    def __init__(self, space=None, device=0, handle=None, owner=None):
        self._owner = owner  # keeps a Driver alive for borrowed handles
        self._borrowed = handle is not None
        self._h = _handle(handle if handle is not None else load_library().jots_ctx_create(space._h, device))
        self.n = load_library().jots_ctx_size(self._h)
        self.dim = space.dim if space is not None else None

    @classmethod
    def from_partition(cls, part, device, comm):
        """Context of one subdomain (jots_ctx_create_part); comm may be None (no exchange: single-rank testing)."""
        h = load_library().jots_ctx_create_part(part._h, device, comm._h if comm is not None else None)
        obj = cls(handle=_handle(h), owner=(part, comm))
        obj._borrowed = False
        obj.n_owned = part.n_owned
        return obj

    def expand_true(self, tvec, out=None):
        """True-DOF vector (n_owned entries, the caller's true-DOF order) -> consistent local vector (collective)."""
        out = np.empty(self.n) if out is None else out
        t = np.ascontiguousarray(tvec, dtype=np.float64)
        buf = np.zeros(self.n)
        buf[:t.size] = t
        _chk(load_library().jots_ctx_expand_true(self._h, _ptr(buf), _ptr(out), _loc(out)))
        return out

    def set_stream(self, stream):
        _chk(load_library().jots_ctx_set_stream(self._h, C.c_void_p(int(stream))))

    def sync(self):
        _chk(load_library().jots_ctx_sync(self._h))

    def set_material(self, which, spec):
        """spec: ('Uniform', v) / ('Polynomial', c0, c1, ...) -- the .ini value split on commas."""
        model = {"Uniform": UNIFORM, "Polynomial": POLYNOMIAL}.get(spec[0])
        if model is None:
            raise JotsError("Unknown/Invalid material model specified")
        c = np.asarray([float(v) for v in spec[1:]], dtype=np.float64)
        _chk(load_library().jots_ctx_set_material(self._h, which, model, _ptr(c), c.size))

    def set_bcs(self, specs):
        """specs[i] <-> boundary attribute i+1: ('Isothermal', v) / ('HeatFlux', v) / ('Sinusoidal_*', A, w, ph, s)."""
        kinds = np.zeros(len(specs), dtype=np.int32)
        params = np.zeros((len(specs), 4))
        for i, s in enumerate(specs):
            if s[0] not in _BC_KIND:
                raise JotsError("Invalid/Unknown boundary condition specified: '%s'" % s[0])
            kinds[i] = _BC_KIND[s[0]]
            vals = [float(v) for v in (s[2:] if s[0].startswith("preCICE") else s[1:])]    # preCICE: kind, coupling mesh, default
            params[i, :len(vals)] = vals
        _chk(load_library().jots_ctx_set_bcs(self._h, len(specs), _ptr(kinds), _ptr(params)))

    def essential_dofs(self):
        n = load_library().jots_ctx_essential_dofs(self._h, None)
        out = np.empty(n, dtype=np.int32)
        load_library().jots_ctx_essential_dofs(self._h, _ptr(out))
        return out

    def update_bcs(self, t):
        _chk(load_library().jots_ctx_update_bcs(self._h, float(t)))

    def apply_dirichlet(self, u, only_time_dependent=False):
        _chk(load_library().jots_ctx_apply_dirichlet(self._h, _ptr(_f64(u)), _loc(u), int(only_time_dependent)))

    def precice_bc_nodes(self, bc):
        """(dofs, coords) of the boundary-element nodes of preCICE BC `bc` (0-based BC index), duplicates kept."""
        n = load_library().jots_ctx_precice_bc_size(self._h, bc)
        if n < 0:
            raise JotsError("not a preCICE boundary condition")
        dim = self._dim()
        dofs = np.empty(n, dtype=np.int32)
        xyz = np.empty((n, dim))
        _chk(load_library().jots_ctx_precice_bc_nodes(self._h, bc, _ptr(xyz), _ptr(dofs)))
        return dofs, xyz

    def _dim(self):
        # coordinates per node: probe with the largest dimension, the library fills dim values per node
        if self.dim is None:
            raise JotsError("set Context.dim (2 or 3) before asking for boundary node coordinates")
        return int(self.dim)

    def precice_bc_set_read_data(self, bc, values):
        v = np.ascontiguousarray(values, dtype=np.float64)
        _chk(load_library().jots_ctx_precice_bc_set_read_data(self._h, bc, _ptr(v)))

    def precice_bc_write_data(self, bc, u):
        n = load_library().jots_ctx_precice_bc_size(self._h, bc)
        out = np.empty(max(n, 0))
        _chk(load_library().jots_ctx_precice_bc_write_data(self._h, bc, _ptr(_f64(u)), _loc(u), _ptr(out)))
        return out

    def set_material_state(self, u):
        _chk(load_library().jots_ctx_set_material_state(self._h, _ptr(_f64(u)), _loc(u)))

    def material_field(self, which, order=0, out=None):
        """Nodal field of property `which` (0 rho, 1 C, 2 k) or its first / second T-derivative at the material state."""
        out = np.empty(self.n) if out is None else out
        _chk(load_library().jots_ctx_material_field(self._h, which, order, _ptr(out), _loc(out)))
        return out

    def stats(self):
        out = np.zeros(12, dtype=np.int64)
        norms = np.zeros(2)
        _chk(load_library().jots_ctx_stats(self._h, _ptr(out), _ptr(norms)))
        keys = ["applies", "diag_assemblies", "krylov_iters", "newton_iters", "residual_evals", "steps", "kernel_launches",
                "dots", "last_linear_iters", "last_linear_converged", "last_newton_iters", "last_newton_converged"]
        d = {k: int(v) for k, v in zip(keys, out)}
        d["last_linear_norm"], d["last_newton_norm"] = float(norms[0]), float(norms[1])
        npch, fill, app = C.c_int64(0), C.c_double(0.0), C.c_int64(0)
        load_library().jots_ctx_patch_info(self._h, C.byref(npch), C.byref(fill), C.byref(app))
        d["n_patches"], d["patch_fill"], d["patch_applies"] = int(npch.value), float(fill.value), int(app.value)
        fd = C.c_int64(0)
        load_library().jots_ctx_fused_dots(self._h, C.byref(fd))
        d["fused_dots"] = int(fd.value)
        load_library().jots_ctx_side_applies(self._h, C.byref(fd))
        d["side_applies"] = int(fd.value)
        return d

    def last_kernel(self):
        """Label of the element kernel the last operator apply launched."""
        return (load_library().jots_ctx_last_kernel(self._h) or b"").decode()

    def reset_stats(self):
        _chk(load_library().jots_ctx_reset_stats(self._h))

    def __del__(self):
        if getattr(self, "_h", None) and not self._borrowed and _LIB is not None:
            _LIB.jots_ctx_destroy(self._h)
        self._h = None


class Operator:
    """Linear / Nonlinear / Steady conduction operator (JOTSIterator + TimeDependentOperator)."""

    def __init__(self, ctx, sim_type, dt=0.0, time_scheme=EULER_IMPLICIT, solver=FGMRES, prec=CHEBYSHEV,
                 rel_tol=1e-10, abs_tol=1e-16, max_iter=100, newton_rel_tol=1e-10, newton_abs_tol=1e-16, newton_max_iter=0,
                 handle=None, owner=None):
        self.ctx = ctx
        self._owner = owner
        self._borrowed = handle is not None
        if handle is not None:
            self._h = _handle(handle)
            return
        self._h = _handle(load_library().jots_op_create(ctx._h, sim_type, dt, time_scheme, solver, prec, rel_tol, abs_tol,
                                                        max_iter, newton_rel_tol, newton_abs_tol, newton_max_iter))

    def iterate(self, u, time=0.0):
        t = C.c_double(time)
        _chk(load_library().jots_op_iterate(self._h, _ptr(_f64(u)), _loc(u), C.byref(t)))
        return t.value

    def implicit_solve(self, dt, u, k):
        _chk(load_library().jots_op_implicit_solve(self._h, dt, _ptr(_f64(u)), _ptr(_f64(k)), _loc(u, k)))

    def mult(self, u, k):
        _chk(load_library().jots_op_mult(self._h, _ptr(_f64(u)), _ptr(_f64(k)), _loc(u, k)))

    def process_mat_prop_update(self, which):
        _chk(load_library().jots_op_process_mat_prop_update(self._h, which))

    def update_neumann(self):
        _chk(load_library().jots_op_update_neumann(self._h))

    def neumann_vector(self):
        b = np.empty(self.ctx.n)
        _chk(load_library().jots_op_get_neumann(self._h, _ptr(b), HOST))
        return b

    def form_apply(self, form, x, y, state=None):
        _chk(load_library().jots_op_form_apply(self._h, form, _ptr(_f64(x)), _ptr(state), _ptr(_f64(y)), _loc(x, state, y)))
        return y

    def form_apply_dot(self, form, x, y, state=None):
        """y = form(x); returns ((x, form(x)) over the owned rows, whether the element kernel formed it in the same pass)."""
        dot, fused = C.c_double(0.0), C.c_int(0)
        _chk(load_library().jots_op_form_apply_dot(self._h, form, _ptr(_f64(x)), _ptr(state), _ptr(_f64(y)), _loc(x, state, y),
                                                   C.byref(dot), C.byref(fused)))
        return dot.value, bool(fused.value)

    def form_apply_side(self, form, x, y, state, alpha, side_d, side_x, side_zero):
        """The conjugate-gradient loop's apply with its side stream (device tensors): y = form(x), side_x += alpha side_d,
        side_zero = 0 in one element-kernel launch; returns ((x, form(x)), whether the kernel carried it)."""
        dot, carried = C.c_double(0.0), C.c_int(0)
        _chk(load_library().jots_op_form_apply_side(self._h, form, _ptr(x), _ptr(state), _ptr(y), DEVICE, C.byref(dot), float(alpha),
                                                    _ptr(side_d), _ptr(side_x), _ptr(side_zero), C.byref(carried)))
        return dot.value, bool(carried.value)

    def form_diagonal(self, form, d, state=None):
        _chk(load_library().jots_op_form_diagonal(self._h, form, _ptr(state), _ptr(_f64(d)), _loc(state, d)))
        return d

    def form_time(self, form, x_dev, y_dev, state_dev=None, reps=10):
        ms = C.c_double()
        _chk(load_library().jots_op_form_time(self._h, form, _ptr(x_dev), _ptr(state_dev), _ptr(y_dev), reps, C.byref(ms)))
        return ms.value

    def __del__(self):
        if getattr(self, "_h", None) and not self._borrowed and _LIB is not None:
            _LIB.jots_op_destroy(self._h)
        self._h = None


class Partition:
    """One subdomain of a partitioned space (host-side index sets only)."""

    def __init__(self, mesh, space, nranks, rank, dims=None, handle=None):
        if handle is None:
            d = None if dims is None else np.asarray(dims, dtype=np.int32)
            handle = load_library().jots_partition_create(mesh._h, space._h, nranks, rank, _ptr(d))
        self._h = _handle(handle)
        nl, no, ng, ne, ns = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64()
        nn = C.c_int()
        _chk(load_library().jots_partition_info(self._h, nl, no, ng, ne, nn, ns))
        self.n_local, self.n_owned, self.n_global, self.n_elem = nl.value, no.value, ng.value, ne.value
        self.n_neighbors, self.n_shared = nn.value, ns.value
        self.rank, self.nranks = rank, nranks

    @classmethod
    def from_tables(cls, dim, order, gather, elem_corners, bdr_gather, bdr_corners, bdr_attr, ldof_global_id, ldof_of_true,
                    neighbors, rank, nranks, attributes=None):
        """The caller's decomposition (ParMesh + ParFiniteElementSpace tables of one rank): neighbors = [(rank, L-vector
        DOFs shared with it), ...]."""
        g = np.ascontiguousarray(gather, dtype=np.int32)
        ec = np.ascontiguousarray(elem_corners, dtype=np.float64)
        bg = np.ascontiguousarray(bdr_gather, dtype=np.int32)
        bc = np.ascontiguousarray(bdr_corners, dtype=np.float64)
        ba = np.ascontiguousarray(bdr_attr, dtype=np.int32)
        gid = np.ascontiguousarray(ldof_global_id, dtype=np.int64)
        lt = np.ascontiguousarray(ldof_of_true, dtype=np.int32)
        nr = np.ascontiguousarray([r for r, _ in neighbors], dtype=np.int32)
        ptr = np.zeros(len(neighbors) + 1, dtype=np.int64)
        for i, (_, d) in enumerate(neighbors):
            ptr[i + 1] = ptr[i] + len(d)
        nd = np.ascontiguousarray(np.concatenate([np.asarray(d, dtype=np.int32) for _, d in neighbors]) if neighbors else np.zeros(0, np.int32))
        h = _handle(load_library().jots_partition_create_from_tables(dim, order, gid.size, g.shape[0], _ptr(g), _ptr(ec), ba.size, _ptr(bg),
                                                                     _ptr(bc), _ptr(ba), _ptr(gid), lt.size, _ptr(lt), len(neighbors), _ptr(nr),
                                                                     _ptr(ptr), _ptr(nd), 0 if attributes is None else len(attributes),
                                                                     _ptr(None if attributes is None else np.ascontiguousarray(attributes, dtype=np.int32)),
                                                                     rank, nranks))
        return cls(None, None, nranks, rank, handle=h)

    def space(self):
        return Space(handle=load_library().jots_partition_space(self._h))

    def global_ids(self):
        g = np.empty(self.n_local, dtype=np.int64)
        e = np.empty(self.n_elem, dtype=np.int64)
        _chk(load_library().jots_partition_copy(self._h, _ptr(g), _ptr(e)))
        return g, e

    def neighbors(self):
        out = []
        for i in range(self.n_neighbors):
            r = C.c_int()
            n = load_library().jots_partition_neighbor(self._h, i, C.byref(r), None)
            idx = np.empty(n, dtype=np.int32)
            load_library().jots_partition_neighbor(self._h, i, C.byref(r), _ptr(idx))
            out.append((r.value, idx))
        return out

    def reduction(self):
        n = load_library().jots_partition_reduction(self._h, None, None, None)
        dof = np.empty(self.n_shared, dtype=np.int32)
        ptr = np.empty(self.n_shared + 1, dtype=np.int32)
        src = np.empty(n, dtype=np.int32)
        load_library().jots_partition_reduction(self._h, _ptr(dof), _ptr(ptr), _ptr(src))
        return dof, ptr, src

    def __del__(self):
        if getattr(self, "_h", None) and _LIB is not None:
            _LIB.jots_partition_destroy(self._h)
            self._h = None


class Comm:
    """NCCL communicator, one rank per GPU.  ``Comm.unique_id()`` on rank 0, broadcast by the caller."""

    @staticmethod
    def unique_id():
        buf = C.create_string_buffer(128)
        _chk(load_library().jots_comm_unique_id(buf))
        return buf.raw

    def __init__(self, uid, rank, nranks, device):
        self._h = _handle(load_library().jots_comm_create(uid, rank, nranks, device))
        self.rank, self.nranks = rank, nranks

    def __del__(self):
        if getattr(self, "_h", None) and _LIB is not None:
            _LIB.jots_comm_destroy(self._h)
            self._h = None


class Driver:
    """Config + the hot-path part of JOTSDriver::Run."""

    def __init__(self, ini_path, device=0, mesh=None, comm=None):
        lib = load_library()
        self._comm = comm
        if comm is not None:
            self._h = _handle(lib.jots_driver_create_dist(os.fsencode(ini_path), mesh._h if mesh is not None else None, device, comm._h))
        elif mesh is None:
            self._h = _handle(lib.jots_driver_create(os.fsencode(ini_path), device))
        else:
            self._h = _handle(lib.jots_driver_create_on_mesh(os.fsencode(ini_path), mesh._h, device))
        self.n = lib.jots_driver_size(self._h)
        self.ctx = Context(handle=lib.jots_driver_ctx(self._h), owner=self)
        self.op = Operator(self.ctx, None, handle=lib.jots_driver_op(self._h), owner=self)

    def run(self, max_steps=-1):
        done = C.c_int()
        _chk(load_library().jots_driver_run(self._h, max_steps, C.byref(done)))
        return done.value

    def get_u(self, out=None):
        u = np.empty(self.n) if out is None else out
        _chk(load_library().jots_driver_get_u(self._h, _ptr(u), _loc(u)))
        return u

    def set_u(self, u):
        _chk(load_library().jots_driver_set_u(self._h, _ptr(_f64(u)), _loc(u)))

    @property
    def time(self):
        return load_library().jots_driver_time(self._h)

    def global_ids(self):
        """(global DOF id of every local entry, number of leading owned entries)."""
        g = np.empty(self.n, dtype=np.int64)
        no = C.c_int64()
        _chk(load_library().jots_driver_copy_global_ids(self._h, _ptr(g), C.byref(no)))
        return g, no.value

    def space(self):
        return Space(handle=load_library().jots_driver_space(self._h))

    def __del__(self):
        if getattr(self, "_h", None) and _LIB is not None:
            self.ctx._h = None
            self.op._h = None
            _LIB.jots_driver_destroy(self._h)
            self._h = None
