"""Oracle (test infrastructure): H1 tensor-product space and dense-per-element assembly.

Restates the MFEM services the reference's conduction operators consume
(SURVEY.md section 2.2, App. A.1-A.3):

* ``H1_FECollection(p, dim)``: nodal Lagrange basis on Gauss-Lobatto points
  (/root/reference/src/jots_driver.cpp:194).
* default quadrature of Mass / Diffusion / MixedScalarWeakDivergence / Convection
  integrators on d-linear tensor elements: Gauss-Legendre with n = p+2 points per
  direction on hexes and n = p+1 on quads; BoundaryLFIntegrator: order p+1 ->
  n = (p+1)//2 + 1; boundary MassIntegrator: n = p+1 (App. A.2).
* element kernels with w = ip.weight*|J| (mass) and w = ip.weight/|J| with adj(J)
  (diffusion) - computed here per quadrature point from the multilinear geometry.
* coefficients are *nodal fields interpolated to quadrature points*
  (``GridFunctionCoefficient`` of ``PolynomialProperty``:
  /root/reference/src/material_property.cpp:31-33,36-49).

Everything is assembled the way the reference does it: dense element matrices ->
global CSR (``ParBilinearForm::Assemble`` / ``ParNonlinearForm::GetGradient``).
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

from . import mesh as M


# ----------------------------------------------------------------------------------------------
# 1-D rules and bases
# ----------------------------------------------------------------------------------------------

def gauss_legendre_01(n):
    x, w = np.polynomial.legendre.leggauss(n)
    return 0.5 * (x + 1.0), 0.5 * w


def gauss_lobatto_01(n):
    """n Gauss-Lobatto points on [0,1] (n >= 2)."""
    if n == 2:
        return np.array([0.0, 1.0])
    # interior points: roots of P'_{n-1}
    c = np.zeros(n)
    c[-1] = 1.0
    dc = np.polynomial.legendre.legder(c)
    r = np.sort(np.polynomial.legendre.legroots(dc).real)
    # Newton polish
    for _ in range(3):
        d1 = np.polynomial.legendre.legval(r, dc)
        d2 = np.polynomial.legendre.legval(r, np.polynomial.legendre.legder(dc))
        r = r - d1 / d2
    x = np.concatenate([[-1.0], r, [1.0]])
    x = 0.5 * (x - x[::-1])  # symmetrise
    return 0.5 * (x + 1.0)


def lagrange_tab(nodes, x):
    """B[q,d] = l_d(x_q), G[q,d] = l_d'(x_q) for the Lagrange basis on ``nodes``."""
    nodes = np.asarray(nodes, dtype=np.float64)
    x = np.asarray(x, dtype=np.float64)
    n = nodes.size
    B = np.ones((x.size, n))
    G = np.zeros((x.size, n))
    for d in range(n):
        den = 1.0
        for m in range(n):
            if m != d:
                den *= nodes[d] - nodes[m]
        for m in range(n):
            if m == d:
                continue
            B[:, d] *= x - nodes[m]
            term = np.ones_like(x)
            for r in range(n):
                if r != d and r != m:
                    term *= x - nodes[r]
            G[:, d] += term
        B[:, d] /= den
        G[:, d] /= den
    return B, G


def domain_quad_points(dim, p):
    """Gauss points per direction of the domain integrators (App. A.2)."""
    return p + 2 if dim == 3 else p + 1


def bdr_lf_quad_points(p):
    return (p + 1) // 2 + 1


def bdr_mass_quad_points(p):
    return p + 1


def tensor_tab(B, G, dim):
    """Phi[q,l], dPhi[q,l,a] on the tensor lattice, lower axis fastest in both q and l."""
    if dim == 1:
        return B.copy(), G[:, :, None].copy()
    if dim == 2:
        Phi = np.einsum("bj,ai->baji", B, B).reshape(B.shape[0] ** 2, -1)
        dx = np.einsum("bj,ai->baji", B, G).reshape(Phi.shape)
        dy = np.einsum("bj,ai->baji", G, B).reshape(Phi.shape)
        return Phi, np.stack([dx, dy], axis=2)
    Phi = np.einsum("ck,bj,ai->cbakji", B, B, B).reshape(B.shape[0] ** 3, -1)
    dx = np.einsum("ck,bj,ai->cbakji", B, B, G).reshape(Phi.shape)
    dy = np.einsum("ck,bj,ai->cbakji", B, G, B).reshape(Phi.shape)
    dz = np.einsum("ck,bj,ai->cbakji", G, B, B).reshape(Phi.shape)
    return Phi, np.stack([dx, dy, dz], axis=2)


def tensor_weights(w, dim):
    if dim == 1:
        return w.copy()
    if dim == 2:
        return np.einsum("b,a->ba", w, w).ravel()
    return np.einsum("c,b,a->cba", w, w, w).ravel()


# ----------------------------------------------------------------------------------------------
# space
# ----------------------------------------------------------------------------------------------

class H1Space:
    """H1 space of order p on a quad/hex mesh with all the tables assembly needs."""

    def __init__(self, mesh: M.Mesh, p: int, structured_numbering=None):
        self.mesh = mesh
        self.dim = dim = mesh.dim
        self.p = p
        self.D = D = p + 1
        self.nloc = D ** dim
        if structured_numbering is None:
            structured_numbering = mesh.brick is not None
        if structured_numbering and mesh.brick is not None:
            nx, ny, nz = mesh.brick[:3]
            self.gather, self.ndof = M.brick_lattice_numbering(nx, ny, nz, p)
        else:
            self.gather, self.ndof = M.lattice_numbering(mesh.elems, dim, p)
        self.nodes = gauss_lobatto_01(D)
        # domain quadrature
        self.Q = Q = domain_quad_points(dim, p)
        self.qx, self.qw = gauss_legendre_01(Q)
        self.B, self.G = lagrange_tab(self.nodes, self.qx)
        self.Phi, self.dPhi = tensor_tab(self.B, self.G, dim)
        self.W = tensor_weights(self.qw, dim)
        # geometry at quadrature points (multilinear map)
        self._geometry()
        # boundary
        self.bdr_elem, self.bdr_face = M.locate_boundary_faces(mesh)
        self._bdr_tables()

    # multilinear geometry ------------------------------------------------------------------
    def _vertex_shape(self, pts1d):
        """value/grad tables of the 2^dim multilinear vertex functions on a tensor lattice."""
        B1, G1 = lagrange_tab(np.array([0.0, 1.0]), pts1d)
        N, dN = tensor_tab(B1, G1, self.dim)  # columns are corner bits, lower axis fastest
        perm = [M.corner_vertex(self.dim, tuple((c >> a) & 1 for a in range(self.dim))) for c in range(2 ** self.dim)]
        Nv = np.zeros_like(N)
        dNv = np.zeros_like(dN)
        Nv[:, perm] = N
        dNv[:, perm, :] = dN
        return Nv, dNv

    def _geometry(self):
        X = self.mesh.element_vertices()  # [e, v, dim]
        _, dNv = self._vertex_shape(self.qx)
        # J[e,q,i,a] = d x_i / d xi_a
        self.J = np.einsum("evi,qva->eqia", X, dNv)
        self.detJ = np.linalg.det(self.J)
        self.invJ = np.linalg.inv(self.J)  # [e,q,a,i] = d xi_a / d x_i
        if (self.detJ <= 0).any():
            raise ValueError("non-positive element Jacobian")

    def node_coordinates(self):
        """[ndof, dim] physical coordinates of the nodes."""
        Nv, _ = self._vertex_shape(self.nodes)
        X = self.mesh.element_vertices()
        pts = np.einsum("lv,evi->eli", Nv, X)
        out = np.zeros((self.ndof, self.dim))
        out[self.gather.ravel()] = pts.reshape(-1, self.dim)
        return out

    # boundary tables -----------------------------------------------------------------------
    def _bdr_tables(self):
        dim, p = self.dim, self.p
        nb = self.mesh.bdr.shape[0]
        nfl = self.D ** (dim - 1)
        self.bdr_gather = np.empty((nb, nfl), dtype=np.int64)
        for f in range(2 * dim):
            sel = self.bdr_face == f
            if sel.any():
                self.bdr_gather[sel] = self.gather[self.bdr_elem[sel]][:, M.face_lattice_nodes(dim, p, f)]

    def face_measure(self, pts1d):
        """Surface Jacobian [nb, nq_face] of every boundary face at a tensor rule."""
        dim = self.dim
        X = self.mesh.element_vertices()
        nb = self.mesh.bdr.shape[0]
        nqf = pts1d.size ** (dim - 1)
        out = np.empty((nb, nqf))
        B1, G1 = lagrange_tab(np.array([0.0, 1.0]), pts1d)
        for f in range(2 * dim):
            sel = np.nonzero(self.bdr_face == f)[0]
            if sel.size == 0:
                continue
            lv = M.face_vertices_local(dim, f)  # corners, lower free axis fastest
            Xf = X[self.bdr_elem[sel]][:, lv]  # [n, 2^(dim-1), dim]
            if dim == 2:
                t = np.einsum("nvi,qv->nqi", Xf, G1)
                out[sel] = np.linalg.norm(t, axis=2)
            else:
                _, dN = tensor_tab(B1, G1, 2)
                t = np.einsum("nvi,qva->nqia", Xf, dN)
                out[sel] = np.linalg.norm(np.cross(t[..., 0], t[..., 1]), axis=2)
        return out

    def essential_dofs(self, attrs):
        """Sorted unique dofs on boundary elements whose attribute is in ``attrs``
        (``GetEssentialTrueDofs``, /root/reference/src/jots_iterator.cpp:30)."""
        sel = np.isin(self.mesh.bdr_attr, np.asarray(list(attrs), dtype=np.int64))
        if not sel.any():
            return np.zeros(0, dtype=np.int64)
        return np.unique(self.bdr_gather[sel])

    # helpers ---------------------------------------------------------------------------------
    def at_quad(self, nodal):
        """Interpolate a nodal field to the domain quadrature points: [e,q]."""
        return nodal[self.gather] @ self.Phi.T

    def grad_at_quad(self, nodal):
        """Physical gradient of a nodal field at quadrature points: [e,q,dim]."""
        gref = np.einsum("el,qla->eqa", nodal[self.gather], self.dPhi)
        return np.einsum("eqa,eqai->eqi", gref, self.invJ)

    def phys_dPhi(self):
        """Physical basis gradients [e,q,l,i] (cached while it stays below ~256 MB)."""
        if getattr(self, "_dP", None) is not None:
            return self._dP
        dP = np.einsum("qla,eqai->eqli", self.dPhi, self.invJ)
        if dP.nbytes <= (256 << 20):
            self._dP = dP
        return dP

    def _scatter_matrix(self, elmat):
        ne, n, _ = elmat.shape
        rows = np.repeat(self.gather, n, axis=1).ravel()
        cols = np.tile(self.gather, (1, n)).ravel()
        A = sp.coo_matrix((elmat.ravel(), (rows, cols)), shape=(self.ndof, self.ndof))
        return A.tocsr()

    def _scatter_vector(self, elvec, gather=None):
        g = self.gather if gather is None else gather
        return np.bincount(g.ravel(), weights=elvec.ravel(), minlength=self.ndof)

    # ------------------------------------------------------------------------------------------
    # bilinear forms (element matrices -> CSR)
    # ------------------------------------------------------------------------------------------
    def _coef(self, c):
        """coefficient -> [e,q]; scalars are constants, arrays are nodal fields."""
        if np.isscalar(c):
            return np.full(self.detJ.shape, float(c))
        return self.at_quad(np.asarray(c, dtype=np.float64))

    def mass_matrix(self, coef=1.0):
        """``MassIntegrator(coef)``: M_ij = sum_q w_q |J| c_q phi_i phi_j."""
        wq = self.W[None, :] * self.detJ * self._coef(coef)
        el = np.einsum("eq,qi,qj->eij", wq, self.Phi, self.Phi, optimize=True)
        return self._scatter_matrix(el)

    def diffusion_matrix(self, coef=1.0):
        """``DiffusionIntegrator(coef)``: K_ij = sum_q w_q |J| c_q grad phi_i . grad phi_j."""
        wq = self.W[None, :] * self.detJ * self._coef(coef)
        dP = self.phys_dPhi()
        el = np.einsum("eq,eqli,eqmi->elm", wq, dP, dP, optimize=True)
        return self._scatter_matrix(el)

    def weakdiv_term_matrix(self, coef, state):
        """+ int coef_q phi_m (grad z . grad phi_i): the second Jacobian term of
        JOTSNonlinearDiffusionIntegrator::AssembleElementGrad
        (/root/reference/src/jots_nlfis.cpp:33-39; minus MixedScalarWeakDivergenceIntegrator)."""
        wq = self.W[None, :] * self.detJ * self._coef(coef)
        gz = self.grad_at_quad(state)
        dP = self.phys_dPhi()
        t = np.einsum("eqli,eqi->eql", dP, gz)
        el = np.einsum("eq,eql,qm->elm", wq, t, self.Phi, optimize=True)
        return self._scatter_matrix(el)

    def convection_matrix(self, coef, state):
        """``ConvectionIntegrator(coef * grad z)``: C_im = int phi_i coef_q grad z . grad phi_m
        (/root/reference/src/jots_nlfis.cpp:80,88)."""
        wq = self.W[None, :] * self.detJ * self._coef(coef)
        gz = self.grad_at_quad(state)
        dP = self.phys_dPhi()
        t = np.einsum("eqmi,eqi->eqm", dP, gz)
        el = np.einsum("eq,ql,eqm->elm", wq, self.Phi, t, optimize=True)
        return self._scatter_matrix(el)

    def gradsq_mass_matrix(self, coef, state):
        """``MassIntegrator(coef * |grad z|^2)`` (/root/reference/src/jots_nlfis.cpp:92)."""
        gz = self.grad_at_quad(state)
        wq = self.W[None, :] * self.detJ * self._coef(coef) * np.einsum("eqi,eqi->eq", gz, gz)
        el = np.einsum("eq,qi,qj->eij", wq, self.Phi, self.Phi, optimize=True)
        return self._scatter_matrix(el)

    # ------------------------------------------------------------------------------------------
    # nonlinear residual vectors (element vectors)
    # ------------------------------------------------------------------------------------------
    def diffusion_action(self, coef, u):
        """``DiffusionIntegrator(coef)::AssembleElementVector``: r_i = int c_q grad u . grad phi_i
        (/root/reference/src/jots_nlfis.cpp:16-22)."""
        wq = self.W[None, :] * self.detJ * self._coef(coef)
        gu = self.grad_at_quad(u)
        dP = self.phys_dPhi()
        el = np.einsum("eq,eqi,eqli->el", wq, gu, dP, optimize=True)
        return self._scatter_vector(el)

    def convection_action(self, coef, u):
        """int phi_i coef_q |grad u|^2 (/root/reference/src/jots_nlfis.cpp:75-81)."""
        gu = self.grad_at_quad(u)
        wq = self.W[None, :] * self.detJ * self._coef(coef) * np.einsum("eqi,eqi->eq", gu, gu)
        el = wq @ self.Phi
        return self._scatter_vector(el)

    # ------------------------------------------------------------------------------------------
    # boundary forms
    # ------------------------------------------------------------------------------------------
    def _face_tab(self, nq):
        x, w = gauss_legendre_01(nq)
        B, _ = lagrange_tab(self.nodes, x)
        if self.dim == 2:
            return x, w, B
        Phi = np.einsum("bj,ai->baji", B, B).reshape(nq * nq, -1)
        return x, np.einsum("b,a->ba", w, w).ravel(), Phi

    def face_nodal_values(self, g_by_attr):
        """[nb, nfl] values of a piecewise boundary coefficient at the face nodes: {attr: scalar} (uniform BC values) or
        {attr: nodal field of length ndof} (a GridFunctionCoefficient, as PreciceBC's coeff_gf:
        /root/reference/src/boundary_condition.cpp:134-147); unmapped attributes give 0."""
        out = np.zeros(self.bdr_gather.shape)
        for attr, val in g_by_attr.items():
            sel = self.mesh.bdr_attr == attr
            if np.isscalar(val):
                out[sel] = val
            else:
                out[sel] = np.asarray(val)[self.bdr_gather[sel]]
        return out

    def neumann_vector(self, g_by_attr, divisor=None):
        """``BoundaryLFIntegrator(PWCoefficient)``: b_i = sum_faces int g phi_i dS, rule order
        p+1 (/root/reference/src/jots_iterator.cpp:33,39-43).  ``g_by_attr``: {attr: value or nodal
        field}; unmapped attributes contribute 0.  ``divisor``: optional nodal field (rho*C is formed by
        the caller as a *ratio of interpolants*: g / (rho_h C_h), see
        /root/reference/src/nl_conduction_operator.cpp:104-107) given as tuple (rho, C) of
        scalars / nodal fields."""
        x, w, Phi = self._face_tab(bdr_lf_quad_points(self.p))
        meas = self.face_measure(x)
        coef = self.face_nodal_values(g_by_attr) @ Phi.T
        if divisor is not None:
            coef = coef / self._face_coef_product(divisor, Phi)
        el = (coef * meas * w[None, :]) @ Phi
        return self._scatter_vector(el, self.bdr_gather)

    def _face_coef(self, c, Phi):
        if np.isscalar(c):
            return np.full((self.mesh.bdr.shape[0], Phi.shape[0]), float(c))
        c = np.asarray(c)
        if c.ndim == 2:                       # values at the face nodes (face_nodal_values)
            return c @ Phi.T
        return c[self.bdr_gather] @ Phi.T

    # preCICE boundary data path (/root/reference/src/boundary_condition.cpp:77-206) ------------------
    def bdr_nodes(self, attr):
        """(dofs, coords) of the nodes of every boundary element of ``attr``, boundary element by boundary element
        (duplicates across neighbouring faces kept, as PreciceBC's bdr_dof_indices / coords)."""
        sel = np.nonzero(self.mesh.bdr_attr == attr)[0]
        dofs = self.bdr_gather[sel].ravel()
        return dofs, self.node_coordinates()[dofs]

    def wall_heat_flux(self, attr, u, k_of_T):
        """PreciceIsothermalBC::RetrieveWriteData (:152-191): -k(T) grad T . n / |n| at every node of every boundary
        element of ``attr`` -- gradient of the ADJACENT element's field at that node, outward normal of the face."""
        dim, D = self.dim, self.D
        sel = np.nonzero(self.mesh.bdr_attr == attr)[0]
        _, Gn = lagrange_tab(self.nodes, self.nodes)            # Gn[i, m] = l_m'(x_i)
        Bn = np.eye(D)
        _, dPhi_n = tensor_tab(Bn, Gn, dim)                      # [node, local dof, axis]: reference gradient at the nodes
        _, dNv = self._vertex_shape(self.nodes)
        X = self.mesh.element_vertices()
        out = []
        for b in sel:
            e, f = self.bdr_elem[b], self.bdr_face[b]
            axis, side = f // 2, f % 2
            ue = u[self.gather[e]]
            J = np.einsum("vi,qva->qia", X[e], dNv)              # [node, i, a]
            invJ = np.linalg.inv(J)                              # [node, a, i]
            gref = np.einsum("qla,l->qa", dPhi_n, ue)
            gphys = np.einsum("qai,qa->qi", invJ, gref)
            for l in M.face_lattice_nodes(dim, self.p, f):
                n = invJ[l, axis, :] * (1.0 if side else -1.0)   # grad xi_axis: normal of the face xi_axis = const, outward
                n = n / np.linalg.norm(n)
                out.append(-k_of_T(ue[l]) * float(gphys[l] @ n))
        return np.array(out)

    def _face_coef_product(self, cs, Phi):
        out = 1.0
        for c in cs:
            out = out * self._face_coef(c, Phi)
        return out

    def bdr_coef_vector(self, coef_fn, nq):
        """int_dOmega lambda_q phi_i with lambda given by coef_fn(face_interp) -> [nb,nq]."""
        x, w, Phi = self._face_tab(nq)
        meas = self.face_measure(x)
        lam = coef_fn(lambda c: self._face_coef(c, Phi))
        el = (lam * meas * w[None, :]) @ Phi
        return self._scatter_vector(el, self.bdr_gather)

    def bdr_mass_matrix(self, coef_fn):
        """int_dOmega lambda'_q phi_i phi_j (boundary MassIntegrator, n = p+1;
        /root/reference/src/jots_nlfis.cpp:56-59)."""
        x, w, Phi = self._face_tab(bdr_mass_quad_points(self.p))
        meas = self.face_measure(x)
        lam = coef_fn(lambda c: self._face_coef(c, Phi))
        el = np.einsum("bq,qi,qj->bij", lam * meas * w[None, :], Phi, Phi, optimize=True)
        n = el.shape[1]
        rows = np.repeat(self.bdr_gather, n, axis=1).ravel()
        cols = np.tile(self.bdr_gather, (1, n)).ravel()
        return sp.coo_matrix((el.ravel(), (rows, cols)), shape=(self.ndof, self.ndof)).tocsr()

    # ------------------------------------------------------------------------------------------
    # error functional of the reference's tests
    # ------------------------------------------------------------------------------------------
    def l2_error(self, u, exact_fn):
        """``GridFunction::ComputeL2Error`` (rule order 2p+3 -> n = p+2 points/direction)."""
        n = (2 * self.p + 3) // 2 + 1
        x, w = gauss_legendre_01(n)
        B, _ = lagrange_tab(self.nodes, x)
        Phi, _ = tensor_tab(B, np.zeros_like(B), self.dim)
        W = tensor_weights(w, self.dim)
        Nv, dNv = self._vertex_shape(x)
        X = self.mesh.element_vertices()
        pts = np.einsum("qv,evi->eqi", Nv, X)
        J = np.einsum("evi,qva->eqia", X, dNv)
        det = np.linalg.det(J)
        uh = u[self.gather] @ Phi.T
        ue = exact_fn(pts.reshape(-1, self.dim)).reshape(uh.shape)
        return float(np.sqrt(np.sum(W[None, :] * det * (uh - ue) ** 2)))


def eliminate_rows_cols(A, ess, diag=1.0):
    """``EliminateRowsCols`` / ``EliminateBC(DIAG_ONE)``: zero essential rows and columns,
    put ``diag`` on the diagonal (App. A.3)."""
    A = A.tocsr(copy=True)
    n = A.shape[0]
    keep = np.ones(n)
    keep[ess] = 0.0
    Dk = sp.diags(keep)
    A = Dk @ A @ Dk
    d = np.zeros(n)
    d[ess] = diag
    return (A + sp.diags(d)).tocsr()
