"""VTK XML reader + the VTK Lagrange cell node order  --  TEST INFRASTRUCTURE ONLY (checks jots_b200/csrc/paraview.cpp).

The reference writes its ParaView output through mfem::ParaViewDataCollection (src/output_manager.cpp:23-28,81-87: binary,
high-order, levels of detail = FE order); MFEM is not in the reference tree, so the format is restated here from the VTK XML
file format and the published node order of vtkLagrangeQuadrilateral / vtkLagrangeHexahedron (VTK >= 9 / file version 2.2):
**unpinned against MFEM's bytes**.  The node order is built constructively (corner list, edge list, face list, interior),
not from the index formula the product uses, so the two are independent statements of the same convention.
"""
import base64
import struct
import xml.etree.ElementTree as ET

import numpy as np

_NP = {"Float64": np.float64, "Float32": np.float32, "Int32": np.int32, "UInt8": np.uint8, "Int64": np.int64}


def lagrange_nodes(dim, p):
    """Lattice coordinates (i, j[, k]) in 0..p of the VTK Lagrange cell's nodes, in VTK order."""
    r = range(1, p)
    if dim == 2:
        out = [(0, 0), (p, 0), (p, p), (0, p)]
        out += [(i, 0) for i in r] + [(p, j) for j in r] + [(i, p) for i in r] + [(0, j) for j in r]
        out += [(i, j) for j in r for i in r]
        return out
    out = [(0, 0, 0), (p, 0, 0), (p, p, 0), (0, p, 0), (0, 0, p), (p, 0, p), (p, p, p), (0, p, p)]
    for k in (0, p):      # the four edges of the bottom face, then of the top face: every edge in increasing coordinate
        out += [(i, 0, k) for i in r] + [(p, j, k) for j in r] + [(i, p, k) for i in r] + [(0, j, k) for j in r]
    for (i, j) in ((0, 0), (p, 0), (0, p), (p, p)):      # vertical edges (VTK 9 order)
        out += [(i, j, k) for k in r]
    for i in (0, p):
        out += [(i, j, k) for k in r for j in r]
    for j in (0, p):
        out += [(i, j, k) for k in r for i in r]
    for k in (0, p):
        out += [(i, j, k) for j in r for i in r]
    out += [(i, j, k) for k in r for j in r for i in r]
    return out


def _block(text, dtype):
    text = "".join(text.split())
    (nbytes,) = struct.unpack("<I", base64.b64decode(text[:8]))
    raw = base64.b64decode(text[8:])
    assert len(raw) == nbytes, (len(raw), nbytes)
    return np.frombuffer(raw, dtype=dtype)


def read_vtu(path):
    root = ET.parse(path).getroot()
    assert root.attrib["type"] == "UnstructuredGrid" and root.attrib["byte_order"] == "LittleEndian"
    piece = root.find("UnstructuredGrid").find("Piece")
    out = {"n_points": int(piece.attrib["NumberOfPoints"]), "n_cells": int(piece.attrib["NumberOfCells"]), "version": root.attrib["version"],
           "point_data": {}, "cell_data": {}}
    da = piece.find("Points").find("DataArray")
    assert da.attrib["format"] == "binary"
    out["points"] = _block(da.text, _NP[da.attrib["type"]]).reshape(-1, int(da.attrib["NumberOfComponents"]))
    for da in piece.find("Cells").findall("DataArray"):
        out[da.attrib["Name"]] = _block(da.text, _NP[da.attrib["type"]])
    for da in piece.find("PointData").findall("DataArray"):
        out["point_data"][da.attrib["Name"]] = _block(da.text, _NP[da.attrib["type"]])
    for da in piece.find("CellData").findall("DataArray"):
        out["cell_data"][da.attrib["Name"]] = _block(da.text, _NP[da.attrib["type"]])
    return out


def read_pvd(path):
    root = ET.parse(path).getroot()
    assert root.attrib["type"] == "Collection"
    return [(float(d.attrib["timestep"]), d.attrib["file"]) for d in root.find("Collection").findall("DataSet")]


def read_pvtu(path):
    root = ET.parse(path).getroot()
    assert root.attrib["type"] == "PUnstructuredGrid"
    g = root.find("PUnstructuredGrid")
    return {"pieces": [p.attrib["Source"] for p in g.findall("Piece")],
            "point_arrays": [a.attrib["Name"] for a in g.find("PPointData").findall("PDataArray")],
            "cell_arrays": [a.attrib["Name"] for a in g.find("PCellData").findall("PDataArray")]}
