"""On-disk formats either side of the hot path: Plot3D grid writer (the reader is grid.Geometry.read) and the
`.vts` export of a solution (same point layout and array names as the reference's
dgfem/visualization.py:52-128, which goes through pyevtk's gridToVTK).

The exporter writes VTK XML StructuredGrid files with raw appended binary data (what pyevtk emits), straight
from NumPy buffers -- no per-point Python loops, so the 38 MDOF cases stream at disk speed.
"""
import os
import struct

import numpy as np


def write_plot3d(path, xn, yn):
    """Single-block Plot3D file as dgfem/grid.py:26-63 reads it: Fortran unformatted records (little-endian
    4-byte markers) [nblocks=1] [il, jl, kl=1] [x | y | z as float64, i fastest].
    xn, yn: node coordinates in file order [jl][il] (NumPy arrays or CPU/CUDA tensors)."""
    xn = np.ascontiguousarray(xn.cpu().numpy() if hasattr(xn, "cpu") else xn, dtype="<f8")
    yn = np.ascontiguousarray(yn.cpu().numpy() if hasattr(yn, "cpu") else yn, dtype="<f8")
    jl, il = xn.shape
    with open(path, "wb") as f:
        def record(payload_bytes, writer):
            f.write(struct.pack("<I", payload_bytes))
            writer()
            f.write(struct.pack("<I", payload_bytes))
        record(4, lambda: f.write(struct.pack("<i", 1)))
        record(12, lambda: f.write(struct.pack("<iii", il, jl, 1)))

        def coords():
            xn.tofile(f)
            yn.tofile(f)
            np.zeros(il * jl, dtype="<f8").tofile(f)
        nbytes = 3 * il * jl * 8
        if nbytes >= 2 ** 32:
            raise ValueError("grid too large for 4-byte Fortran record markers")
        record(nbytes, coords)
    return path


def _vts(path, x, y, point_data=None):
    """x, y: [nx, ny] point coordinates; point_data: name -> [nx, ny] array or a tuple of three (vector)."""
    nx, ny = x.shape
    arrays = []          # (xml attributes, bytes-like producer)
    pts = np.empty((ny, nx, 3), dtype="<f8")          # VTK point order: i fastest
    pts[:, :, 0] = x.T
    pts[:, :, 1] = y.T
    pts[:, :, 2] = 0.0
    blobs = [pts]
    pd_xml = []
    offset = 8 + pts.nbytes
    for name, data in (point_data or {}).items():
        if isinstance(data, tuple):
            v = np.empty((ny, nx, 3), dtype="<f8")
            for c in range(3):
                v[:, :, c] = np.asarray(data[c]).reshape(nx, ny).T
            ncomp = 3
        else:
            v = np.ascontiguousarray(np.asarray(data, dtype="<f8").reshape(nx, ny).T)
            ncomp = 1
        pd_xml.append(f'<DataArray type="Float64" Name="{name}" NumberOfComponents="{ncomp}" format="appended" '
                      f'offset="{offset}"/>')
        blobs.append(v)
        offset += 8 + v.nbytes
    ext = f"0 {nx - 1} 0 {ny - 1} 0 0"
    head = ('<?xml version="1.0"?>\n<VTKFile type="StructuredGrid" version="1.0" byte_order="LittleEndian" '
            'header_type="UInt64">\n'
            f'<StructuredGrid WholeExtent="{ext}">\n<Piece Extent="{ext}">\n<PointData>\n' + "\n".join(pd_xml) +
            '\n</PointData>\n<CellData>\n</CellData>\n<Points>\n'
            '<DataArray type="Float64" Name="points" NumberOfComponents="3" format="appended" offset="0"/>\n'
            '</Points>\n</Piece>\n</StructuredGrid>\n<AppendedData encoding="raw">\n_')
    if not path.endswith(".vts"):
        path += ".vts"
    with open(path, "wb") as f:
        f.write(head.encode())
        for b in blobs:
            f.write(struct.pack("<Q", b.nbytes))
            b.tofile(f)
        f.write(b"\n</AppendedData>\n</VTKFile>\n")
    return path


def grid_to_vtk(filename, x, y):
    """dgfem/visualization.py:52-65: the grid nodes x[il, jl], y[il, jl] as a structured grid."""
    return _vts(os.path.join(os.getcwd(), filename), np.asarray(x), np.asarray(y))


def element_node_arrays(xn, yn, Ni, Nj, Pg):
    """x_el, y_el [Ni, Nj, N1, N1] (node a along i, c along j) from the file-order node arrays [jl][il]."""
    N1 = Pg + 1
    ii = (np.arange(Ni)[:, None] * Pg + np.arange(N1)[None, :])        # [Ni, a]
    jj = (np.arange(Nj)[:, None] * Pg + np.arange(N1)[None, :])        # [Nj, c]
    x_el = xn[jj[None, :, None, :], ii[:, None, :, None]]
    y_el = yn[jj[None, :, None, :], ii[:, None, :, None]]
    return x_el, y_el


def _spread(a):
    """[Ni, Nj, N1, N1] -> [Ni*N1, Nj*N1] (dgfem/visualization.py:73-74: element nodes side by side, interface
    nodes duplicated)."""
    return a.transpose(0, 2, 1, 3).reshape(a.shape[0] * a.shape[2], a.shape[1] * a.shape[3])


def elements_to_vtk(filename, x_el, y_el, problem="Poisson", point_data=None):
    """dgfem/visualization.py:67-117.  x_el, y_el, point_data[name]: [Ni, Nj, N1, N1]."""
    pd = {}
    if point_data:
        vel, vel_exact = [None, None], [None, None]
        for key, data in sorted(point_data.items(), key=lambda kv: kv[0].lower(), reverse=True):
            data = _spread(np.asarray(data))
            if problem == "Stokes" and key in ("u", "v"):
                vel[("u", "v").index(key)] = data
            elif problem == "Stokes" and key in ("u_exact", "v_exact"):
                vel_exact[("u_exact", "v_exact").index(key)] = data
            else:
                pd[key] = data
        if problem == "Stokes":
            for name, comps in (("velocity", vel), ("velocity_exact", vel_exact)):
                if comps[0] is not None and comps[1] is not None:
                    pd[name] = (comps[0], comps[1], np.zeros_like(comps[0]))
    return _vts(os.path.join(os.getcwd(), filename), _spread(np.asarray(x_el)), _spread(np.asarray(y_el)), pd)


def nodal_to_elements(u_nodal, Ni, Nj, Pg):
    """[N, ng] (element m = j*Ni + i, node a + N1*c) -> [Ni, Nj, N1(a), N1(c)] (dgfem.py:203-205 layout)."""
    N1 = Pg + 1
    a = u_nodal.cpu().numpy() if hasattr(u_nodal, "cpu") else np.asarray(u_nodal)
    return a.reshape(Nj, Ni, N1, N1).transpose(1, 0, 3, 2)


def modal_to_vtk(filename, dgfem, u_nodal, u_exact_nodes):
    """The export of DGFEM.solve (dgfem/dgfem.py:236-247, Poisson): phi, phi_exact, abs_error_phi at the element
    nodes.  u_nodal: [N, ng] device tensor (dgb_nodal_error); u_exact_nodes: [jl, il] exact solution at the nodes."""
    g = dgfem.grids[-1]
    xn, yn = (t.cpu().numpy() for t in dgfem.geometry.device_nodes())
    x_el, y_el = element_node_arrays(xn, yn, g.Ni, g.Nj, g.P_grid)
    ex = u_exact_nodes.cpu().numpy() if hasattr(u_exact_nodes, "cpu") else np.asarray(u_exact_nodes)
    ex_el, _ = element_node_arrays(ex, ex, g.Ni, g.Nj, g.P_grid)
    phi = nodal_to_elements(u_nodal, g.Ni, g.Nj, g.P_grid)
    return elements_to_vtk(filename, x_el, y_el, "Poisson",
                           {"phi_exact": ex_el, "phi": phi, "abs_error_phi": np.abs(phi - ex_el)})
