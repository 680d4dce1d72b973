"""VTK output: the .vtu/.pvd files are well-formed and carry the data back."""
import xml.etree.ElementTree as ET

import numpy as np

import mupopp_b200 as mp
from mupopp_b200.io import PvdSeries, write_vtu


def test_vtu_roundtrip(tmp_path):
    m = mp.BoxMesh((0, 0, 0), (1, 1, 1), 2, 2, 2)
    c = np.sin(m.coords[:, 0])
    u = np.stack([m.coords[:, 1], -m.coords[:, 0], 0 * c], 1)
    fn = tmp_path / "a.vtu"
    write_vtu(str(fn), m.coords, m.cells, {"c0": c, "u": u})
    root = ET.parse(fn).getroot()
    piece = root.find("UnstructuredGrid/Piece")
    assert int(piece.get("NumberOfPoints")) == m.num_vertices() and int(piece.get("NumberOfCells")) == m.num_cells()
    arrs = {a.get("Name"): np.array(a.text.split(), dtype=float) for a in piece.find("PointData")}
    assert np.allclose(arrs["c0"], c, rtol=0, atol=1e-15) and np.allclose(arrs["u"].reshape(-1, 3), u, rtol=0, atol=1e-15)
    conn = np.array(piece.find("Cells")[0].text.split(), dtype=int).reshape(-1, 4)
    assert np.array_equal(conn, m.cells)
    assert set(piece.find("Cells")[2].text.split()) == {"10"}


def test_pvd_series(tmp_path):
    m = mp.UnitSquareMesh(2, 2)
    s = PvdSeries(str(tmp_path / "out" / "conc.pvd"))
    for k in range(3):
        s.write(m.coords, m.cells, {"c": np.full(m.num_vertices(), float(k))}, time=0.1 * k)
    root = ET.parse(tmp_path / "out" / "conc.pvd").getroot()
    ds = root.find("Collection").findall("DataSet")
    assert [d.get("file") for d in ds] == ["conc000000.vtu", "conc000001.vtu", "conc000002.vtu"]
    assert abs(float(ds[2].get("timestep")) - 0.2) < 1e-15
