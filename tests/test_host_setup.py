"""CPU tests of the host setup library (include/nsb_host.h): meshes, gmsh I/O,
Taylor-Hood numbering and block sparsity pattern (bit-exact against the
oracle's independent restatement), Dirichlet / obstacle-face lists, partition."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import CASES, ROOT, make_case

ALL_MESHES = [("2d-cylinder", 0.05), ("3d-square", 0.1), ("3d-cylinder", 0.1), ("naca2412", 0.1),
              ("channel2d", 0.1), ("channel3d", 0.15)]
EXACT_VOLUME = {"3d-square": (2.5 * 0.41 - 0.01) * 0.41, "channel2d": 2.2 * 0.41, "channel3d": 2.5 * 0.41 * 0.41}


def _faces(cells, dim):
    loc = [(0, 1), (1, 2), (2, 0)] if dim == 2 else [(0, 1, 2), (1, 0, 3), (0, 2, 3), (2, 1, 3)]
    return np.sort(np.concatenate([cells[:, list(f)] for f in loc]), axis=1)


@pytest.mark.parametrize("name,h", ALL_MESHES)
def test_generated_mesh_is_valid(pkg, name, h):
    prob = pkg.Problem.generate(name, h)
    s = prob.sizes()
    dim = s["dim"]
    xyz = prob.array("xyz").reshape(-1, dim)
    cells = prob.array("cells").reshape(-1, dim + 1)
    vol = np.linalg.det(xyz[cells[:, 1:]] - xyz[cells[:, :1]]) / (2 if dim == 2 else 6)
    assert vol.min() > 0, "all cells positively oriented"
    if name in EXACT_VOLUME:
        assert abs(vol.sum() - EXACT_VOLUME[name]) < 1e-12
    # conforming: every facet belongs to one (boundary) or two (interior) cells
    f, cnt = np.unique(_faces(cells, dim), axis=0, return_counts=True)
    assert cnt.max() == 2
    bnd = f[cnt == 1]
    tagged = np.unique(np.sort(prob.array("bfaces").reshape(-1, dim), axis=1), axis=0)
    assert np.array_equal(bnd, tagged), "every boundary facet is tagged exactly once"
    assert set(np.unique(prob.array("bids"))) <= {0, 1, 2, 3, 4}
    assert len(np.unique(cells)) == s["n_verts"], "no unused vertices"


def test_mesh_generation_is_deterministic(pkg):
    a = pkg.Problem.generate("3d-cylinder", 0.1)
    b = pkg.Problem.generate("3d-cylinder", 0.1)
    for k in ("xyz", "cells", "bfaces", "bids"):
        assert np.array_equal(a.array(k), b.array(k))


@pytest.mark.parametrize("name,h", [("2d-cylinder", 0.08), ("3d-square", 0.15)])
def test_msh_roundtrip(pkg, tmp_path, name, h):
    a = pkg.Problem.generate(name, h)
    path = str(tmp_path / "m.msh")
    a.write_msh(path)
    b = pkg.Problem.read_msh(path, a.sizes()["dim"])
    for k in ("xyz", "cells", "bfaces", "bids"):
        assert np.array_equal(a.array(k), b.array(k)), k


def test_msh_v41_reader_and_errors(pkg, tmp_path):
    # two triangles, one tagged boundary edge (physical 3), one unused node, one negatively oriented cell
    txt = """$MeshFormat
4.1 0 8
$EndMeshFormat
$Entities
0 1 1 0
7 0 0 0 0 1 0 1 3 0
1 0 0 0 1 1 0 1 10 0
$EndEntities
$Nodes
1 5 1 5
2 1 0 5
1
2
3
4
5
0 0 0
1 0 0
1 1 0
0 1 0
9 9 0
$EndNodes
$Elements
2 3 1 3
1 7 1 1
1 4 1
2 1 2 2
2 1 2 3
3 1 4 3
$EndElements
"""
    p = tmp_path / "v41.msh"
    p.write_text(txt)
    m = pkg.Problem.read_msh(str(p), 2)
    s = m.sizes()
    assert (s["n_verts"], s["n_cells"], s["n_bfaces"]) == (4, 2, 1)
    assert m.array("bids")[0] == 3
    xyz, cells = m.array("xyz").reshape(-1, 2), m.array("cells").reshape(-1, 3)
    assert (np.linalg.det(xyz[cells[:, 1:]] - xyz[cells[:, :1]]) > 0).all()
    with pytest.raises(pkg.HostError):
        pkg.Problem.read_msh(str(tmp_path / "missing.msh"), 2)
    with pytest.raises(pkg.HostError):
        pkg.Problem.generate("no-such-mesh", 0.1)


@pytest.mark.parametrize("key", list(CASES))
def test_numbering_and_pattern_match_oracle_bit_exact(pkg, oracle_mod, key):
    """DoF numbering and block pattern (A.5) of the product's
    host library vs the oracle's independent std::map / per-row-set version."""
    prob, orc, dim, nu, um = make_case(pkg, oracle_mod, key)
    assert np.array_equal(orc.cell_dofs(), prob.array("cell_dofs"))
    for b in ("a00", "a01", "a10", "s"):
        rp, ci = orc.pattern(b)
        assert np.array_equal(rp, prob.array(b + ".rowptr")), b
        assert np.array_equal(ci, prob.array(b + ".colind")), b
    s = prob.sizes()
    assert s["n_u"] == orc.n_u and s["n_p"] == orc.n_p
    # canonical structure: velocity dof = dim*node + c, A00 = nodes (x) ones(dim, dim)
    nrp, nci = prob.array("nodes.rowptr"), prob.array("nodes.colind")
    rp, ci = prob.array("a00.rowptr"), prob.array("a00.colind")
    assert rp[-1] == nrp[-1] * dim * dim
    A = 7 % s["n_nodes"]
    exp = (dim * nci[nrp[A]:nrp[A + 1]][:, None] + np.arange(dim)[None, :]).ravel()
    for c in range(dim):
        assert np.array_equal(ci[rp[dim * A + c]:rp[dim * A + c + 1]], exp)


@pytest.mark.parametrize("key", list(CASES))
def test_dirichlet_list_matches_oracle(pkg, oracle_mod, key):
    prob, orc, dim, nu, um = make_case(pkg, oracle_mod, key)
    orc.assemble(0.01)
    d, v = orc.bc()
    assert np.array_equal(d, prob.array("bc.dofs"))
    assert np.allclose(v, prob.array("bc.values"), rtol=0, atol=1e-15)
    assert prob.mean_velocity(0.3) == orc.mean_velocity(0.3)


def test_time_dependent_inlet_factor(pkg):
    prob = pkg.Problem.generate("channel2d", 0.2).build(inlet=(pkg.INLET_PARABOLIC, 1.5, 0.41, 1))
    assert prob.inlet_time_factor(0.0) == 0.0  # get_mean_vel() = 0 at t = 0
    assert abs(prob.inlet_time_factor(4.0) - 1.0) < 1e-15
    assert abs(prob.mean_velocity(4.0) - 1.0) < 1e-15


@pytest.mark.parametrize("name,h", [("2d-cylinder", 0.05), ("3d-cylinder", 0.1), ("3d-square", 0.1)])
def test_force_faces_close_the_obstacle(pkg, name, h):
    prob = pkg.Problem.generate(name, h).build()
    dim = prob.sizes()["dim"]
    n = prob.array("ff.normal").reshape(-1, dim)
    m = prob.array("ff.measure")
    assert np.allclose(np.linalg.norm(n, axis=1), 1.0)
    # closed curve / prismatic surface: integral of the normal vanishes in x and y
    assert np.abs((n * m[:, None]).sum(axis=0)[:2]).max() < 1e-12
    if name == "3d-square":
        assert abs(m.sum() - 4 * 0.1 * 0.41) < 1e-12
    # normals point out of the fluid, i.e. into the obstacle: towards its centre
    xyz = prob.array("xyz").reshape(-1, dim)
    cells = prob.array("cells").reshape(-1, dim + 1)
    cen = xyz[cells[prob.array("ff.cell")]].mean(axis=1)
    c0 = np.array([0.2, 0.2]) if name == "2d-cylinder" else np.array([0.45 if name == "3d-cylinder" else 0.5, 0.2])
    assert (((c0 - cen[:, :2]) * n[:, :2]).sum(axis=1) > 0).all()


@pytest.mark.parametrize("parts", [2, 4, 8])
def test_partition_is_balanced(pkg, parts):
    prob = pkg.Problem.generate("3d-cylinder", 0.1)
    p = prob.partition(parts)
    cnt = np.bincount(p, minlength=parts)
    assert cnt.min() > 0 and cnt.max() - cnt.min() <= parts
    assert np.array_equal(p, prob.partition(parts)), "deterministic"


def test_c_abi_exports_every_declared_symbol(pkg):
    """libnsb.so loads without a GPU and exports what include/nsb.h declares;
    the product refuses to run (no CPU fallback) when no CUDA device exists."""
    subprocess.check_call(["make", "-C", ROOT, "cuda"], stdout=subprocess.DEVNULL)
    hdr = open(os.path.join(ROOT, "include", "nsb.h")).read()
    declared = set(re.findall(r"\b(nsb_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(pkg.device.SYMBOLS), declared ^ set(pkg.device.SYMBOLS)
    lib = pkg.device_lib()
    for s in declared:
        assert hasattr(lib, s), s
    hdr_h = open(os.path.join(ROOT, "include", "nsb_host.h")).read()
    hl = pkg.host_lib()
    for s in set(re.findall(r"\b(nsh_[a-z0-9_]+)\s*\(", hdr_h)):
        assert hasattr(hl, s), s
    if lib.nsb_device_count() == 0:
        with pytest.raises(pkg.DeviceError):
            pkg.Device(3)


def test_product_does_not_touch_the_oracle():
    """The oracle is test infrastructure: nothing under the package, include/
    or the Makefile's product targets may reference it."""
    pkgdir = os.path.join(ROOT, "navierstokes-capoferri_cecchettini_untila_b200")
    for base, _, files in os.walk(pkgdir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                txt = open(os.path.join(base, f), errors="ignore").read()
                assert "ns_oracle" not in txt and "oracle/" not in txt, os.path.join(base, f)
    # the experiment scripts under tools/ are not checkers either
    for f in os.listdir(os.path.join(ROOT, "tools")):
        if f.endswith((".py", ".cu")):
            txt = open(os.path.join(ROOT, "tools", f), errors="ignore").read()
            assert "ns_oracle" not in txt and "from oracle" not in txt, f


def _airfoil_contour(prob):
    xyz = np.array(prob.array("xyz")).reshape(-1, 2)
    bf = np.array(prob.array("bfaces")).reshape(-1, 2)
    bid = np.array(prob.array("bids"))
    return xyz[np.unique(bf[bid == 4])], xyz, bid


def test_airfoil_preprocessing_matches_the_reference_script(pkg, tmp_path):
    """mesh/test.py:25-41, 155-168 + tests/2D/test_naca/run_test.sh:7-9: contour points shifted to mid-chord,
    scaled to the chord, turned clockwise by the angle of attack, placed at (0.4, 0.5) in the 2.2 x 1.0 box.  The
    boundary nodes tagged 4 must be exactly those transformed points (restated here the way test.py does it)."""
    import math
    dat = tmp_path / "foil.dat"
    # a small symmetric contour in the mesh/naca.dat layout: name line, TE -> upper -> LE -> lower
    xs = 0.5 * (1 + np.cos(np.linspace(0, math.pi, 12)))
    yt = 0.6 * (0.2969 * np.sqrt(xs) - 0.126 * xs - 0.3516 * xs**2 + 0.2843 * xs**3 - 0.1036 * xs**4)
    pts = [(x, y) for x, y in zip(xs, yt)] + [(x, -y) for x, y in zip(xs[-2:0:-1], yt[-2:0:-1])]
    dat.write_text("TEST FOIL\n" + "\n".join(f"{x:.6f}     {y:.6f}" for x, y in pts) + "\n")
    chord, angle = 0.4, 7.0
    prob = pkg.Problem.generate_airfoil(0.03, dat_path=str(dat), chord=chord, aoa_deg=angle).build(
        inlet=(pkg.INLET_UNIFORM, 1.0, 0.41, 0))
    foil, xyz, bid = _airfoil_contour(prob)
    # test.py: data = (x - 0.5, y); resize(chord); rotate(angle): angle -= ...; cos(-a), sin(-a)
    a = -angle * math.pi / 180.0
    want = []
    for x, y in np.loadtxt(dat, skiprows=1):
        X, Y = (x - 0.5) * chord, y * chord
        want.append((0.4 + math.cos(a) * X - math.sin(a) * Y, 0.5 + math.sin(a) * X + math.cos(a) * Y))
    want = np.array(want)
    # every contour point of the file is a boundary-4 vertex (the O-grid adds 4 corner rays)
    d = np.abs(foil[None, :, :] - want[:, None, :]).sum(axis=2).min(axis=1)
    assert d.max() < 1e-12 and len(foil) in (len(want), len(want) + 4)
    assert set(np.unique(bid)) == {0, 1, 2, 3, 4}
    assert xyz[:, 0].min() == 0.0 and abs(xyz[:, 0].max() - 2.2) < 1e-12 and abs(xyz[:, 1].max() - 1.0) < 1e-12
    # nose down for a positive angle of attack: the leading edge (x = 0 of the file) moves up, the trailing edge down
    le = want[np.argmin(np.loadtxt(dat, skiprows=1)[:, 0])]
    assert le[1] > 0.5
    # the cells are positively oriented and the space builds
    s = prob.sizes()
    assert s["n_cells"] > 500 and s["n_bc"] > 0 and s["n_force_faces"] == len(foil)


def test_make_mesh_airfoil_cli(pkg, tmp_path):
    """`make_mesh airfoil nacaDDDD|file.dat chord angle h out.msh` stands in for test.py + gmsh (run_test.sh:7-9)."""
    import subprocess
    subprocess.check_call(["make", "-C", ROOT, "drivers"], stdout=subprocess.DEVNULL)
    exe = os.path.join(ROOT, "navierstokes-capoferri_cecchettini_untila_b200", "drivers", "make_mesh")
    out = tmp_path / "domain2D.msh"
    subprocess.check_call([exe, "airfoil", "naca2408", "0.4", "10", "0.03", str(out)], stdout=subprocess.DEVNULL)
    prob = pkg.Problem.read_msh(str(out), 2).build(inlet=(pkg.INLET_UNIFORM, 1.0, 0.41, 0))
    ref = pkg.Problem.generate_airfoil(0.03, naca4=2408, chord=0.4, aoa_deg=10.0).build(
        inlet=(pkg.INLET_UNIFORM, 1.0, 0.41, 0))
    assert prob.sizes() == ref.sizes()
    assert np.allclose(np.array(prob.array("xyz")), np.array(ref.array("xyz")), atol=1e-12)
    if os.path.exists("/root/reference/mesh/naca.dat"):  # the reference's own contour file, where it lies
        subprocess.check_call([exe, "airfoil", "/root/reference/mesh/naca.dat", "0.4", "5", "0.03", str(out)],
                              stdout=subprocess.DEVNULL)
        assert pkg.Problem.read_msh(str(out), 2).sizes()["n_cells"] > 500


def test_msh_v22_file_as_gmsh_writes_it(pkg, tmp_path):
    """Format 2.2 in the layout gmsh itself produces (`gmsh -2 -format msh22`, what deal.II's GridIn::read_msh of
    the reference reads, src/NavierStokes.cpp:9-17): $PhysicalNames, 15-node (point) elements that must be skipped,
    line elements with two tags (physical, elementary), node numbers that do not start at the first used node, and
    a clockwise triangle."""
    txt = """$MeshFormat
2.2 0 8
$EndMeshFormat
$PhysicalNames
3
1 0 "inlet"
1 4 "obstacle"
2 10 "fluid"
$EndPhysicalNames
$Nodes
6
1 0 0 0
2 1 0 0
3 1 1 0
4 0 1 0
5 0.5 0.5 0
6 7 7 7
$EndNodes
$Elements
9
1 15 2 0 1 1
2 15 2 0 2 2
3 1 2 0 1 4 1
4 1 2 4 2 1 2
5 1 2 4 2 2 3
6 2 2 10 1 1 2 5
7 2 2 10 1 2 3 5
8 2 2 10 1 3 4 5
9 2 2 10 1 1 5 4
$EndElements
"""
    p = tmp_path / "v22.msh"
    p.write_text(txt)
    m = pkg.Problem.read_msh(str(p), 2)
    s = m.sizes()
    assert (s["n_verts"], s["n_cells"], s["n_bfaces"]) == (5, 4, 3)  # node 6 is unused, points are skipped
    assert sorted(m.array("bids").tolist()) == [0, 4, 4]
    xyz, cells = m.array("xyz").reshape(-1, 2), m.array("cells").reshape(-1, 3)
    area = np.linalg.det(xyz[cells[:, 1:]] - xyz[cells[:, :1]]) / 2
    assert (area > 0).all() and abs(area.sum() - 1.0) < 1e-15  # the clockwise cell (1 5 4) was re-oriented
    m.build(inlet=(pkg.INLET_PARABOLIC, 0.3, 1.0, 0))
    assert m.sizes()["n_nodes"] == 5 + 8 and m.sizes()["n_p"] == 5  # vertices + edges
