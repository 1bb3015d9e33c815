"""The C++ NavierStokes facade (host/NavierStokes.hpp) and its driver mains."""
import glob
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, PKG_NAME

PKG_DIR = os.path.join(ROOT, PKG_NAME)
REF = "/root/reference"


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree only exists in the build container")
@pytest.mark.parametrize("case,dim", [("2D/test_01", 2), ("2D/test_02", 2), ("2D/test_03", 2), ("2D/test_naca", 2),
                                      ("3D/test_01", 3), ("3D/test_02", 3), ("3D/test_03", 3)])
def test_reference_drivers_compile_unmodified_against_facade(pkg, tmp_path, case, dim):
    """Drop-in check: the reference's own driver sources (compiled where they
    lie, never copied) build against this repo's NavierStokes.hpp."""
    subprocess.check_call(["make", "-C", ROOT, "host", "cuda"], stdout=subprocess.DEVNULL)
    src = glob.glob(os.path.join(REF, "tests", case, "src", "*.cpp"))
    assert len(src) == 1
    out = str(tmp_path / "drv")
    cmd = ["/usr/bin/g++", "-O0", "-std=c++17", "-fopenmp", f"-DDIM={dim}", "-DNS_INPUT=", "-I" + os.path.join(PKG_DIR, "host"),
           "-I" + os.path.join(ROOT, "include"), "-o", out, src[0], os.path.join(PKG_DIR, "host", "NavierStokes.cpp"),
           "-L" + PKG_DIR, "-lnsb_host", "-lnsb", "-Wl,-rpath," + PKG_DIR]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]


@pytest.mark.parametrize("prec", ["NSB_PREC_AYOSIDA", "NSB_PREC_IDENTITY"])
def test_facade_builds_with_the_alternative_preconditioners(pkg, tmp_path, prec):
    """The reference selects PreconditionAYosida / PreconditionIdentity by (un)commenting blocks of
    NavierStokes.cpp:352-373; the facade takes -DNS_PRECONDITIONER= instead."""
    subprocess.check_call(["make", "-C", ROOT, "host", "cuda"], stdout=subprocess.DEVNULL)
    out = str(tmp_path / "drv")
    cmd = ["/usr/bin/g++", "-O0", "-std=c++17", "-fopenmp", "-DDIM=2", "-DNS_INPUT=", f"-DNS_PRECONDITIONER={prec}",
           "-I" + os.path.join(PKG_DIR, "host"), "-I" + os.path.join(ROOT, "include"), "-o", out,
           os.path.join(PKG_DIR, "drivers", "d2_test_01.cpp"), os.path.join(PKG_DIR, "host", "NavierStokes.cpp"),
           "-L" + PKG_DIR, "-lnsb_host", "-lnsb", "-Wl,-rpath," + PKG_DIR]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]


def _run_driver(tmp_path, binary, mesh_name, h, T):
    subprocess.check_call(["make", "-C", ROOT, "drivers"], stdout=subprocess.DEVNULL)
    for d in ("build", "output", "cache", "mesh"):
        os.makedirs(tmp_path / d, exist_ok=True)
    msh = str(tmp_path / "mesh" / "domain.msh")
    subprocess.check_call([os.path.join(PKG_DIR, "drivers", "make_mesh"), mesh_name, str(h), msh])
    r = subprocess.run([os.path.join(PKG_DIR, "drivers", binary), "../mesh/domain.msh", str(T)],
                       cwd=str(tmp_path / "build"), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return msh, r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("binary,mesh_name,h,dim,um", [("d2_test_01", "2d-cylinder", 0.05, 2, 0.3),
                                                       ("d3_test_01", "3d-cylinder", 0.1, 3, 0.45)])
def test_driver_time_loop_matches_oracle(pkg, oracle_mod, tmp_path, binary, mesh_name, h, dim, um):
    """ctor -> set_re_number -> setup -> compute_ordered_dofs_indices -> solve
    (reference tests/3D/test_01/src/test_01.cpp:57-61) on a generated .msh;
    forces_vs_time.csv against the oracle stepping the same mesh file."""
    nsteps = 4
    msh, stdout = _run_driver(tmp_path, binary, mesh_name, h, nsteps * 0.01)
    assert "GMRES iterations" in stdout and "Drag coefficient (Cd):" in stdout
    lines = open(tmp_path / "build" / "forces_vs_time.csv").read().strip().split("\n")
    assert lines[0] == "time,deltat,GMRES_iters,time_prec_init,time_sol,Drag,Lift,Cd,Cl"
    rows = np.array([[float(x) for x in l.split(",")] for l in lines[1:]])
    assert rows.shape == (nsteps, 9)
    assert np.allclose(rows[:, 0], 0.01 * np.arange(1, nsteps + 1)) and (rows[:, 2] > 0).all()
    prob = pkg.Problem.read_msh(msh, dim).build(inlet=(0, um, 0.41, 0))
    orc = oracle_mod.Oracle(dim, prob.array("xyz"), prob.array("cells"), prob.array("bfaces"), prob.array("bids"))
    orc.set_inlet(0, um, 0.41, 0)
    orc.set_params(0.01, 1e-3)
    orc.set_re_number(20)
    orc.set_solver(1e-10, 30, 10000, 1e-8)
    t = 0.0
    for n in range(nsteps):
        t += 0.01
        orc.assemble(t)
        orc.solve_time_step()
        f = orc.compute_forces(t)
        # both solvers stop at their own 1e-6 / 1e-10 tolerance: compare to 1e-3 relative
        assert abs(rows[n, 7] - f[2]) < 2e-3 * abs(f[2]), (n, rows[n], f)
    # checkpoint files: N raw doubles (reference NavierStokes.cpp:560-567); step 0 holds the zero initial state
    st0 = np.fromfile(tmp_path / "cache" / "state-ns-0.dat")
    assert st0.size == orc.N and not st0.any()


@pytest.mark.gpu
def test_checkpoint_and_vtu_output(pkg, tmp_path):
    msh, _ = _run_driver(tmp_path, "d2_test_02", "2d-cylinder", 0.06, 0.08)  # dt 0.02, output step 2
    out = sorted(os.listdir(tmp_path / "output"))
    assert "output-stokes_2.0.vtu" in out and "output-stokes_2.pvtu" in out and "output-stokes_4.pvtu" in out
    prob = pkg.Problem.read_msh(msh, 2).build()
    s = prob.sizes()
    N = s["n_u"] + s["n_p"]
    st = np.fromfile(tmp_path / "cache" / "state-ns-4.dat")
    assert st.size == N and np.isfinite(st).all() and np.abs(st).max() > 0.1
    txt = open(tmp_path / "output" / "output-stokes_4.0.vtu").read()
    assert f'NumberOfCells="{s["n_cells"]}"' in txt and 'Name="velocity"' in txt and 'Name="partitioning"' in txt


def _restart_driver(base, binary, *args):
    r = subprocess.run([os.path.join(PKG_DIR, "drivers", binary), "../mesh/domain.msh", *[str(a) for a in args]],
                       cwd=str(base / "build"), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return r.stdout


def _csv_rows(path):
    lines = open(path).read().strip().split("\n")[1:]
    return np.array([[float(x) for x in l.split(",")] for l in lines])


@pytest.mark.gpu
@pytest.mark.parametrize("binary,mesh_name,h,dim,um", [("d2_restart", "2d-cylinder", 0.05, 2, 0.3),
                                                       ("d3_restart", "3d-cylinder", 0.1, 3, 0.45)])
def test_restart_and_post_process(pkg, oracle_mod, tmp_path, binary, mesh_name, h, dim, um):
    """Restart path of the reference (src/NavierStokes.cpp:457-463, 501-568, 787-805) and post_process
    (:808-828, src/postprocess.cpp): n steps in one go; a NEW object continues from the checkpoint of step k and
    must reach the same state and force rows as the uninterrupted run; post_process re-reads the checkpoints
    and reproduces the Cd of the time loop; the on-disk order of a checkpoint is the first-encounter order of
    the dofs over the cells (:571-784 on one rank), checked against the oracle's independent numbering."""
    subprocess.check_call(["make", "-C", ROOT, "drivers"], stdout=subprocess.DEVNULL)
    n, k = 6, 3
    for d in ("build", "output", "cache", "mesh"):
        os.makedirs(tmp_path / d, exist_ok=True)
    msh = str(tmp_path / "mesh" / "domain.msh")
    subprocess.check_call([os.path.join(PKG_DIR, "drivers", "make_mesh"), mesh_name, str(h), msh])
    T = n * 0.01
    _restart_driver(tmp_path, binary, T, "full")
    rows_full = _csv_rows(tmp_path / "build" / "forces_vs_time.csv")
    assert rows_full.shape == (n, 9)
    states_full = {s: np.fromfile(tmp_path / "cache" / f"state-ns-{s}.dat") for s in range(n + 1)}
    # --- the on-disk order: file[perm[i]] = solution[i], perm = first encounter over the cells' dof lists ---
    prob = pkg.Problem.read_msh(msh, dim).build(inlet=(0, um, 0.41, 0))
    orc = oracle_mod.Oracle(dim, prob.array("xyz"), prob.array("cells"), prob.array("bfaces"), prob.array("bids"))
    cd = orc.cell_dofs()
    _, first = np.unique(cd, return_index=True)       # first occurrence of every dof in the cell walk
    perm = np.empty(orc.N, np.int64)
    perm[np.argsort(first)] = np.arange(orc.N)        # dof -> rank of its first occurrence
    dev = pkg.Device(dim).load_problem(prob)
    nu = prob.mean_velocity(0.0) * 0.4 / 20
    dev.set_params(0.01, nu)
    t = 0.0
    for _ in range(k):
        t += 0.01
        dev.assemble(t)
        dev.solve_time_step()
    x = dev.solution()
    expect = np.empty(orc.N)
    expect[perm] = x
    # two runs of the same solver at its 1e-6 stopping rule (atomics order differs): agreement far below the O(1)
    # error of a wrong permutation
    assert np.linalg.norm(states_full[k] - expect) < 1e-4 * np.linalg.norm(expect)
    dev.close()
    # --- post_process over the checkpoints of the full run: same Cd as the time loop printed ---
    out = _restart_driver(tmp_path, binary, T, "post", 1, n)
    assert out.count("Importing time step") == n and out.count("Exporting pvtu files") == n
    cds = [float(l.split("(Cd):")[1].split()[0]) for l in out.split("\n") if "Drag coefficient (Cd):" in l]
    assert len(cds) == n and np.allclose(cds, rows_full[:, 7], rtol=1e-5)
    assert os.path.exists(tmp_path / "output" / f"output-stokes_{n}.pvtu")
    # --- restart: a new object continues from step k ---
    out = _restart_driver(tmp_path, binary, T, "restart", k)
    assert f"Continuing execution from time step {k}" in out
    rows_re = _csv_rows(tmp_path / "build" / "forces_vs_time.csv")
    assert rows_re.shape == (n - k, 9)
    assert np.allclose(rows_re[:, 0], rows_full[k:, 0])
    assert np.allclose(rows_re[:, 5:], rows_full[k:, 5:], rtol=2e-4, atol=1e-7), (rows_re[:, 5:], rows_full[k:, 5:])
    st = np.fromfile(tmp_path / "cache" / f"state-ns-{n}.dat")
    assert np.linalg.norm(st - states_full[n]) < 1e-4 * np.linalg.norm(states_full[n])
    # the imported state is written back unchanged at the restart step (export_data right after import_data)
    assert np.array_equal(np.fromfile(tmp_path / "cache" / f"state-ns-{k}.dat"), states_full[k])


def test_export_data_reports_an_unwritable_checkpoint(pkg, tmp_path):
    """a failed write of ../cache/state-ns-*.dat must not pass silently (source-level check: the
    facade tests the stream and throws)."""
    src = open(os.path.join(PKG_DIR, "host", "NavierStokes.cpp")).read()
    body = src[src.index("void NavierStokes::export_data"):src.index("void NavierStokes::import_data")]
    assert "if (!f) throw" in body
