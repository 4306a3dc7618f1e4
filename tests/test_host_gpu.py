"""End-to-end test of the C host (graphical-edmd_b200/host/edmd_host.c) driving
the CUDA library: the first GPU sweep must equal the host's own per-particle
predictors bit for bit (--verify), the thermostat must hold the temperature,
the ovito dump must keep the reference's byte format, and -- chaotic dynamics
can only be compared statistically -- the virial pressure must sit on the
hard-disk equation of state."""
import re
import subprocess
from pathlib import Path

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu
HOST = ROOT / "graphical-edmd_b200" / "host" / "edmd_host"


def run_host(tmp_path, *args):
    if not HOST.exists():
        subprocess.run(["make", "-C", str(ROOT), "host"], check=True, capture_output=True)
    r = subprocess.run([str(HOST), *map(str, args), "--outdir", str(tmp_path)],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    return r.stdout


def test_host_reference_cli_defaults_verify_and_dump(tmp_path):
    # BASELINE configs[0] flavour: -N 2000 --phi 0.7, 30 % small disks, growth start
    # (reference defaults); --verify then checks a GROW-mode sweep
    out = run_host(tmp_path, "-N", 2000, "--phi", 0.7, "-t", 30, "-D", 10, "-o", 10, "--verify",
                   "--noise", 2, "--dtnoise", 0.5, "--boop", "--pcf", "--quiet")
    assert "0 mismatches" in out
    m = re.search(r"(\d+) collisions.*?([\d.e+]+) coll/s.*?(\d+) GPU sweeps.*?E/N = ([\d.]+)", out)
    assert m, out
    ncol, rate, sweeps, e_n = int(m.group(1)), float(m.group(2)), int(m.group(3)), float(m.group(4))
    assert ncol > 100000 and sweeps >= 60
    assert abs(e_n - 1.0) < 1e-9          # velocity-rescale thermostat holds E/N = T
    dump = next(tmp_path.glob("*.dump")).read_text().splitlines()
    assert dump[0] == "ITEM: TIMESTEP"
    assert dump[2] == "ITEM: NUMBER OF ATOMS" and dump[3] == "1980"
    assert dump[8] == "ITEM: ATOMS id type x y vx vy radius m coll q5 q6 q7 argq6 neighbors"
    assert len(dump[9].split()) == 14
    pcf = np.loadtxt(next(tmp_path.glob("*.pcf")))
    last = pcf[pcf[:, 0] == pcf[-1, 0]]
    g_far = last[last[:, 1] > 15][:, 2]
    assert abs(g_far.mean() - 1.0) < 0.05   # g(r) -> 1 at large r
    assert last[last[:, 1] < 0.75][:, 2].max() == 0.0   # nothing inside the smallest contact (2*0.4)


def test_host_pressure_on_hard_disk_equation_of_state(tmp_path):
    """Monodisperse fluid at phi = 0.5: Z = p / (rho T) = 4.17 +- 3 % (Kolafa-Rottner /
    Henderson 4.13); energy is conserved without thermostat."""
    out = run_host(tmp_path, "-N", 4000, "--phi", 0.5, "-x", 0, "-t", 80, "-D", 1000, "-o", 10, "--quiet",
                   "--init", "lattice")
    th = np.loadtxt(next(tmp_path.glob("*.thermo")), skiprows=1)
    th = th[th[:, 0] > 15]                  # drop the lattice melt
    n = 3968                                # lattice rounding of 4000
    lx_ly = np.pi * n / 0.5
    z = th[:, 3].mean() / (n / lx_ly * 1.0)
    assert abs(z - 4.17) / 4.17 < 0.03, z
    assert np.abs(th[:, 2] - 1.0).max() < 1e-9


def test_host_bulk_calendar_ingest_equals_sequential(tmp_path):
    """The calendar rebuilt from the device's ingest plan (edmd_cuda_calendar_plan)
    pops the same events in the same order as 2N sequential insertions: the two
    runs are the same trajectory, byte for byte."""
    outs = {}
    for mode in ("seq", "bulk"):
        d = tmp_path / mode
        d.mkdir()
        out = run_host(d, "-N", 3000, "--phi", 0.6, "-x", 0, "-t", 12, "-D", 4, "-o", 4, "--quiet",
                       "--init", "lattice", "--noise", 2, "--dtnoise", 0.25, "--ingest", mode)
        m = re.search(r"(\d+) collisions, (\d+) crossings.*?(\d+) GPU sweeps.*?\((\d+) from the device plan\)", out)
        assert m, out
        outs[mode] = (int(m.group(1)), int(m.group(2)), int(m.group(3)), int(m.group(4)),
                      next(d.glob("*.dump")).read_bytes(), next(d.glob("*.thermo")).read_bytes())
    assert outs["seq"][3] == 0 and outs["bulk"][3] == outs["bulk"][2] >= 40
    assert outs["seq"][:3] == outs["bulk"][:3]
    assert outs["seq"][4] == outs["bulk"][4]      # identical dumps
    assert outs["seq"][5] == outs["bulk"][5]      # identical thermo records


def test_host_voronoi_area_struc_and_g6_outputs(tmp_path):
    """The reference's other frame-analysis switches (boopThermo == 1, areaThermo,
    strucThermo, pcfg6Thermo; src/EDMD.c:5053-5064, 5620-5641) through the host:
    dump columns in the reference's format, S(q) file with initStructureFactor's
    header, g6 correlation file in save_pcf_g6's layout."""
    run_host(tmp_path, "-N", 2500, "--phi", 0.72, "-x", 0, "-t", 6, "-D", 3, "-o", 3, "--quiet", "--init", "lattice",
             "--boop-voronoi", "--area", "--struc", 1, "--qmax", 0.5, "--pcfg6")
    dump = next(tmp_path.glob("*.dump")).read_text().splitlines()
    assert dump[8] == "ITEM: ATOMS id type x y vx vy radius m coll q5 q6 q7 argq6 neighbors packingFraction"
    n = int(dump[3])
    rows = np.array([[float(v) for v in line.split()] for line in dump[9:9 + n]])
    assert rows.shape == (n, 15)
    assert rows[:, 13].sum() == 6 * n                 # Voronoi neighbours: mean coordination exactly 6
    lx, ly = float(dump[5].split()[1]), float(dump[6].split()[1])
    # sum_i pi r_i^2 / (local packing fraction) = the box area (columns are %lf: 1e-6 each)
    assert abs((np.pi * rows[:, 6] ** 2 / rows[:, 14]).sum() - lx * ly) < 1e-3 * lx * ly
    struc = next(tmp_path.glob("*.struc")).read_text().splitlines()
    qx = np.array(struc[1].split(), float)
    qy = np.array(struc[2].split(), float)
    assert abs(qx[1] - qx[0] - 2 * np.pi / lx) < 1e-5 and abs(qy[1] - qy[0] - 2 * np.pi / ly) < 1e-5
    frames = (len(struc) - 3) // len(qx)
    assert frames >= 2 and (len(struc) - 3) % len(qx) == 0
    s0 = np.array([line.split() for line in struc[3:3 + len(qx)]], float)
    assert s0.shape == (len(qx), len(qy))
    assert s0[(len(qx) - 1) // 2, (len(qy) - 1) // 2] == n      # S(q = 0) = N
    g6 = next(tmp_path.glob("*.pcfg6")).read_text().splitlines()
    r = np.array(g6[0].split(), float)
    assert r[0] == 1.0 and len(g6) == 1 + frames
    c = np.array(g6[1].split(), float)
    assert len(c) == len(r) and np.abs(c).max() <= 1.0


def test_host_langevin_thermostat_reaches_the_bath_temperature(tmp_path):
    """--noise 1: the Langevin kick of addNoise (randomGaussian, src/EDMD.c:5802-5826) done on the
    device (edmd_cuda_langevin_kick) every dtnoise; the velocities come back to the host's event
    loop.  A fluid at the bath temperature T = 0.4 stays there under the kicks (no spurious heating)."""
    out = run_host(tmp_path, "-N", 3000, "--phi", 0.5, "-x", 0, "-t", 40, "-D", 1000, "-o", 2, "--quiet",
                   "--init", "lattice", "--noise", 1, "--gamma", 0.5, "--dtnoise", 0.25, "-T", 0.4)
    th = np.loadtxt(next(tmp_path.glob("*.thermo")), skiprows=1)
    late = th[th[:, 0] > 15][:, 2]                   # E/N column; exp(-2 gamma t) is gone by t = 15
    assert len(late) >= 10
    assert abs(late.mean() - 0.4) < 0.03, late.mean()   # 1/sqrt(N) = 1.8 % per sample
    m = re.search(r"(\d+) collisions", out)
    assert m and int(m.group(1)) > 50000


def test_host_writes_the_reference_thermo_columns_and_file_names(tmp_path):
    """saveThermo's record (src/EDMD.c:1229-1285 header, :5407-5577 line) and customName's file names
    (:6015-6116): `t Ncol E p px py pxy pyx q6 a2`, the pressure tensor's trace consistent with p,
    the mean q6 column from the device, versioned names that do not overwrite an earlier run."""
    args = ("-N", 1500, "--phi", 0.6, "-x", 0, "-t", 20, "-D", 1000, "-o", 2, "--quiet", "--init", "lattice", "--boop")
    run_host(tmp_path, *args)
    run_host(tmp_path, *args)
    names = sorted(p.name for p in tmp_path.glob("*.thermo"))
    assert len(names) == 2 and names[0].endswith("v_0.thermo") and names[1].endswith("v_1.thermo"), names
    assert re.match(r"N_1482res_1\.000phi_0\.600000q_0\.000rat_0\.400Lx_\d+\.\d{3}Ly_\d+\.\d{3}v_0\.thermo", names[0]), names[0]
    lines = (tmp_path / names[0]).read_text().splitlines()
    assert lines[0] == "t Ncol E p px py pxy pyx q6 a2 "
    th = np.loadtxt(tmp_path / names[0], skiprows=1)
    assert th.shape[1] == 10
    t, ncol, e, p, pxx, pyy, pxy, pyx, q6, a2 = th.T
    assert (np.diff(ncol) > 0).all() and np.abs(e - 1.0).max() < 1e-6     # `%lf`: six decimals
    assert np.abs(0.5 * (pxx + pyy) - p).max() < 2e-6                      # p = (px + py)/2 (:5324-5328)
    assert np.abs(pxy - pyx).max() < 1e-9 and np.abs(pxy).max() < 0.2 * p.mean()
    assert ((q6 > 0.2) & (q6 < 1.0)).all()                                 # a phi = 0.6 fluid: local sixfold order ~0.4-0.5
    assert np.abs(a2).max() < 0.2                                          # Maxwellian: a2 ~ 0 +- 1/sqrt(N)
    # two identical runs are the same trajectory
    assert (tmp_path / names[0]).read_bytes() == (tmp_path / names[1]).read_bytes()


def test_host_growth_stop_normalises_on_the_device(tmp_path):
    """stopGrow's normalizePhysicalQ runs on the device (edmd_cuda_normalize_velocities) with the
    reference's --initial-energy: after the growth phase E/N = Einit and stays (no thermostat)."""
    out = run_host(tmp_path, "-N", 1500, "--phi", 0.6, "-x", 0, "-t", 12, "-D", 1000, "-o", 2, "--quiet", "-E", 1.7)
    th = np.loadtxt(next(tmp_path.glob("*.thermo")), skiprows=1)
    assert len(th) >= 4 and np.abs(th[:, 2] - 1.7).max() < 2e-6
    assert re.search(r"E/N = 1\.70000", out), out


def test_host_on_slabs_is_the_same_trajectory(tmp_path):
    """--gpus / --slabs: the re-predict sweeps and the thermo's mean q6 through edmd_cuda_create_mg (here
    three slabs dealt to the GPUs present) -- the same events bit for bit, hence the same run byte for byte."""
    import torch
    outs = {}
    for tag, extra in (("one", ()), ("slabs", ("--gpus", min(2, torch.cuda.device_count()), "--slabs", 3))):
        d = tmp_path / tag
        d.mkdir()
        out = run_host(d, "-N", 6000, "--phi", 0.6, "-x", 0, "-t", 12, "-D", 4, "-o", 2, "--quiet", "--init", "lattice",
                       "--noise", 2, "--dtnoise", 0.5, "--boop", "--ingest", "seq", *extra)
        outs[tag] = (next(d.glob("*.dump")).read_bytes(), next(d.glob("*.thermo")).read_bytes(),
                     re.search(r"(\d+) collisions", out).group(1))
    assert outs["one"][2] == outs["slabs"][2] and int(outs["one"][2]) > 20000
    assert outs["one"][1] == outs["slabs"][1]
    assert outs["one"][0] == outs["slabs"][0]
