#!/usr/bin/env python
"""bench.py -- the hot-path benchmark (contract: see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one full re-predict sweep (K0 cell index + K1 prediction) of the
BASELINE.json workload `N=1000000 phi=0.70` (configs[2]) on synthetic
non-overlapping disks.  `value` is particles/s with the state resident in HBM
(L2 flushed between steps, CUDA events on the library's own stream); `e2e` is
the same sweep through the public C ABI call pair edmd_cuda_upload +
edmd_cuda_predict_all with pinned HOST buffers, copies inside the timed region.
The per-frame analysis (psi6 + full-range g(r)) is reported under "analysis".

`--impl reference` times the reference's own CPU implementation of the path
(oracle/_ref = the unmodified reference's addNoise() re-predict loop; falls back
to the oracle port when _ref is absent) on the host cores.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

N_PART = 1_000_000
PHI = 0.70
SEED = 12345
BYTES_PER_PARTICLE = 62          # 40 B in (x,y,vx,vy,rad) + 22 B out
L2_FLUSH_BYTES = 256 << 20       # > 126 MB L2
METRIC = "full re-predict particles/s at N=1M"
UNIT = "particles/s"


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        return json.loads(p.read_text()).get("hbm_gbs", 6650.0), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                    "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)),
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_baseline(cfg, repeats=3):
    """The reference's own re-predict (addNoise loop) on one host core."""
    from oracle.oracle_py import Oracle, Reference
    n = cfg["n"]
    if Reference.available():
        ref = Reference()
        ref.setup(n, cfg["lx"], cfg["ly"], 0.0, cfg["x"], cfg["y"], cfg["vx"], cfg["vy"], cfg["rad"])
        ref.predict_first()
        secs = [ref.repredict()["seconds"] for _ in range(repeats)]
        cal = ref.calendar_only()
        # the frame analysis of the same snapshot with the reference's own functions:
        # computeBOOPCutoff over all N; calculate_pcf (single-threaded, all pairs) on the first
        # n_sub particles -- a bounded sample -- and extrapolated to N(N-1)/2 pairs
        out = {"value": n / min(secs), "unit": UNIT, "cores": 1, "kind": "reference",
               "sample": f"{repeats} full re-predict sweeps (reference addNoise loop incl. calendar "
                         f"remove/insert) of the same N={n} snapshot, best of {repeats}",
               "seconds_per_sweep": min(secs), "calendar_only_seconds": cal}
        try:
            out["psi6_seconds"] = float(ref.boop_cutoff(2.5)["seconds"])
            out["psi6_particles_per_s"] = n / out["psi6_seconds"]
            n_sub = min(n, 30000)
            sec = float(ref.pcf(0.1, min(cfg["lx"], cfg["ly"]) / 2, n_sub)["seconds"])
            rate = n_sub * (n_sub - 1) / 2 / sec
            out["gr_sample"] = f"calculate_pcf on the first {n_sub} particles of the snapshot (same box, dr=0.1, max_r=min(L)/2)"
            out["gr_pairs_per_s"] = rate
            out["gr_full_seconds_extrapolated"] = n * (n - 1) / 2 / rate
        except Exception as e:   # the analysis timings are extras of the baseline
            out["analysis_error"] = str(e)[:200]
        ref.teardown()
        return out
    orc = Oracle()
    secs = []
    for _ in range(repeats):
        t0 = time.perf_counter()
        orc.predict_all(n, cfg["lx"], cfg["ly"], 0.0, cfg["x"], cfg["y"], cfg["vx"], cfg["vy"], cfg["rad"])
        secs.append(time.perf_counter() - t0)
    return {"value": n / min(secs), "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"{repeats} sweeps of the oracle port (no calendar) at N={n}, best",
            "seconds_per_sweep": min(secs)}


def end_to_end_coll_rate(t_short=60.0, t_long=150.0):
    """Third part of BASELINE.json's metric: end-to-end collisions per second of the whole program on
    configs[0] (`-N 2000 --phi 0.7`, the reference CLI defaults: 30 % small disks, growth start).
    Ours = graphical-edmd_b200/host/edmd_host (sequential C host + the CUDA library for the whole-system
    sweeps); reference = the unmodified reference's own main(), called through oracle/_ref.  Both are
    sequential event loops on one host core.  LIKE FOR LIKE: each program runs twice (to t_short and to
    t_long); rate = (collisions(t_long) - collisions(t_short)) / (wall(t_long) - wall(t_short)) -- setup,
    growth phase and process start cancel -- with EXACT collision counters (ours: the host's own;
    reference: its global `ncol` read through the shim after main() returns, not a console value)."""
    import re
    import tempfile
    out = {"config": f"-N 2000 --phi 0.7, runs to t = {t_short:g} and t = {t_long:g} (reference CLI defaults, "
                     f"BASELINE configs[0]); rate = difference of the two runs"}
    host = ROOT / "graphical-edmd_b200" / "host" / "edmd_host"

    def ours(t_end):
        with tempfile.TemporaryDirectory() as d:
            r = subprocess.run([str(host), "-N", "2000", "--phi", "0.7", "-t", str(t_end), "-D", "1000",
                                "-o", "1000", "--quiet", "--outdir", d], capture_output=True, text=True, timeout=300)
        m = re.search(r"(\d+) collisions", r.stdout)
        w = re.search(r"whole run ([\d.]+) s", r.stdout)
        return int(m.group(1)), float(w.group(1))

    try:
        (c1, w1), (c2, w2) = ours(t_short), ours(t_long)
        out["ours"] = {"collisions": c2 - c1, "seconds": w2 - w1, "coll_per_s": (c2 - c1) / (w2 - w1), "cores": 1,
                       "runs": [[c1, w1], [c2, w2]]}
    except Exception as e:   # the host program is optional for the headline number
        out["ours_error"] = str(e)[:200]
    ref_so = ROOT / "oracle" / "_ref" / "libedmd_ref.so"
    if ref_so.exists():
        code = ("import ctypes as C, sys, time\n"
                f"lib = C.CDLL({str(ref_so)!r})\n"
                "a = [b'a.out', b'-N', b'2000', b'--phi', b'0.7', b'-t', sys.argv[1].encode()]\n"
                "argv = (C.c_char_p * (len(a) + 1))(*a, None)\n"
                "t0 = time.time(); lib.edmd_reference_main(len(a), argv); w = time.time() - t0\n"
                "lib.ref_last_ncol.restype = C.c_ulong\n"
                "sys.stdout.flush(); print('\\nNCOL', lib.ref_last_ncol(), 'WALL', w)\n")

        def ref(t_end):
            with tempfile.TemporaryDirectory() as d:
                os.makedirs(os.path.join(d, "dump"), exist_ok=True)
                r = subprocess.run([sys.executable, "-c", code, str(t_end)], cwd=d, capture_output=True,
                                   text=True, timeout=300)
            m = re.search(r"NCOL (\d+) WALL ([\d.]+)", r.stdout)
            return int(m.group(1)), float(m.group(2))

        try:
            (c1, w1), (c2, w2) = ref(t_short), ref(t_long)
            out["reference"] = {"collisions": c2 - c1, "seconds": w2 - w1, "coll_per_s": (c2 - c1) / (w2 - w1),
                                "cores": 1, "runs": [[c1, w1], [c2, w2]]}
        except Exception as e:
            out["reference_error"] = str(e)[:200]
    return out


def thermostat_run_1m():
    """The end-to-end number where the GPU matters: edmd_host at N = 10^6 with the velocity-rescale thermostat
    (every tick = upload + sweep + download + calendar ingest from the device's plan).  The reference's CLI
    build has `noise` compiled out (const int noise = 0), so its side is the per-tick cost of its own addNoise()
    (cpu_baseline.seconds_per_sweep: re-predict incl. calendar remove/insert)."""
    import re
    import tempfile
    host = ROOT / "graphical-edmd_b200" / "host" / "edmd_host"
    try:
        with tempfile.TemporaryDirectory() as d:
            r = subprocess.run([str(host), "-N", "1000000", "--phi", "0.7", "-x", "0", "--init", "lattice", "-t", "0.6",
                                "-D", "1000", "-o", "1000", "--noise", "2", "--dtnoise", "0.1", "--quiet",
                                "--outdir", d], capture_output=True, text=True, timeout=600)
        m = re.search(r"(\d+) collisions, \d+ crossings in ([\d.]+) s => ([\d.e+]+) coll/s.*?(\d+) GPU sweeps, "
                      r"([\d.]+) ms each.*?calendar ingest ([\d.]+) ms each", r.stdout)
        if not m:
            return {"error": (r.stdout + r.stderr)[-300:]}
        return {"config": "edmd_host -N 1000000 --phi 0.7 -x 0 --init lattice --noise 2 --dtnoise 0.1 -t 0.6",
                "collisions": int(m.group(1)), "seconds": float(m.group(2)), "coll_per_s": float(m.group(3)),
                "gpu_sweeps": int(m.group(4)), "ms_per_sweep_upload_k0_k1_download": float(m.group(5)),
                "ms_calendar_ingest_per_tick": float(m.group(6)),
                "ms_per_tick": float(m.group(5)) + float(m.group(6)),
                "note": "per tick: GPU sweep through the C ABI + the host's calendar rebuild from the device's ingest "
                        "plan; compare cpu_baseline.seconds_per_sweep (the reference's addNoise tick, one core)"}
    except Exception as e:
        return {"error": str(e)[:200]}


def run_reference(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle.oracle_py import Oracle, Reference
    n = cfg["n"]
    steps, warm = args.steps, args.warmup
    if Reference.available():
        kind = "reference"
        ref = Reference()
        ref.setup(n, cfg["lx"], cfg["ly"], 0.0, cfg["x"], cfg["y"], cfg["vx"], cfg["vy"], cfg["rad"])
        ref.predict_first()
        for _ in range(warm):
            ref.repredict()
        secs = [ref.repredict()["seconds"] for _ in range(steps)]
        ref.teardown()
    else:
        kind = "port"
        orc = Oracle()
        secs = []
        for it in range(warm + steps):
            t0 = time.perf_counter()
            orc.predict_all(n, cfg["lx"], cfg["ly"], 0.0, cfg["x"], cfg["y"], cfg["vx"], cfg["vy"], cfg["rad"])
            if it >= warm:
                secs.append(time.perf_counter() - t0)
    ms = 1e3 * float(np.mean(secs))
    val = n / (ms * 1e-3)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT,
        "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": workload_config(cfg, args.gpus),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": 1, "kind": kind,
                         "sample": f"{steps} full re-predict sweeps of the N={n} snapshot "
                                   "(single-threaded, as the reference ships)"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def workload_config(cfg, gpus=1):
    which = "BASELINE configs[2]" if gpus == 1 else \
        f"{gpus} x BASELINE configs[2] in one periodic box (configs[3] at 4 GPUs)"
    return {"workload": f"N={cfg['n']} phi={PHI} monodisperse liquid-density jittered-lattice "
                        f"snapshot, {cfg.get('order', 'shuffled')} ids, full re-predict sweep ({which})",
            "n_particles": cfg["n"], "phi": PHI, "seed": SEED,
            "box": [cfg["lx"], cfg["ly"]],
            "l2": f"flushed between steps ({L2_FLUSH_BYTES >> 20} MiB written then {L2_FLUSH_BYTES >> 20} MiB "
                  "read, outside the timed events: cold and clean)"}


def profile_traffic():
    """dram bytes per launch of K1 from the committed ncu summary, if any."""
    p = ROOT / "profiles" / "k1_traffic.json"
    if p.exists():
        try:
            return json.loads(p.read_text()).get("dram_bytes_per_launch")
        except Exception:
            return None
    return None


def bind_to_gpu_numa_node(local):
    """Run this rank's host thread (and so its pinned staging buffers, first-touch) on the CPUs next to
    its GPU (NVML's CPU affinity of the device) when the box exposes more than one NUMA node."""
    try:
        import pynvml
        pynvml.nvmlInit()
        hdl = pynvml.nvmlDeviceGetHandleByIndex(local)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(hdl, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:
        return 0


def timed_sweeps(ctx, B, steps, warm):
    """(ms per step, ms of K1) of the resident sweep: exactly `steps` timed steps of the chain AS SHIPPED (no event
    between K0 and K1: the second kernel's programmatic launch overlaps the first one's tail), then a second run
    with an event between the two kernels for K1's own duration (the roofline figure)."""
    tot, _ = ctx.bench(B.BENCH_SWEEP, warmup=warm, iters=steps, flush_bytes=L2_FLUSH_BYTES, split=False)
    _, main = ctx.bench(B.BENCH_SWEEP, warmup=2, iters=steps, flush_bytes=L2_FLUSH_BYTES)
    return tot, main


def slab_sweep_ms(pkg, dist, torch, cfg, rank, world, local, steps, warm):
    """Device-resident multi-GPU step of one system: ms per step (max over ranks), K1 part, n_owned."""
    slab = pkg.slab
    N, lx, ly = cfg["n"], cfg["lx"], cfg["ly"]
    fx, fy = 1.0 / (lx / int(lx / 2)), 1.0 / (ly / int(ly / 2))     # cellxFac, cellyFac (src/EDMD.c:694-706)
    cells = np.stack([(cfg["x"] * fx).astype(np.int32), (cfg["y"] * fy).astype(np.int32)], 1)
    sr = slab.SlabRank(pkg, N, lx, ly, rank, world, local)
    sr.connect_p2p(dist)   # halo = peer stores over NVLink (csrc/halo.cu)
    gid = sr.load_owned(cfg, cells, 0.0)
    sr.exchange(dist)
    dist.barrier()
    torch.cuda.synchronize()
    tot, main = timed_sweeps(sr.ctx, pkg.binding, steps, warm)
    return sr, gid, cells, tot, main


def run_slabs(args, pkg, rank, world, local):
    """N > 1: weak scaling -- the system is N x 1M disks in one periodic box, cut
    into N row slabs of the cell grid (one per GPU).  A step = halo exchange
    (boundary rows stored straight into the neighbours' memory over NVLink, on a
    second stream beside the partition of the owned particles) + K0 + K1 on every
    rank, all inside the CUDA-event bracket; outputs are disjoint, no collective
    on the data path.  Plus: psi6 per slab with the mean q6 all-reduced, g(r)
    split over the ranks, and the STRONG-scaling point of BASELINE configs[3]
    (N = 4*10^6 fixed, on this many GPUs)."""
    import torch
    import torch.distributed as dist

    B = pkg.binding
    ncpu_bound = bind_to_gpu_numa_node(local)
    n_total = args.n * world
    cfg = pkg.synth.lattice_config(n_total, PHI, SEED, shuffle=(args.order == "shuffled"))
    N, lx, ly = cfg["n"], cfg["lx"], cfg["ly"]
    steps, warm = args.steps, max(args.warmup, 3)

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    with ClockSampler(local) as clk:
        sr, gid, cells, tot, main = slab_sweep_ms(pkg, dist, torch, cfg, rank, world, local, steps, warm)
        n_owned = len(gid)
        halo_bytes = 2 * sr.halo_capacity * 48
        _, n_local = sr.ctx.counts()
        l0 = sr.ctx.launches
        sr.ctx.bench(B.BENCH_SWEEP, warmup=0, iters=1, flush_bytes=0)
        launches_per_step = sr.ctx.launches - l0
        # ---- end to end: pinned host buffers -> upload owned, exchange + sweep (one device sequence),
        # results into pinned host buffers; every call goes through the C ABI
        keep, own = [], {}
        for k in ("x", "y", "vx", "vy", "rad"):
            tpin = torch.from_numpy(np.ascontiguousarray(cfg[k][gid])).pin_memory()
            keep.append(tpin)
            own[k] = tpin.numpy()
        for k, a in (("cells", cells[gid]), ("gid", gid.astype(np.int32))):
            tpin = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
            keep.append(tpin)
            own[k] = tpin.numpy()
        outs = {}
        for k, dt in (("t_cross", torch.float64), ("dir", torch.uint8), ("t_coll", torch.float64), ("partner", torch.int32)):
            tpin = torch.empty(n_owned, dtype=dt).pin_memory()
            keep.append(tpin)
            outs[k] = tpin.numpy()
        ov = np.zeros(2, np.int32)
        lib, h, P = sr.ctx.lib, sr.ctx._h, B._ptr

        def e2e_step():
            # a thermostat tick: positions, velocities and the host's cells (radii and ids are resident)
            rc = lib.edmd_cuda_upload_owned(h, n_owned, P(own["x"]), P(own["y"]), P(own["vx"]), P(own["vy"]),
                                            None, P(own["cells"]), None, 0.0)
            assert rc == 0, lib.edmd_cuda_last_error(h)
            rc = lib.edmd_cuda_exchange_predict_device(h, B.MODE_NORMAL)
            assert rc == 0, lib.edmd_cuda_last_error(h)
            rc = lib.edmd_cuda_fetch_predictions(h, P(outs["t_cross"]), P(outs["dir"]), P(outs["t_coll"]),
                                                 P(outs["partner"]), None, P(ov))
            assert rc == 0, lib.edmd_cuda_last_error(h)

        e2e = []
        for it in range(3 + steps):
            barrier()
            t0 = time.perf_counter()
            e2e_step()
            if it >= 3:
                e2e.append(time.perf_counter() - t0)
        barrier()
        # ---- psi6 of every slab (halo rows of the last exchange supply the cross-boundary neighbours)
        bt, bm = sr.ctx.bench(B.BENCH_BOOP, dr=2.5, warmup=3, iters=10, flush_bytes=L2_FLUSH_BYTES)
        bo = sr.boop(dist, N)
    clocks = clk.summary()
    ms_step, ms_k1, ms_e2e = float(np.mean(tot)), float(np.mean(main)), 1e3 * float(np.mean(e2e))
    ms_psi6 = float(np.mean(bt))
    # BASELINE configs[3] second half: g(r) of the whole N-GPU system, pairs split over the ranks
    ms_gr, gr_pairs = 0.0, 0
    if args.analysis == "full":
        max_r = min(lx, ly) / 2
        counts, ms_gr = sr.pcf(dist, cfg["x"][gid], cfg["y"][gid], N, 0.1, max_r)
        gr_pairs = int(counts.sum())
    sr.close()
    # ---- strong scaling of BASELINE configs[3]: N = 4*10^6 fixed on `world` GPUs
    n_strong = 4 * args.n
    cfg4 = cfg if world == 4 else pkg.synth.lattice_config(n_strong, PHI, SEED, shuffle=(args.order == "shuffled"))
    sr4, gid4, _, tot4, main4 = slab_sweep_ms(pkg, dist, torch, cfg4, rank, world, local, steps, warm)
    sr4.close()
    ms_strong = float(np.mean(tot4))
    tt = torch.tensor([ms_step, ms_k1, ms_e2e, float(n_local), ms_gr, ms_psi6, ms_strong], device="cuda",
                      dtype=torch.float64)
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms_step, ms_k1, ms_e2e, n_local_max, ms_gr, ms_psi6, ms_strong = tt.tolist()
    if rank == 0:
        peak, how = peaks()
        achieved = BYTES_PER_PARTICLE * n_owned / (ms_k1 * 1e-3) / 1e9
        wc = workload_config({**cfg, "order": args.order}, world)
        wc["parallelism"] = (f"{world} row slabs of the cell grid, one-cell-row halo by peer stores over NVLink "
                             f"({halo_bytes} B/rank/step inbox capacity) on a second stream beside the partition, "
                             f"no data-path collective")
        wc["step_breakdown_ms"] = {"halo_exchange_plus_K0": ms_step - ms_k1, "K1": ms_k1}
        line = {
            "metric": METRIC, "value": N / (ms_step * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": steps, "warmup": warm, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": wc,
            "clocks": clocks,
            "e2e": {"value": N / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": 40 * n_owned, "d2h_bytes_per_step": 21 * n_owned,
                    "api": "per rank, pinned host buffers: edmd_cuda_upload_owned(x,y,vx,vy,cell_xy; radii and ids resident) + "
                           "edmd_cuda_exchange_predict_device + edmd_cuda_fetch_predictions(t_cross,dir,t_coll,partner)",
                    "pcie_gb_per_s_per_rank": 61 * n_owned / (ms_e2e * 1e-3) / 1e9,
                    "host_cpus_bound_per_rank": ncpu_bound,
                    "note": "every rank moves 61 B/particle over its own PCIe link at the same time; the per-rank "
                            "rate falls when the ranks share the host's memory system"},
            "gpu_launches": int(launches_per_step) * steps * world,
            "roofline": {"bound": "hbm", "kernel": "k_cell_sweep (K1 of the cell-slot sweep), rank 0", "achieved": achieved,
                         "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "peak_source": how + " (burst copy)", "traffic": None,
                         "ms_kernel": ms_k1, "bytes_per_particle": BYTES_PER_PARTICLE},
            "strong_scaling_4M": {"n_particles": cfg4["n"], "n_gpus": world, "ms_per_step": ms_strong,
                                  "particles_per_s": cfg4["n"] / (ms_strong * 1e-3),
                                  "note": "BASELINE configs[3]: the same N = 4*10^6 phi = 0.70 system on 1/2/4/8 GPUs; "
                                          "efficiency = t(1 GPU) / (n_gpus * t(n_gpus)) across the lines of the scaling run"},
            "analysis": {"psi6_ms": ms_psi6, "psi6_particles_per_s": N / (ms_psi6 * 1e-3),
                         "mean_q6_all_reduced": bo["mean_q6_global"],
                         "psi6_note": "computeBOOPCutoff per slab (partition + tile kernel + unpack, max over ranks); "
                                      "the per-rank q6 sums are all-reduced (NCCL) into the thermo column's mean q6"},
        }
        if args.analysis == "full":
            line["analysis"].update({
                "gr_full_ms": ms_gr, "gr_full_bins": int(min(lx, ly) / 2 / 0.1),
                "gr_full_pairs_per_s": N * (N - 1) / 2 / (ms_gr * 1e-3), "gr_pairs_binned": gr_pairs,
                "note": "calculate_pcf dr=0.1 max_r=min(L)/2 of the whole system: all-gather of the positions "
                        "(NCCL) + sorted-tile kernel on the tile pairs w = rank (mod N) + all-reduce of the "
                        "counts, CUDA events, max over ranks"})
        print(json.dumps(line, default=float))


def bench_config(pkg, local, name, n, phi, sf, steps, warm, with_psi6=True):
    """Sweep (+ psi6) timing of one more BASELINE configuration on one GPU."""
    B = pkg.binding
    c = pkg.synth.lattice_config(n, phi, SEED, small_fraction=sf, shuffle=True)
    peak, _ = peaks()
    with pkg.EdmdCuda(c["n"], c["lx"], c["ly"], device=local) as ctx:
        ctx.upload(c["x"], c["y"], c["vx"], c["vy"], c["rad"], t=0.0)
        r0 = ctx.stat(B.STAT_EXACT_RESCANS)
        tot, main = timed_sweeps(ctx, B, steps, warm)
        rescans = (ctx.stat(B.STAT_EXACT_RESCANS) - r0) / (2 * steps + warm + 2)
        out = {"config": name, "n_particles": c["n"], "phi": phi, "small_fraction": sf,
               "ms_per_step": float(np.mean(tot)), "particles_per_s": c["n"] / (float(np.mean(tot)) * 1e-3),
               "k1_ms": float(np.mean(main)),
               "k1_hbm_frac": BYTES_PER_PARTICLE * c["n"] / (float(np.mean(main)) * 1e-3) / 1e9 / peak,
               "exact_rescans_per_sweep": rescans,
               "tile_sweep": bool(ctx.stat(B.STAT_LEAN_ELIGIBLE) and not ctx.stat(B.STAT_LEAN_DECLINES))}
        if with_psi6:
            bt, bm = ctx.bench(B.BENCH_BOOP, dr=2.5, warmup=2, iters=5, flush_bytes=L2_FLUSH_BYTES)
            out["psi6_ms"] = float(np.mean(bt))
            out["psi6_kernel_ms"] = float(np.mean(bm))
    return out


def run_ours(args, cfg):
    import torch
    import torch.distributed as dist
    import __graft_entry__ as entry

    pkg = entry.load_package()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        run_slabs(args, pkg, rank, world, local)
        dist.barrier()
        dist.destroy_process_group()
        return

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    n = cfg["n"]
    steps, warm = args.steps, max(args.warmup, 3)
    B = pkg.binding
    ctx = pkg.EdmdCuda(n, cfg["lx"], cfg["ly"], device=local)

    # pinned host buffers for the end-to-end leg
    def pinned(a):
        t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        return t, t.numpy()
    keep, host = [], {}
    for k in ("x", "y", "vx", "vy", "rad"):
        t, host[k] = pinned(cfg[k])
        keep.append(t)
    outs = {}
    for k, dt in (("t_cross", np.float64), ("dir", np.uint8), ("t_coll", np.float64),
                  ("partner", np.int32), ("ctype", np.uint8)):
        t = torch.empty(n, dtype=getattr(torch, np.dtype(dt).name)).pin_memory()
        keep.append(t)
        outs[k] = t.numpy()
    ov = np.zeros(2, np.int32)
    lib, h = ctx.lib, ctx._h
    P = B._ptr
    # the host's cell[2] of every particle (the mid-run form of the upload: include/edmd_cuda.h)
    fx, fy = 1.0 / (cfg["lx"] / int(cfg["lx"] / 2)), 1.0 / (cfg["ly"] / int(cfg["ly"] / 2))
    t, host["cells"] = pinned(np.stack([(cfg["x"] * fx).astype(np.int32), (cfg["y"] * fy).astype(np.int32)], 1))
    keep.append(t)
    # calendar geometry of the reference (boxConstantHelper: dtPaul = 5/N, paulListN = N)
    cal = {}
    for k, m in (("bucket", 2 * n), ("next", 2 * n), ("prev", 2 * n), ("head", n + 1)):
        t = torch.empty(m, dtype=torch.int32).pin_memory()
        keep.append(t)
        cal[k] = t.numpy()
    n_tree = C.c_int32(0)

    def e2e_step():
        # a thermostat tick: positions, velocities and the host's cells up (radii are unchanged:
        # rad = NULL), crossing + collision events down (ctype is the constant COLLISION: not fetched)
        rc = lib.edmd_cuda_upload(h, P(host["x"]), P(host["y"]), P(host["vx"]), P(host["vy"]),
                                  None, P(host["cells"]), 0.0)
        assert rc == 0, lib.edmd_cuda_last_error(h)
        rc = lib.edmd_cuda_predict_all(h, B.MODE_NORMAL, None, P(outs["t_cross"]), P(outs["dir"]),
                                       P(outs["t_coll"]), P(outs["partner"]), None, P(ov))
        assert rc == 0, lib.edmd_cuda_last_error(h)

    def plan_step():
        # the calendar ingest plan of those events (edmd_cuda_calendar_plan), downloaded: what the host
        # needs to rebuild its calendar in one streaming pass instead of 2N addEventToQueue calls
        rc = lib.edmd_cuda_calendar_plan(h, C.c_double(0.0), C.c_double(5.0 / n), n, 17, P(cal["bucket"]),
                                         P(cal["next"]), P(cal["prev"]), P(cal["head"]), C.byref(n_tree))
        assert rc == 0, lib.edmd_cuda_last_error(h)

    ctx.upload(host["x"], host["y"], host["vx"], host["vy"], host["rad"], t=0.0)

    # ---- device-resident sweep -------------------------------------------
    barrier()
    l0 = ctx.launches
    with ClockSampler(local) as clk:
        tot, _ = ctx.bench(B.BENCH_SWEEP, warmup=warm, iters=steps, flush_bytes=L2_FLUSH_BYTES, split=False)
        barrier()
        launches_timed = (ctx.launches - l0) * steps // (steps + warm)
        # K1's own duration (the roofline figure): a second run with an event between the two kernels
        _, main = ctx.bench(B.BENCH_SWEEP, warmup=2, iters=steps, flush_bytes=L2_FLUSH_BYTES)
        # ---- end to end through the C ABI, host buffers ------------------
        for _ in range(3):
            e2e_step()
            plan_step()
        barrier()
        e2e_times, plan_times = [], []
        for _ in range(steps):
            t0 = time.perf_counter()
            e2e_step()
            t1 = time.perf_counter()
            plan_step()
            e2e_times.append(t1 - t0)
            plan_times.append(time.perf_counter() - t1)
        barrier()
    clocks = clk.summary()

    ms_step = float(np.mean(tot))
    ms_k1 = float(np.mean(main))
    e2e_ms = float(np.mean(e2e_times)) * 1e3
    if world > 1:
        tt = torch.tensor([ms_step, ms_k1, e2e_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_step, ms_k1, e2e_ms = tt.tolist()

    # ---- per-frame analysis (rank-local; reported, not the headline) -----
    analysis = {}
    if args.analysis != "none" and rank == 0:
        bt, bm = ctx.bench(B.BENCH_BOOP, dr=2.5, warmup=3, iters=10, flush_bytes=L2_FLUSH_BYTES)
        analysis["psi6_ms"] = float(np.mean(bt))
        analysis["psi6_kernel_ms"] = float(np.mean(bm))
        analysis["psi6_particles_per_s"] = float(n / (np.mean(bt) * 1e-3))
        analysis["psi6_hbm_frac"] = 56.0 * n / (np.mean(bt) * 1e-3) / 1e9 / peaks()[0]
        # the same frame right after a sweep: the tile buckets are re-used, no partition, no re-index
        ctx.predict_device()
        ctx.stat(B.STAT_EXACT_RESCANS)
        torch.cuda.synchronize()
        reuse = []
        for _ in range(5):
            t0 = time.perf_counter()
            lib.edmd_cuda_boop_cutoff(h, C.c_double(2.5), None, None, None, None, None, None)
            reuse.append(time.perf_counter() - t0)
        analysis["psi6_after_sweep_wall_ms"] = float(np.median(reuse) * 1e3)
        analysis["psi6_roofline"] = {
            "bound": "fp64 pipe + instruction issue (not HBM: 56 B/particle is 0.06 % of a millisecond of HBM traffic)",
            "fp64_ops_per_particle": "8 branch-free plane-0 candidates x 30 (distance, 1/r by one third-order step, "
                                     "z^2 z^4 by unit-modulus squarings, z^6, sum6 and the A/B sums that give sum5 / sum7) "
                                     "+ 68 (three moduli, 1/n, atan2 by one rotation + degree-8 polynomial) "
                                     "+ 30 per extra disk of the 3 x 3 block",
            "hbm_frac": analysis["psi6_hbm_frac"]}
        max_r_cut = 12.0
        if world == 1:
            # K5: Voronoi psi6 + cell area / perimeter (computeBOOPVoronoi, get_particle_voronoi_area)
            vt, vm = ctx.bench(B.BENCH_VORONOI, warmup=2, iters=5, flush_bytes=L2_FLUSH_BYTES)
            analysis["voronoi_ms"] = float(np.mean(vt))
            analysis["voronoi_cells_kernel_ms"] = float(np.mean(vm))
            analysis["voronoi_particles_per_s"] = float(n / (np.mean(vt) * 1e-3))
            # a whole thermostat tick on the resident state: free flight (dt = 0), kinetic energy,
            # velocity rescale, K0 + K1 -- wall clock around the C-ABI calls, nothing crosses PCIe
            # but the two scalars of the rescale
            torch.cuda.synchronize()
            ticks = []
            for _ in range(5):
                t0 = time.perf_counter()
                ctx.free_fly(0.0)              # the snapshot was uploaded at t = 0: dt = 0
                ctx.rescale_velocities(1.0)
                ctx.predict_device()
                ctx.stat(B.STAT_EXACT_RESCANS)   # synchronises the context's stream
                ticks.append(time.perf_counter() - t0)
            analysis["device_tick_ms"] = float(np.median(ticks) * 1e3)
        if args.analysis == "full":
            max_r = min(cfg["lx"], cfg["ly"]) / 2
            # one untimed frame first: the first call allocates the sort scratch and loads the kernels
            pt, _ = ctx.bench(B.BENCH_PCF, dr=0.1, max_r=max_r, warmup=1, iters=2)
            pairs = n * (n - 1) / 2
            analysis["gr_full_ms"] = float(np.mean(pt))
            analysis["gr_full_bins"] = int(max_r / 0.1)
            analysis["gr_full_pairs_per_s"] = pairs / (float(np.mean(pt)) * 1e-3)
            analysis["gr_full_reference_op_rate_tflops_9op"] = 9 * pairs / (float(np.mean(pt)) * 1e-3) / 1e12
            analysis["frames_per_s_gr_full_plus_psi6"] = 1e3 / (float(np.mean(pt)) + np.mean(bt))
            analysis["gr_roofline"] = {
                "bound": "instruction issue + shared-memory atomics (positions fit L2; bytes negligible)",
                "pairs": pairs, "instructions_per_pair": 10,
                "fp64_equivalent_tflops_9op": 9 * pairs / (float(np.mean(pt)) * 1e-3) / 1e12,
                "fp64_peak_tflops_nominal": 37.0,
                "frac_of_nominal_fp64_peak": 9 * pairs / (float(np.mean(pt)) * 1e-3) / 1e12 / 37.0,
                "note": "SURVEY 8d counts 9 FP64 operations per pair for the reference's arithmetic; the kernel decides "
                        "99.6 % of the pairs in FP32 (certified), so the FP64 pipe itself is ~1 % busy"}
            # the same frame with every bin certified in FP64 (the previous default kernel):
            # its time, and the two histograms compared bin by bin at full size
            e0 = ctx.stat(B.STAT_PCF_EXACT_PAIRS)
            fast = ctx.pcf(0.1, max_r)["counts"]
            analysis["gr_full_exact_path_share"] = (ctx.stat(B.STAT_PCF_EXACT_PAIRS) - e0) / pairs
            ctx.set_option(B.OPT_PCF_LEGACY, 2)
            p64, _ = ctx.bench(B.BENCH_PCF, dr=0.1, max_r=max_r, warmup=1, iters=1)
            slow = ctx.pcf(0.1, max_r)["counts"]
            ctx.set_option(B.OPT_PCF_LEGACY, 0)
            analysis["gr_full_ms_fp64_certified_kernel"] = float(p64[0])
            analysis["gr_full_counts_equal_fp64_certified"] = bool(np.array_equal(fast, slow))
            analysis["gr_full_pairs_in_range"] = int(fast.sum())
        analysis["note"] = ("psi6 = computeBOOPCutoff r_c=2.5 (56 B/particle, HBM); g(r) full range = "
                            "calculate_pcf dr=0.1 max_r=min(L)/2, all N(N-1)/2 pairs; bins decided in FP32 under "
                            "a rigorous error bound, undecided pairs (exact_path_share) redone in FP64")
        del max_r_cut

    # ---- second input family: the reference-grown liquid, tiled ------------
    liquid = None
    fixture = ROOT / "tests" / "golden" / "liquid_n10000_phi070.npz"
    if rank == 0 and world == 1 and args.analysis != "none" and fixture.exists() and n >= 40000:
        k = max(2, int(round((n / 10000) ** 0.5)))
        lc = pkg.synth.tiled_config(np.load(fixture), k, SEED)
        with pkg.EdmdCuda(lc["n"], lc["lx"], lc["ly"], device=local) as lctx:
            lctx.upload(lc["x"], lc["y"], lc["vx"], lc["vy"], lc["rad"], t=0.0)
            eligible = lctx.stat(B.STAT_LEAN_ELIGIBLE)
            ltot, lmain = timed_sweeps(lctx, B, steps, warm)
            declines = lctx.stat(B.STAT_LEAN_DECLINES)
            liquid = {
                "workload": f"liquid grown and equilibrated by the reference itself (N0=10000, phi=0.70, "
                            f"tests/golden/make_liquid.py) tiled {k}x{k}, fresh Maxwell velocities, shuffled ids",
                "n_particles": lc["n"], "distinct_radii": int(len(np.unique(lc["rad"]))),
                "ms_per_step": float(np.mean(ltot)), "particles_per_s": lc["n"] / (float(np.mean(ltot)) * 1e-3),
                "k1_ms": float(np.mean(lmain)), "lean_path": bool(eligible and not declines),
            }
            if args.analysis == "full":
                lp, _ = lctx.bench(B.BENCH_PCF, dr=0.1, max_r=min(lc["lx"], lc["ly"]) / 2, warmup=1, iters=1)
                liquid["gr_full_ms"] = float(lp[0])

    # ---- the other BASELINE configurations + the strong-scaling point of configs[3] on this one GPU ----
    others, strong = [], None
    if rank == 0 and world == 1 and args.analysis != "none":
        ctx.close()
        ctx = None
        for name, nn, phi, sf in (("configs[1]: N=100000 phi=0.72 near liquid-hexatic", 100000, 0.72, 0.0),
                                  ("configs[4]: N=1000000 phi=0.85 dense crystal", n, 0.85, 0.0),
                                  ("reference CLI default mixture (30 % of radius 0.4) at N=1000000 phi=0.70", n, 0.70, 0.3)):
            others.append(bench_config(pkg, local, name, nn, phi, sf, steps, warm))
        st = bench_config(pkg, local, "configs[3] on one GPU", 4 * n, PHI, 0.0, steps, warm, with_psi6=False)
        strong = {"n_particles": st["n_particles"], "n_gpus": 1, "ms_per_step": st["ms_per_step"],
                  "particles_per_s": st["particles_per_s"],
                  "note": "BASELINE configs[3]: the same N = 4*10^6 phi = 0.70 system on 1/2/4/8 GPUs; "
                          "efficiency = t(1 GPU) / (n_gpus * t(n_gpus)) across the lines of the scaling run"}

    cpu = cpu_baseline(cfg) if (rank == 0 and world == 1 and not args.no_cpu) else None

    if rank == 0:
        peak, how = peaks()
        achieved = BYTES_PER_PARTICLE * n / (ms_k1 * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": world * n / (ms_step * 1e-3), "unit": UNIT,
            "n_gpus": world, "steps": steps, "warmup": warm, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(cfg),
            "clocks": clocks,
            "e2e": {"value": world * n / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": 40 * n, "d2h_bytes_per_step": 21 * n,
                    "api": "edmd_cuda_upload(x,y,vx,vy,cell_xy; radii resident) + edmd_cuda_predict_all"
                           "(t_cross,dir,t_coll,partner), pinned host buffers",
                    "pcie_gb_per_s": 61 * n / (e2e_ms * 1e-3) / 1e9,
                    "calendar_plan_ms": float(np.mean(plan_times)) * 1e3,
                    "calendar_plan_d2h_bytes": 4 * (6 * n + n + 1),
                    "ms_per_tick_with_calendar_plan": e2e_ms + float(np.mean(plan_times)) * 1e3,
                    "note": "a tick = upload (40 B/particle incl. the host's cells) + sweep + download (21 B/particle); "
                            "upload and download cannot overlap (the results depend on every particle), so the tick is "
                            "PCIe-serial: pcie_gb_per_s is the link rate it achieves.  calendar_plan_ms = the device's "
                            "ingest plan + its 28 B/particle download, timed next to it (the reference arm's loop "
                            "includes its calendar remove/insert)"},
            "gpu_launches": int(launches_timed),
            "roofline": {"bound": "hbm", "kernel": "K1 = k_cell_sweep (persistent workers: frame by TMA bulk copies, FP32 screening, exact FP64 winner)",
                         "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "peak_source": how + " (burst copy)",
                         "traffic": profile_traffic(),
                         "ms_kernel": ms_k1, "bytes_per_particle": BYTES_PER_PARTICLE},
            "analysis": analysis,
        }
        if world > 1:
            line["config"]["parallelism"] = f"{world} independent replicas (slab path: see DESIGN.md)"
        if liquid:
            line["liquid_input"] = liquid
        if others:
            line["configs"] = others
        if strong:
            line["strong_scaling_4M"] = strong
        if cpu:
            line["cpu_baseline"] = cpu
        if world == 1 and not args.no_cpu:
            line["end_to_end"] = end_to_end_coll_rate()
            line["end_to_end"]["thermostat_run_n1e6"] = thermostat_run_1m()
        print(json.dumps(line, default=float))
    if ctx is not None:
        ctx.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--analysis", choices=["none", "psi6", "full"], default="full")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--n", type=int, default=N_PART)
    ap.add_argument("--order", choices=["shuffled", "lattice"], default="shuffled",
                    help="particle id order: uncorrelated with position (default, like a "
                         "reference-made configuration) or lattice order")
    args = ap.parse_args()

    import __graft_entry__ as entry
    pkg = entry.load_package()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl != "reference" and world > 1:
        run_ours(args, None)
        return
    # the reference arm sweeps the SAME system as our arm: N x 10^6 disks in one box at --gpus N
    n_sys = args.n * (args.gpus if args.impl == "reference" else 1)
    cfg = pkg.synth.lattice_config(n_sys, PHI, SEED, shuffle=(args.order == "shuffled"))
    cfg["order"] = args.order
    if args.impl == "reference":
        run_reference(args, cfg)
    else:
        run_ours(args, cfg)


if __name__ == "__main__":
    main()
