#!/usr/bin/env python
"""bench.py — particle-steps/s of the full 3D WCSPH step (BASELINE.json's metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload block|droplet]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A "step" is one Integrator::integrate (reference src/Integration.cpp:233-303): find_timestep, 2 neighbour
builds, 2 x (prestep, surface detection, dissipation, shifting), 1 + k force evaluations with Newmark-Beta
updates and the on-device residual, pn = pnp1.  The workload is config C5 of SURVEY.md 8d (BASELINE.json
configs[4]): a synthetic 500 x 250 x 100 FREE-particle block per GPU (12.5 M particles, dx = 1 mm, jitter
+-0.1 dx, smooth density / velocity fields), slab-decomposed along x for N > 1.  k is pinned to 4
sub-iterations (max_subits = 3, min_residual = -30), so one step holds 5 force evaluations.

One JSON line on stdout (rank 0).  `value` times K steps on device-resident state with CUDA events on the
engine's stream; `e2e` times the same K steps through fjsph_step_host with pinned HOST buffers (upload of
x, v, acc, rho, Rrho, p, m, b and download of x, v, acc, rho, Rrho, p inside the timed region).  `roofline`
is the force kernel (get_acc_and_Rrho), `cpu_baseline` the reference's CPU path on this box's host cores: FJSPH's own
time-step sources compiled unmodified with the reference's flags (makefile:16) against stand-in Eigen / nanoflann
headers (oracle/_ref/liborc_ref3d_fast.so, built where /root/reference exists by oracle/Makefile.ref; kind
"reference"), or, where that library is absent, the CPU restatement under oracle/ built the same way (kind "port").
`--impl reference` times that arm alone.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "particle-steps/sec (full 3D WCSPH step incl. neighbour build)"
UNIT = "particle-steps/s"
K_SUBITS = 4
# SURVEY.md 8d, algorithmic (compulsory) HBM bytes and FP64 flops of one force evaluation
FORCE_BYTES_PER_PARTICLE = 292.0
FORCE_FLOP_PER_PAIR = 150.0
STEP_BYTES_PER_PARTICLE = 2120.0 + 444.0 * (1 + K_SUBITS)
FP64_NOMINAL_TFLOPS = 148 * 64 * 2 * 1.965e9 / 1e12  # 148 SMs x 64 DFMA/clk x 2 flop x clocks.max.sm


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="block", choices=["block", "jet", "droplet"])
    ap.add_argument("--jet-columns", type=int, default=255, help="jet workload: lattice columns along x per GPU (R = 125 dx)")
    ap.add_argument("--cells", default="500,250,100", help="block lattice per GPU (x,y,z)")
    ap.add_argument("--droplet-dx", type=float, default=0.0008)
    ap.add_argument("--solver", default="newmark_beta", choices=["newmark_beta", "runge_kutta"])
    ap.add_argument("--cpu-sample", default="64,64,64", help="lattice of the bounded CPU-baseline sample (block workload)")
    ap.add_argument("--no-check", action="store_true",
                    help="N > 1: skip the slab parity check (2 steps of the slab engines on a ~1 M-particle block against one "
                         "engine, reported as `slab_parity`)")
    ap.add_argument("--no-extras", action="store_true",
                    help="N > 1: skip the side lines (`workloads.jet`: the 12.5 M-per-GPU jet; `strong`: 12.5 M particles in total)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-cpu-port", action="store_true", help="--impl reference: skip the `cpu_port` side line")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def step_params(args, base):
    p = dict(base)
    p.update(delta_t_min=1e-9, frame_time_interval=1e9)
    if args.solver == "newmark_beta":
        p.update(solver_type=0, max_subits=K_SUBITS - 1, min_residual=-30.0)
    else:
        p.update(solver_type=1)
    return p


def make_case(args, rank=0, cells=None):
    from fjsph_b200 import cases

    if args.workload == "droplet":
        return cases.droplet(dx=args.droplet_dx)
    if args.workload == "jet":
        nx = int(cells.split(",")[0]) if cells else args.jet_columns
        return cases.synthetic_jet(nx=nx, dx=1e-3, jitter=0.1, seed=1234, x_offset_cells=rank * nx)
    n = tuple(int(k) for k in (cells or args.cells).split(","))
    return cases.synthetic_block(n, 1e-3, jitter=0.1, seed=1234, x_offset_cells=rank * n[0])


def workload_name(args, world):
    if args.workload == "droplet":
        return "Examples/Droplet 3D (para3D), dx=%g" % args.droplet_dx
    if args.workload == "jet":
        return ("synthetic 3D jet C5: cylinder R=125dx along x, %d columns = ~%.2fM FREE particles per GPU, dx=1e-3, "
                "jitter 0.1dx, v=(30,0,0), Gissler aero in a (0,100,0) cross flow%s" % (
                    args.jet_columns, 49080 * args.jet_columns / 1e6, ", slab-decomposed along x" if world > 1 else ""))
    n = [int(k) for k in args.cells.split(",")]
    return "synthetic 3D block C5, %dx%dx%d = %.2fM FREE particles per GPU, dx=1e-3, jitter 0.1dx%s" % (
        n[0], n[1], n[2], n[0] * n[1] * n[2] / 1e6, ", slab-decomposed along x" if world > 1 else "")


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "200"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(",") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[0]))
                smax.append(float(r[1]))
                power.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for k, nm in enumerate(names):
                if len(r) > 3 + k and r[3 + k].strip().lower().startswith("active"):
                    reasons.add(nm)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(smax)), power_w_max=float(max(power)),
                       reasons=sorted(reasons), samples=len(sm))
        return out


# ----------------------------------------------------------------------------- CPU reference arm
class quiet_fd1:
    """FJSPH prints its step table with printf (Integration.cpp:250-265); bench.py prints ONE JSON line.  Routes the C
    level stdout to /dev/null while the reference's own code runs."""

    def __enter__(self):
        import ctypes

        self.libc = ctypes.CDLL(None)
        sys.stdout.flush()
        self.libc.fflush(None)
        self.saved = os.dup(1)
        devnull = os.open(os.devnull, os.O_WRONLY)
        os.dup2(devnull, 1)
        os.close(devnull)

    def __exit__(self, *exc):
        self.libc.fflush(None)
        os.dup2(self.saved, 1)
        os.close(self.saved)


def cpu_sample_case(args, sample_cells):
    """The bounded sample of the workload the CPU arm steps: same generator, same spacing, jitter and fields, fewer
    particles (the reference's list layout needs ~4.2 KB per particle, and a step of 12.5 M particles takes minutes)."""
    from fjsph_b200 import cases

    if args.workload == "block":
        return make_case(args, 0, cells=sample_cells), "%s lattice of the block workload" % sample_cells.replace(",", "x")
    if args.workload == "jet":
        return (cases.synthetic_jet(nx=64, radius_cells=32, dx=1e-3, jitter=0.1, seed=1234),
                "64-column, R=32dx cylinder of the jet workload")
    return make_case(args), "the droplet itself"


def cpu_reference(args, steps, warmup, sample_cells, kind_pref="reference"):
    """Times the reference's CPU implementation of the step on a bounded sample of the same workload, with every host
    thread: exactly `warmup` untimed steps, then `steps` timed ones; the value is particles / MEDIAN step time.
    kind_pref "reference": oracle/_ref/liborc_ref3d_fast.so -- FJSPH's OWN sources compiled with the reference's flags
    (oracle/Makefile.ref), dissipation_terms serial as the reference has it (Shifting.cpp:126-186 carries no OpenMP
    pragma) -- falling back to the port where that library is absent; "port": the CPU restatement built the same way,
    every loop threaded.  Returns a dict (value, cores, kind, sample, ms_per_step, n, step_ms)."""
    from oracle import oracle as orc  # bench.py's CPU-baseline leg: the checker timed as the baseline

    use_ref = kind_pref == "reference" and orc.have_ref("ref3d_fast")
    if not use_ref:
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "lib/liborc3d_fast.so"])
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    os.environ["OMP_NUM_THREADS"] = str(cores)  # torchrun exports OMP_NUM_THREADS=1; the baseline gets every host core
    try:
        import ctypes

        ctypes.CDLL("libgomp.so.1").omp_set_num_threads(cores)  # in case libgomp was initialised before
    except OSError:
        pass
    case, what_sample = cpu_sample_case(args, sample_cells)
    params = step_params(args, case["params"])
    kind = "ref3d_fast" if use_ref else "3d_fast"
    o = orc.Oracle(orc.default_params(3, **params), kind=kind)
    o.set_particles(case["xi"], case["v"], case["rho"], case["p"], case["m"], case["b"], case["bound_points"])
    n = case["xi"].shape[0] - case["bound_points"]
    its = 0
    step_s = []
    with quiet_fd1():
        for _ in range(warmup):
            o.integrate()
        for _ in range(steps):
            t0 = time.perf_counter()
            _, st = o.integrate()
            step_s.append(time.perf_counter() - t0)
            its = st.iterations
    med = float(np.median(step_s))
    what = ("FJSPH's own sources (Neighbours, Shifting, Resid, Geometry, Containment, Newmark_Beta, Integration .cpp compiled "
            "unmodified, -O3 -ffast-math -funroll-loops -fopenmp -march=x86-64-v3; stand-in Eigen / nanoflann headers, a "
            "uniform-grid radius search in place of the KD-tree; dissipation_terms serial as in the reference)" if use_ref else
            "oracle/fjsph_oracle.cpp, the CPU restatement, every loop threaded (-O3 -ffast-math -funroll-loops -fopenmp "
            "-march=native)")
    desc = "%s, %d threads, median of %d timed steps after %d warm-up on a %s, %d particles, %d sub-iterations" % (
        what, cores, steps, warmup, what_sample, n, its)
    return {"value": n / med, "unit": UNIT, "cores": cores, "kind": "reference" if use_ref else "port", "sample": desc,
            "ms_per_step": med * 1e3, "n": int(n), "what_sample": what_sample,
            "step_ms_min_max": [float(min(step_s)) * 1e3, float(max(step_s)) * 1e3]}


def cpu_baseline_leg(args):
    """The `cpu_baseline` leg of the GPU arm: the reference arm itself (`--impl reference`, 2 warm-up + 4 timed steps, ~30 s)
    in a FRESH process, so that the CPU code runs under the conditions of the reference arm whatever this process holds (a
    CUDA context, pinned buffers), and it is run FIRST (main).  The value has two regimes on these boxes -- 58-61 k
    particle-steps/s on a quiet host, 35-38 k right after other large processes (the GPU arm's own allocations, a test
    suite that has just exited): BASELINE.md 3."""
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "4", "--warmup", "2",
           "--workload", args.workload, "--cpu-sample", args.cpu_sample, "--solver", args.solver, "--no-cpu-port"]
    env = dict(os.environ)
    env.pop("OMP_NUM_THREADS", None)
    try:
        out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=900, env=env)
        line = json.loads([ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1])
        cpu = dict(line["cpu_baseline"])
        cpu["step_ms_min_max"] = line.get("step_ms_min_max")
        cpu["how"] = "bench.py --impl reference --steps 4 --warmup 2 in a fresh process"
        return cpu
    except (subprocess.SubprocessError, ValueError, IndexError, KeyError) as ex:
        print("cpu_baseline: the fresh-process run failed (%s); timing in this process instead" % ex, file=sys.stderr)
        r = cpu_reference(args, 4, 1, args.cpu_sample)
        return {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}


def run_reference(args, rank, world):
    """`--impl reference`: the reference's own CPU implementation of the step, alone.  It runs what it says: `warmup`
    untimed and `steps` timed steps of a bounded SAMPLE of the workload (named in config.workload with its particle
    count -- the 12.5 M-particle configuration itself would need ~50 GB of the reference's list and minutes per step)."""
    if rank != 0:
        return
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    r = cpu_reference(args, steps, warmup, args.cpu_sample)
    port = None
    if r["kind"] == "reference" and not args.no_cpu_port:
        # BASELINE.md 3, line (a): the restatement with every loop threaded, on the same sample (a short run: it is the
        # side line; the headline is the reference's own code above)
        p = cpu_reference(args, 3, 1, args.cpu_sample, kind_pref="port")
        port = {k: p[k] for k in ("value", "unit", "cores", "kind", "sample")}
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "%s -- CPU arm: a bounded sample of it, %s, %d particles" % (
                       workload_name(args, world), r["what_sample"], r["n"]),
                   "particles_total": r["n"], "solver": args.solver,
                   "force_evals_per_step": 1 + K_SUBITS if args.solver == "newmark_beta" else 4,
                   "same_config": False,
                   "note": "the reference's CPU path on host cores; every step is one step of the sample named in workload; "
                           "value = particles / median step time; step_ms_min_max shows the spread"},
        "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "cpu_port": port, "step_ms_min_max": r["step_ms_min_max"],
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ----------------------------------------------------------------------------- our arm
_RESULT_FD = None


def slab_parity(rank, world, local_rank):
    """2 steps of the slab engines (x-slabs over all ranks, ghosts and migration over NCCL) against ONE engine on rank 0, on a
    ~1 M-particle block of the bench workload: every particle compared by id (tools/slab_check.py).  Carries the multi-GPU
    parity into the scaling run itself."""
    import io
    from contextlib import redirect_stdout

    from fjsph_b200 import cases
    from tools import slab_check

    case = cases.synthetic_block((208, 70, 70), 1e-3, jitter=0.1, seed=77)
    buf = io.StringIO()
    with redirect_stdout(buf):
        ok = slab_check.run("bench block", case, dict(case["params"], delta_t_min=1e-9, max_subits=K_SUBITS - 1,
                                                      min_residual=-30.0), 2, rank, world, local_rank)
    errs = {}
    for line in buf.getvalue().splitlines():
        t = line.split()
        if len(t) >= 3 and t[1] == "relerr":
            errs[t[0]] = float(t[2])
        elif len(t) >= 4 and t[1] == "differing":
            errs[t[0] + "_flags_differing"] = int(t[3])
    return {"ok": bool(ok), "particles": int(case["xi"].shape[0]), "steps": 2, "world": world,
            "what": "slab engines vs one engine on rank 0, particle by particle (normwise relative errors; flags: count)",
            "errors": errs}


def time_slab_workload(args, rank, world, local_rank, workload, cells, steps, warmup):
    """A side measurement at N > 1: K steps of another workload / decomposition, device-resident, max over ranks."""
    import argparse

    import torch
    import torch.distributed as dist

    from fjsph_b200 import engine as eng, slab

    a = argparse.Namespace(**vars(args))
    a.workload = workload
    if workload == "jet":
        a.jet_columns = int(cells)
        nx = a.jet_columns
        case = make_case(a, rank, cells=str(nx))
    else:
        a.cells = cells
        nx = int(cells.split(",")[0])
        case = make_case(a, rank)
    params = step_params(a, case["params"])
    n = case["xi"].shape[0]
    dx = case["params"]["particle_step"]
    x_lo = -1e300 if rank == 0 else (rank * nx - 0.5) * dx
    x_hi = 1e300 if rank == world - 1 else ((rank + 1) * nx - 0.5) * dx
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        e = slab.SlabEngine(eng.default_params(3, **params), case, rank, world, x_lo, x_hi, device=local_rank, stream=stream,
                            part_id=np.arange(n, dtype=np.int64) + rank * n)
        for _ in range(warmup):
            e.integrate()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dist.barrier()
        torch.cuda.synchronize()
        ev0.record(stream)
        for _ in range(steps):
            e.integrate()
        ev1.record(stream)
        dist.barrier()
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1)
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    cnt = torch.tensor([float(n - case["bound_points"])], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    e.close()
    del e
    torch.cuda.empty_cache()
    total = float(cnt.item())
    return {"value": total * steps / (float(t.item()) * 1e-3), "unit": UNIT, "ms_per_step": float(t.item()) / steps,
            "particles_total": int(total), "steps": steps, "warmup": warmup, "workload": workload_name(a, world)}


def emit(line: dict) -> None:
    """The ONE JSON line of the contract, on the process's real stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


def main():
    global _RESULT_FD
    args = parse_args()
    # Libraries write to the C-level stdout behind Python's back (NCCL prints "NCCL version ..." when the first
    # communicator comes up, FJSPH prints its step table): everything but the result line goes to stderr.
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    # The cpu_baseline leg first, while this process holds nothing: run after the GPU legs -- beside ~16 GB of host
    # arrays, 2.5 GB of page-locked buffers and a CUDA context -- the same fresh-process run measured 35-38 k particle-steps/s
    # three times on three boxes against 58-61 k as the reference arm.
    cpu_early = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_early = cpu_baseline_leg(args)

    import torch
    import torch.distributed as dist

    from fjsph_b200 import engine as eng

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    case = make_case(args, rank)
    params = step_params(args, case["params"])
    n = case["xi"].shape[0]
    n_fluid = n - case["bound_points"]
    stream = torch.cuda.Stream()
    if world > 1:
        from fjsph_b200 import slab

        # rank r owns lattice columns [r*nx, (r+1)*nx): faces half a spacing outside its first / last column
        nx = args.jet_columns if args.workload == "jet" else int(args.cells.split(",")[0])
        dx = case["params"]["particle_step"]
        x_lo = -1e300 if rank == 0 else (rank * nx - 0.5) * dx
        x_hi = 1e300 if rank == world - 1 else ((rank + 1) * nx - 0.5) * dx
        with torch.cuda.stream(stream):
            e = slab.SlabEngine(eng.default_params(3, **params), case, rank, world, x_lo, x_hi, device=local_rank,
                                stream=stream, part_id=np.arange(n, dtype=np.int64) + rank * n)
    else:
        e = eng.Engine(eng.default_params(3, **params), n, device=local_rank)
        e.set_stream(stream.cuda_stream)
        e.upload_state(case["xi"], case["v"], case["rho"], case["p"], case["m"], case["b"], case["bound_points"])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.cuda.stream(stream):
        for _ in range(args.warmup):
            stats = e.integrate()
        # ---- timed region: K steps on device-resident state
        e.timers_reset()
        e.timers_enable(True)
        sampler = ClockSampler(local_rank)
        launches0 = e.launch_count
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record(stream)
        force_evals = 0
        for _ in range(args.steps):
            stats = e.integrate()
            force_evals += stats.force_evals
        ev1.record(stream)
        barrier()
        ms = ev0.elapsed_time(ev1)
        clocks = sampler.stop()
        launches = e.launch_count - launches0
        timers = e.timers()
        e.timers_enable(False)
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    cnt = torch.tensor([float(n_fluid), float(launches)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    ms = float(t.item())
    total_fluid = float(cnt[0].item())
    value = total_fluid * args.steps / (ms * 1e-3)

    # ---- roofline of the dominant kernel (force evaluation), measured over the timed region
    pairs = float(np.sum(e.neighbour_counts() - 1)) if world == 1 else float(e.pair_count())
    fk = timers.get("force", {"ms": 0.0, "calls": 0})
    force_ms = fk["ms"] / max(1, fk["calls"])
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = FORCE_BYTES_PER_PARTICLE * n_fluid / (force_ms * 1e-3) / 1e9 if force_ms > 0 else 0.0
    # Hardware counters of the same kernels from one `ncu --set full` capture of this build (profiles/kernel_counters.json,
    # made by tools/kernel_counters.py): DRAM bytes and FP64-pipe busy time per launch, scaled by the particle count.  They
    # are profile data, NOT measured in this run (ncu replays every kernel ~40 times); the durations they are set against
    # are this run's CUDA-event times.
    counters = {}
    try:
        counters = json.load(open(os.path.join(ROOT, "profiles", "kernel_counters.json")))
    except (OSError, ValueError):
        pass
    kc = counters.get("kernels", {})
    kscale = n_fluid / float(counters.get("particles", n_fluid) or n_fluid)

    def family_counters(prefixes):
        """Sum over the kernels of one timed family (one launch of each), e.g. the lean + near launches of the fused sweep."""
        picked = [v for k, v in kc.items() if any(k.startswith(p_) for p_ in prefixes)]
        if not picked:
            return None
        return {"dram_bytes": sum(v["dram_bytes"] for v in picked) * kscale,
                "fp64_pipe_busy_ms": sum(v["fp64_pipe_busy_ms"] * v["sm_mhz"] for v in picked) * kscale,  # ms x MHz
                "kernels": [k for k in kc if any(k.startswith(p_) for p_ in prefixes)]}

    def mean_variants(prefix):
        """Mean over the captured instantiations of one kernel (force: the FROZEN and the plain one)."""
        picked = [v for k, v in kc.items() if k.startswith(prefix)]
        if not picked:
            return None
        return {"dram_bytes": float(np.mean([v["dram_bytes"] for v in picked])) * kscale,
                "fp64_pipe_busy_ms": float(np.mean([v["fp64_pipe_busy_ms"] * v["sm_mhz"] for v in picked])) * kscale,
                "kernels": [k for k in kc if k.startswith(prefix)]}

    fcnt = mean_variants("k_force")
    traffic = fcnt["dram_bytes"] if fcnt else None
    fp64_peak, fp64_src = FP64_NOMINAL_TFLOPS, "nominal (148 SM x 64 DFMA/clk x 1.965 GHz)"
    try:
        fp64_peak = float(json.load(open(os.path.join(ROOT, "profiles", "fp64_peak.json")))["fp64_tflops"])
        fp64_src = "measured (profiles/fp64_peak.json, tools/fp64_peak.cu)"
    except (OSError, ValueError, KeyError):
        pass
    fp64_achieved = FORCE_FLOP_PER_PAIR * pairs / (force_ms * 1e-3) / 1e12 if force_ms > 0 else 0.0
    roofline = {
        "kernel": "k_force (get_acc_and_Rrho, Resid.cpp:243-469)", "bound": "hbm", "achieved": achieved,
        "peak": hbm_peak, "peak_source": "measured (MEASURED_PEAKS.json)" if peaks else "fallback", "unit": "GB/s",
        "frac": achieved / hbm_peak, "traffic": traffic,
        "traffic_source": ("profile: profiles/kernel_counters.json (ncu --set full of this build: %s), dram__bytes_read + "
                           "dram__bytes_write per launch, scaled by the particle count; not measured in this run"
                           % counters.get("how", "")) if traffic is not None else None,
        "ms_per_launch": force_ms,
        "algorithmic_bytes_per_launch": FORCE_BYTES_PER_PARTICLE * n_fluid,
        "fp64": {"achieved_tflops": fp64_achieved, "peak_tflops": fp64_peak, "peak_source": fp64_src,
                 "frac": fp64_achieved / fp64_peak, "nominal_peak_tflops": FP64_NOMINAL_TFLOPS,
                 "flop_per_pair": FORCE_FLOP_PER_PAIR, "pairs": pairs,
                 "note": "the pair sweeps are FP64-pipe bound (SURVEY.md 8d), the HBM figure is the yardstick north_star names"},
        "step_hbm_frac": (value / max(1, world)) * STEP_BYTES_PER_PARTICLE / 1e9 / hbm_peak,
    }
    sm_mhz_live = float(clocks.get("sm_mhz") or peaks.get("sm_max_mhz", 1965.0))
    if fcnt and force_ms > 0:
        # FP64-pipe utilisation of the force sweep from the hardware counter (sm__inst_executed_pipe_fp64), not from a
        # flop-per-pair estimate: pipe-busy time of the profile, rescaled to this run's clock, over this run's duration
        roofline["fp64"]["pipe_busy_frac"] = fcnt["fp64_pipe_busy_ms"] / sm_mhz_live / force_ms
        roofline["fp64"]["pipe_busy_source"] = "profile counter sm__inst_executed_pipe_fp64 x this run's CUDA-event duration"
    # per-kernel rooflines: HBM for the streaming kernels, the FP64 pipe for the pair sweeps (SURVEY 8d algorithmic bytes)
    FAMILIES = [("force", ["k_force"], 292.0, True), ("prestep", ["k_prestep"], 204.0, False),
                ("surf1+diss", ["k_surf1_diss"], 460.0, False), ("surf2+3+shift", ["k_surf23_shift"], 252.0, False),
                ("nb_list", ["k_exact_runs"], 40.0, False), ("nb_skin", ["k_build_skin_runs"], 40.0, False),
                ("nb_update", ["k_nb_update"], 152.0, False)]
    roofline_kernels = []
    whole_busy = 0.0
    for fam, prefixes, alg_bytes, variants in FAMILIES:
        tm = timers.get(fam)
        if not tm or tm["calls"] == 0:
            continue
        ms_launch = tm["ms"] / tm["calls"]
        cnt_f = mean_variants(prefixes[0]) if variants else family_counters(prefixes)
        ent = {"family": fam, "ms_per_call": ms_launch, "calls_per_step": tm["calls"] / args.steps,
               "hbm": {"algorithmic_bytes": alg_bytes * n_fluid, "achieved_gbs": alg_bytes * n_fluid / (ms_launch * 1e-3) / 1e9,
                       "frac": alg_bytes * n_fluid / (ms_launch * 1e-3) / 1e9 / hbm_peak}}
        if cnt_f:
            ent["hbm"]["dram_traffic_bytes_profile"] = cnt_f["dram_bytes"]
            ent["hbm"]["dram_frac_profile"] = cnt_f["dram_bytes"] / (ms_launch * 1e-3) / 1e9 / hbm_peak
            ent["fp64_pipe_busy_frac"] = cnt_f["fp64_pipe_busy_ms"] / sm_mhz_live / ms_launch
            ent["profile_kernels"] = cnt_f["kernels"]
            whole_busy += cnt_f["fp64_pipe_busy_ms"] / sm_mhz_live * tm["calls"] / args.steps
        roofline_kernels.append(ent)
    if whole_busy > 0:
        roofline["step_fp64_pipe_busy_frac"] = whole_busy / (ms / args.steps)
    total_ms = sum(v["ms"] for v in timers.values())
    kernels = {k: {"ms_per_step": v["ms"] / args.steps, "launches_per_step": v["launches"] / args.steps,
                   "share": v["ms"] / total_ms if total_ms > 0 else 0.0} for k, v in sorted(timers.items())}

    # ---- end to end: the same steps through the C-ABI call with pinned HOST buffers
    e2e = None
    if not args.no_e2e and world == 1:
        def pinned(a):
            t_ = torch.empty(a.shape, dtype=torch.from_numpy(a).dtype, pin_memory=True)
            t_.numpy()[...] = a
            return t_

        st = e.download(("xi", "v", "acc", "rho", "Rrho", "p", "m", "b"))
        ins_t = {k: pinned(v) for k, v in st.items()}
        out_fields = ("xi", "v", "acc", "rho", "Rrho", "p")
        with torch.cuda.stream(stream):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                ins = {k: v.numpy() for k, v in ins_t.items()}
                outs, s2 = e.step_host(ins, case["bound_points"], 1, out_fields=out_fields, out=ins)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
        h2d = sum(v.numpy().nbytes for v in ins_t.values())
        d2h = sum(ins_t[k].numpy().nbytes for k in out_fields)
        e2e = {"value": n_fluid * args.steps / dt, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "ms_per_step": dt / args.steps * 1e3}
    elif not args.no_e2e:
        # every rank round-trips its own slab through pinned host memory each step (upload -> step -> download)
        def pinned(a):
            t_ = torch.empty(a.shape, dtype=torch.from_numpy(a).dtype, pin_memory=True)
            t_.numpy()[...] = a
            return t_

        fields = ("xi", "v", "acc", "rho", "Rrho", "p", "m", "b", "part_id")
        out_fields = fields  # migration reorders the owned set, so every per-particle array comes back
        with torch.cuda.stream(stream):
            st = e.download(fields)
            ins_t = {k: pinned(v) for k, v in st.items()}
            n_own = st["xi"].shape[0]
            barrier()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                ins = {k: v.numpy()[:n_own] for k, v in ins_t.items()}
                e.reupload_owned(ins)
                e.integrate()
                n_own = e.n
                for k in ins_t:  # migration changes the owned count by a few particles
                    if ins_t[k].shape[0] < n_own:
                        ins_t[k] = pinned(np.resize(ins_t[k].numpy(), (int(n_own * 1.01),) + tuple(ins_t[k].shape[1:])))
                e.download(out_fields, out={k: ins_t[k].numpy()[:n_own] for k in out_fields})
            barrier()
            dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        h2d = sum(v.numpy()[:n_own].nbytes for v in ins_t.values())
        d2h = sum(ins_t[k].numpy()[:n_own].nbytes for k in out_fields)
        e2e = {"value": total_fluid * args.steps / dt, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "ms_per_step": dt / args.steps * 1e3, "note": "bytes are per rank"}

    slab_stats = e.slab_stats() if world > 1 else None
    cpu = cpu_early

    extras = {}
    if world > 1 and not args.no_extras:
        e.close()
        del e
        torch.cuda.empty_cache()
        n_cells = [int(k) for k in args.cells.split(",")]
        if args.workload != "jet":
            # north_star's 100 M-particle jet: 12.5 M particles per GPU, Gissler aero in a cross flow
            extras["workloads"] = {"jet": time_slab_workload(args, rank, world, local_rank, "jet", str(args.jet_columns), 3, 2)}
        # strong scaling: the single-GPU workload (12.5 M particles in total) cut over the N GPUs
        sx = max(16, n_cells[0] // world)
        extras["strong"] = dict(time_slab_workload(args, rank, world, local_rank, "block", "%d,%d,%d" % (sx, n_cells[1], n_cells[2]),
                                                   5, 3), scaling="strong")

    parity = None
    if world > 1 and not args.no_check:
        # after every timed region: the check builds and frees several engines, and a measurement taken after it ran 2.3x
        # slower (profiles/r2e_bench_n2_after_check.json) -- fragmented device memory is the suspect
        if not extras:
            e.close()
            del e
            torch.cuda.empty_cache()
        parity = slab_parity(rank, world, local_rank)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args, world), "particles_total": int(total_fluid), "solver": args.solver,
                       "force_evals_per_step": force_evals / args.steps, "neighbour_builds_per_step": 2,
                       "mean_neighbours": pairs / max(1, n_fluid),
                       "l2": "inputs larger than L2 (state + neighbour list >> 126 MB), no explicit flush"
                       if n_fluid > 2_000_000 else "working set may fit L2 (small workload)",
                       "parallelism": "slab%d" % world if world > 1 else "single"},
            "roofline": roofline, "roofline_kernels": roofline_kernels, "cpu_baseline": cpu, "e2e": e2e,
            "gpu_launches": int(cnt[1].item()),
            "clocks": clocks, "kernels": kernels,
        }
        if world > 1:
            line["slab"] = slab_stats
            line["slab_parity"] = parity
            line.update(extras)
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
