#!/usr/bin/env python
"""Benchmark of the implicit conduction solve (BASELINE.json metric: CG GDOF/s + implicit steps/s on
Q2 hexes; % of HBM roofline).

A "step" is one implicit (backward Euler) time step of the nonlinear unsteady conduction solve --
Newton on R(k) with a preconditioned Krylov solve per Newton iteration -- on a synthetic refined
brick (BASELINE.json configs[3]: Q2, N=128 -> 257^3 = 16 974 593 DOFs, k(T) = 0.09 T - 17,
heat flux on x0, isothermal x1), exactly SURVEY.md section 8d "C4".

value  = N_dof x Krylov iterations / time of the timed steps, vectors resident in HBM;
e2e    = the same through the C-ABI driver with HOST buffers (u copied host->device before and
         device->host after every step inside the timed region);
roofline = the Jacobian-vector apply (the dominant kernel), algorithmic 24 B/DOF, timed with CUDA
         events on the launching stream.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

PROPS = [("Density_rho", "Uniform, 8000"), ("Specific_Heat_C", "Uniform, 500"),
         ("Thermal_Conductivity_k", "Polynomial, 0.09, -17")]
BCS = ["HeatFlux, 7.5e5", "HeatFlux, 0", "HeatFlux, 0", "Isothermal, 300", "HeatFlux, 0", "HeatFlux, 0"]
METRIC = "Krylov GDOF/s (N_dof x Krylov iterations / solve time), nonlinear unsteady implicit step, Q2 hex"


PROPS_B = [("Density_rho", "Uniform, 8000"), ("Specific_Heat_C", "Polynomial, 4.5, -850"),
           ("Thermal_Conductivity_k", "Polynomial, 0.09, -17")]   # SURVEY.md 8d variant B: C(T) exercises the BN terms
VARIANT = {"A": PROPS, "B": PROPS_B}
_variant = "A"


def write_ini(path, order, solver, prec, steps):
    from cases import ini_text
    with open(path, "w") as f:
        f.write(ini_text("Nonlinear_Unsteady", "synthetic-brick", 300, VARIANT[_variant], BCS, newton=100, solver=solver, prec=prec,
                         steps=steps, order=order))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index=0):
        self.rows, self.stop_flag, self.index = [], False, index
        self.thread = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([t.strip() for t in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def __enter__(self):
        self.thread.start()
        return self

    def __exit__(self, *a):
        self.stop_flag = True
        self.thread.join(timeout=6)

    def summary(self):
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.rows)}


CPU_BIN = os.path.join(ROOT, "oracle", "_build", "jots_cpu_port")


def host_threads():
    """Threads the CPU arm may use: every core this process is allowed on (torchrun exports OMP_NUM_THREADS=1 to its
    workers, which is why the thread count is passed explicitly and the variable overridden)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_port_run(order, n, steps, threads=None):
    """The C/OpenMP restatement of the reference's algorithm (oracle/c/jots_cpu_port.c: dense element
    Jacobians -> CSR at every Newton iteration, SpMV-based Jacobi-PCG) on a bounded sample of the same
    workload (same physics, n^3 elements), on all host cores."""
    threads = host_threads() if not threads else threads
    if not os.path.exists(CPU_BIN):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle", "c")], stdout=subprocess.DEVNULL)
    cmd = [CPU_BIN, str(n), str(n), str(n), "0.01", "0.01", "0.01", str(order), "1e-2", str(steps), "8000", "500", "300",
           "7.5e5", "300", "1e-10", "1e-16", "100", "1e-10", "100", str(threads), "2", "0.09", "-17"]
    env = dict(os.environ, OMP_NUM_THREADS=str(threads))
    env.pop("OMP_PROC_BIND", None)
    try:
        out = subprocess.check_output(cmd, env=env)
    except (subprocess.CalledProcessError, OSError):
        # binary built on another CPU generation: rebuild on this host and retry once
        subprocess.check_call(["make", "-B", "-C", os.path.join(ROOT, "oracle", "c")], stdout=subprocess.DEVNULL)
        out = subprocess.check_output(cmd, env=env)
    r = json.loads(out.decode())
    r["gdofs"] = r["ndof"] * r["krylov_iters"] / r["time"] / 1e9
    r["steps_per_s"] = r["steps"] / r["time"]
    return r


def cpu_baseline_entry(r, n, order):
    return {"value": r["gdofs"], "unit": "GDOF/s", "cores": r["threads"], "kind": "port",
            "sample": "same physics on a %d^3-element Q%d brick (%d DOFs), %d steps, %.1f s; C/OpenMP restatement of the "
                      "reference's algorithm (CSR re-assembly per Newton iteration + SpMV Jacobi-PCG); the reference itself "
                      "(MFEM/hypre/MPI) cannot be built here" % (n, order, r["ndof"], r["steps"], r["time"]),
            "steps_per_s": r["steps_per_s"], "krylov_iters": r["krylov_iters"]}


def cpu_oracle_run(order, solver, prec, n, steps):
    """The NumPy/SciPy oracle on a bounded sample (kept for solver/preconditioner variants the C port lacks)."""
    from oracle import jots as OJ, mesh as OM
    with tempfile.TemporaryDirectory() as d:
        ini = os.path.join(d, "c.ini")
        write_ini(ini, order, solver, prec, steps)
        drv = OJ.JOTSDriver(OJ.Config(ini), mesh=OM.make_brick(n, n, n, 0.01, 0.01, 0.01))
        t0 = time.perf_counter()
        drv.run()
        dt = time.perf_counter() - t0
    st = drv.it.stats
    return dict(ndof=drv.space.ndof, time=dt, krylov_iters=st["krylov_iters"], steps=st["steps"],
                gdofs=drv.space.ndof * st["krylov_iters"] / dt / 1e9, steps_per_s=st["steps"] / dt)


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle port: full sparse assembly per Newton
    iteration + SpMV Krylov; the reference itself cannot be built here)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = cpu_sample_n(args)
    use_c = (args.solver == "CG" and args.prec == "Jacobi")
    if use_c:
        if args.warmup:
            cpu_port_run(args.order, max(4, n // 4), 1)
        r = cpu_port_run(args.order, n, args.steps)
    else:
        n = min(n, 12)
        r = cpu_oracle_run(args.order, args.solver, args.prec, n, args.steps)
        r["threads"] = 1
    # the config names the workload of the GPU arm AND the bounded sample of it that was actually timed here
    cfg = workload_config(args, max(1, args.gpus))
    cfg["workload"] += "; TIMED HERE: a bounded sample of it, %dx%dx%d Q%d hexes = %d DOFs per step (same physics, solver and tolerances)" % (
        n, n, n, args.order, r["ndof"])
    cfg["sample_n_elem"] = [n, n, n]
    cfg["sample_n_dof"] = r["ndof"]
    line = {"impl": "reference", "metric": METRIC, "value": r["gdofs"], "unit": "GDOF/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * r["time"] / max(r["steps"], 1), "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": cfg,
            "steps_per_s": r["steps_per_s"],
            "cpu_baseline": cpu_baseline_entry(r, n, args.order),
            "e2e": {"value": r["gdofs"], "unit": "GDOF/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def cpu_sample_n(args):
    """Elements per direction of the CPU sample.  The port's cost per step grows like n^4 (DOFs x Krylov iterations), so the
    sample is as large as a run of steps + warm-up steps allows within a few minutes: small samples spend a larger share in
    the per-Newton CSR assembly and understate the CPU's GDOF/s."""
    if args.cpu_n > 0:
        return args.cpu_n
    total = args.steps + (1 if args.warmup else 0)
    return 64 if total <= 4 else 56 if total <= 8 else 48 if total <= 24 else 40


def box_dims(world):
    d = [1, 1, 1]
    r, a = world, 0
    while r > 1:
        d[a % 3] *= 2
        r //= 2
        a += 1
    return d


def global_brick(args, world):
    """elements per direction of the global brick: fixed (strong) or n per GPU in every direction (weak)"""
    if args.scaling == "strong" or world == 1:
        return [args.n, args.n, args.n]
    return [args.n * k for k in box_dims(world)]


def workload_config(args, world):
    ne = global_brick(args, world)
    nd = 1
    for k in ne:
        nd *= k * args.order + 1
    # BASELINE.json: configs[3] = Q2, ~16 M DOFs on one B200 (the default); configs[4] = Q3, ~128 M DOFs over 2/4/8 B200
    which = "configs[4]" if args.order == 3 else "configs[3]" if args.order == 2 else "configs[3] physics at another order"
    return {"workload": "%s: synthetic refined brick, %dx%dx%d Q%d hexes, %d DOFs, nonlinear unsteady (backward Euler, "
                        "dt=1e-2), k(T)=0.09T-17%s, Newton + %s/%s rel 1e-10" % (which, ne[0], ne[1], ne[2], args.order, nd,
                                                                                    ", C(T)=4.5T-850" if args.variant == "B" else "", args.solver, args.prec),
            "n_elem": ne, "order": args.order, "n_dof": nd, "solver": args.solver, "preconditioner": args.prec,
            "parallelism": ("box partition %s, 1 subdomain per GPU, NCCL interface sum-exchange + scalar all-reduce" % "x".join(map(str, box_dims(world))))
                           if world > 1 else "single GPU",
            "l2_policy": "vectors (136 MB each at 128^3 elements per GPU) exceed the 126 MB L2; no explicit flush"}


def partition_parity_check(args, J, dist, comm, rank, world, local):
    """N > 1: the same small problem (24^3 Q-order hexes, this bench's physics and solver, 2 steps) solved by the partitioned
    ranks and by rank 0 alone on one GPU: relative l2 difference of the temperature fields (north_star bound 1e-8) and
    equality of the Newton / Krylov iteration counts.  Evidence in the SCALE record that the multi-GPU numbers are numbers of
    a correct solve."""
    import numpy as np
    n = 24
    tmp = tempfile.mkdtemp()
    ini = os.path.join(tmp, "parity.ini")
    write_ini(ini, args.order, args.solver, args.prec, 2)
    drv = J.Driver(ini, local, mesh=J.Mesh.brick(n, n, n, 0.01, 0.01, 0.01), comm=comm)
    drv.run()
    gid, n_owned = drv.global_ids()
    st = drv.ctx.stats()
    mine = (np.asarray(gid[:n_owned]), drv.get_u()[:n_owned], st["krylov_iters"], st["newton_iters"])
    parts = [None] * world if rank == 0 else None
    dist.gather_object(mine, parts, dst=0)
    del drv
    if rank != 0:
        return None
    ref = J.Driver(ini, local, mesh=J.Mesh.brick(n, n, n, 0.01, 0.01, 0.01))
    ref.run()
    u_ref = ref.get_u()
    st1 = ref.ctx.stats()
    u = np.full(u_ref.shape, np.nan)
    for g, v, _, _ in parts:
        u[g] = v
    rel = float(np.linalg.norm(u - u_ref) / np.linalg.norm(u_ref))
    return {"n_elem": [n, n, n], "order": args.order, "steps": 2, "rel_l2_vs_single_gpu": rel, "tolerance": 1e-8, "ok": bool(rel <= 1e-8),
            "krylov_iters": {"partitioned": parts[0][2], "single_gpu": st1["krylov_iters"]},
            "newton_iters": {"partitioned": parts[0][3], "single_gpu": st1["newton_iters"]},
            "all_dofs_covered": bool(np.isfinite(u).all())}


def strong_scaling_leg(args, J, torch, dist, comm, rank, world, local, barrier):
    """A FIXED global brick (default 256^3 Q2 hexes = 135 005 697 DOFs, domain 0.01^3 for every N) box-partitioned over the
    N GPUs: 1 warm-up + 2 timed steps, device timed, max over ranks.  north_star's strong-scaling number."""
    ne = args.strong_elems
    tmp = tempfile.mkdtemp()
    ini = os.path.join(tmp, "strong.ini")
    write_ini(ini, args.order, args.solver, args.prec, 10 ** 9)
    t0 = time.perf_counter()
    drv = J.Driver(ini, local, mesh=J.Mesh.brick(ne, ne, ne, 0.01, 0.01, 0.01), comm=comm)
    setup_s = time.perf_counter() - t0
    stream = torch.cuda.current_stream()
    drv.ctx.set_stream(stream.cuda_stream)
    drv.run(1)
    barrier()
    drv.ctx.reset_stats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    drv.run(2)
    e1.record(stream)
    barrier()
    st = drv.ctx.stats()
    t = torch.tensor([e0.elapsed_time(e1), float(st["krylov_iters"]), setup_s], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    del drv
    nd = (ne * args.order + 1) ** 3
    ms = float(t[0])
    return {"n_elem": [ne, ne, ne], "domain": [0.01, 0.01, 0.01], "n_dof": nd, "steps": 2, "warmup": 1, "ms_per_step": ms / 2,
            "value": nd * int(t[1]) / (ms * 1e-3) / 1e9, "unit": "GDOF/s", "krylov_iters": int(t[1]), "newton_iters": st["newton_iters"],
            "host_setup_s": float(t[2]), "scaling": "strong"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--elems", dest="n", type=int, default=128, help="elements per direction of the synthetic brick")
    ap.add_argument("--order", type=int, default=2)
    ap.add_argument("--solver", default="CG")
    ap.add_argument("--prec", default="Jacobi")
    ap.add_argument("--cpu-n", type=int, default=0, help="elements per direction of the CPU sample (0: by the number of steps, 40..64)")
    ap.add_argument("--apply-reps", type=int, default=100, help="timed operator applies (after 20 warm-up applies)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--variant", default="A", choices=["A", "B"], help="B: C(T) = 4.5 T - 850 as well (FGMRES recommended)")
    ap.add_argument("--strong-elems", type=int, default=256,
                    help="elements per direction of the FIXED global brick of the extra strong-scaling leg, run at every N so that the N = 1, 2, 4, 8 "
                         "lines give north_star's strong-scaling curve (256 -> 135 M DOFs at Q2; 0: off)")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling leg (it costs about half a minute of host set-up at every N)")
    ap.add_argument("--no-multi-checks", action="store_true", help="N > 1: skip the partitioned-vs-single-GPU parity check")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="N>1: weak = --n elements per direction per GPU, strong = --n for the global brick")
    args = ap.parse_args()
    global _variant
    _variant = args.variant
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import jots_b200 as J

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if J.load_library().jots_device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device: jots_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    t_start = time.perf_counter()

    def trace(msg):
        if os.environ.get("JOTS_BENCH_TRACE"):
            print("[bench rank %d %.1fs] %s" % (rank, time.perf_counter() - t_start, msg), file=sys.stderr, flush=True)

    W, K = args.warmup, args.steps
    tmp = tempfile.mkdtemp()
    ini = os.path.join(tmp, "bench.ini")
    write_ini(ini, args.order, args.solver, args.prec, 10 ** 9)
    ne = global_brick(args, world)
    mesh = J.Mesh.brick(ne[0], ne[1], ne[2], 0.01 * ne[0] / args.n, 0.01 * ne[1] / args.n, 0.01 * ne[2] / args.n)
    comm = None
    if world > 1:
        uid = [J.Comm.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        comm = J.Comm(uid[0], rank, world, local)
    trace("mesh + communicator ready")
    drv = J.Driver(ini, local, mesh=mesh, comm=comm)
    del mesh
    trace("driver ready (n_local = %d)" % drv.n)
    stream = torch.cuda.current_stream()
    drv.ctx.set_stream(stream.cuda_stream)
    ndof = drv.n

    # ---- device-resident timing ---------------------------------------------------------------
    for _ in range(W):
        drv.run(1)
        trace("warm-up step done")
    barrier()
    drv.ctx.reset_stats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        e0.record(stream)
        for _ in range(K):
            drv.run(1)
        e1.record(stream)
        barrier()
    ms = e0.elapsed_time(e1)
    st = drv.ctx.stats()
    t = torch.tensor([ms, float(st["krylov_iters"])], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t[0])
    iters = int(t[1])
    total_dof = workload_config(args, world)["n_dof"]
    value = total_dof * iters / (ms * 1e-3) / 1e9

    trace("timed steps done: %.1f ms" % ms)
    # ---- end to end through the C ABI with host buffers ---------------------------------------------
    u_host = torch.empty(ndof, dtype=torch.float64).pin_memory().numpy()
    drv.get_u(u_host)
    drv.set_u(u_host); drv.run(1); drv.get_u(u_host)
    barrier()
    drv.ctx.reset_stats()
    t0 = time.perf_counter()
    for _ in range(K):
        drv.set_u(u_host)
        drv.run(1)
        drv.get_u(u_host)
    barrier()
    e2e_s = time.perf_counter() - t0
    st2 = drv.ctx.stats()
    t = torch.tensor([e2e_s, float(st2["krylov_iters"])], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = total_dof * int(t[1]) / float(t[0]) / 1e9

    trace("e2e done")
    # ---- dominant kernel: Jacobian-vector apply -------------------------------------------------------
    x = torch.empty(ndof, dtype=torch.float64, device="cuda").uniform_(-1.0, 1.0)
    y = torch.empty_like(x)
    z = torch.empty_like(x)
    drv.get_u(z)
    torch.cuda.synchronize()
    drv.op.form_time(J.FORM_SYSTEM, x, y, state_dev=z, reps=min(20, max(2, args.apply_reps)))   # warm-up applies
    apply_plain_ms = drv.op.form_time(J.FORM_SYSTEM, x, y, state_dev=z, reps=args.apply_reps)
    apply_ms, fused = apply_plain_ms, st["fused_dots"] > 0
    if fused:
        # the conjugate-gradient loop launched the variant that also forms (d, A d) -- and, when it could, carries the loop's
        # side stream (x += alpha d of the previous iteration, zeroing of the next output): that is the dominant kernel of the step
        os.environ["JOTS_FORM_TIME_DOT"] = "2" if st.get("side_applies", 0) > 0 else "1"
        drv.op.form_time(J.FORM_SYSTEM, x, y, state_dev=z, reps=min(20, max(2, args.apply_reps)))
        apply_ms = drv.op.form_time(J.FORM_SYSTEM, x, y, state_dev=z, reps=args.apply_reps)
        os.environ.pop("JOTS_FORM_TIME_DOT")
    kernel_label = drv.ctx.last_kernel()          # the element kernel these applies actually launched
    res_ms = drv.op.form_time(J.FORM_RESIDUAL, x, y, state_dev=z, reps=max(2, args.apply_reps // 4))
    peaks = {}
    pk_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk_path):
        peaks = json.load(open(pk_path))
    peak = float(peaks.get("hbm_gbs", 6650.0))
    alg_bytes = 24.0 * ndof
    achieved = alg_bytes / (apply_ms * 1e-3) / 1e9
    # ncu --set full figures of the SAME kernel (label and vector length must match, else they are left out):
    # dram__bytes_read + dram__bytes_write, FP64 pipe utilisation, FP64 thread instructions per launch
    traffic, fp64_pct, fp64_instr, stats_src = None, None, None, None
    ks_path = os.path.join(ROOT, "profiles", "kernel_stats.json")
    if os.path.exists(ks_path):
        for e in json.load(open(ks_path)):
            if e.get("kernel") == kernel_label and e.get("n_dof") == ndof:
                traffic, fp64_pct = e["dram_bytes_per_launch"], e.get("fp64_pipe_active_pct")
                fp64_instr, stats_src = e.get("fp64_thread_instr_per_launch"), e.get("source")

    # ---- N > 1: correctness of the partitioned solve + a strong-scaling leg, both in the driver's view ----------------
    extra = {}
    if world > 1 and not args.no_multi_checks:
        extra["parity_check"] = partition_parity_check(args, J, dist, comm, rank, world, local)
        trace("parity check done")
    if args.strong_elems > 0 and not args.no_strong:
        extra["strong"] = strong_scaling_leg(args, J, torch, dist if world > 1 else None, comm, rank, world, local, barrier)
        trace("strong-scaling leg done")

    def finish():
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()

    if rank != 0:
        del drv, comm
        return finish()
    line = {"metric": METRIC, "value": value, "unit": "GDOF/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": workload_config(args, world),
            "steps_per_s": K / (ms * 1e-3), "krylov_iters": iters, "newton_iters": st["newton_iters"],
            "apply_gdofs": ndof / (apply_ms * 1e-3) / 1e9, "apply_ms": apply_ms, "apply_without_inner_product_ms": apply_plain_ms,
            "residual_ms": res_ms, "fused_inner_products": st["fused_dots"], "side_stream_applies": st.get("side_applies", 0),
            "clocks": clk.summary(),
            "e2e": {"value": e2e_value, "unit": "GDOF/s", "h2d_bytes_per_step": 8 * ndof, "d2h_bytes_per_step": 8 * ndof,
                    "ms_per_step": 1e3 * e2e_s / K},
            "gpu_launches": st["kernel_launches"],
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "algorithmic_bytes": alg_bytes,
                         # what the launch moves on top of the apply's 24 B/DOF when it carries the CG loop's side stream (x in / out,
                         # d_old in, zeroing of the next output: 32 B/DOF); `achieved` does not count it, `traffic` contains it
                         "side_stream_bytes": (32.0 * ndof if st.get("side_applies", 0) > 0 else 0.0),
                         "kernel": kernel_label,
                         "timed": "Jacobian-vector apply, 24 B/DOF algorithmic (x, z in, y out); CUDA events on the launching stream around "
                                  "%d applies, each = %s" % (args.apply_reps,
                                      "this kernel as the CG loop launches it: it also forms (x, A x) and carries the loop's side stream "
                                      "(x_sol += alpha d_old, zeroing of the next output: 4 more vector passes, not counted in the 24 B/DOF)"
                                      if st.get("side_applies", 0) > 0 else
                                      "zeroing of y + this kernel" + (" (which also forms (x, A x), as in the CG loop)" if fused else "")),
                         "note": "this kernel is FP64-pipe bound (fp64.instr_per_dof FP64 instructions against 24 B/DOF), not HBM "
                                 "bound: fp64.frac_of_peak is the fraction that matters (DESIGN.md 5)",
                         "fp64_pipe_active_pct": fp64_pct,
                         "kernel_stats_source": stats_src,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6.65 TB/s",
                         "frac_of_8TBs_spec": achieved / 8000.0}}
    clk_mhz = (line["clocks"].get("sm_mhz") or line["clocks"].get("sm_max_mhz") or 1965.0)
    if fp64_instr:
        # FP64 peak of this GPU: SMs x 64 FP64 lanes per clock x SM clock under load (instructions, not flops)
        sms = torch.cuda.get_device_properties(local).multi_processor_count
        fp64_peak = sms * 64 * clk_mhz * 1e6
        line["roofline"]["fp64"] = {"instr_per_launch": fp64_instr, "instr_per_dof": fp64_instr / ndof,
                                    "achieved_ginstr_s": fp64_instr / (apply_ms * 1e-3) / 1e9, "peak_ginstr_s": fp64_peak / 1e9,
                                    "frac_of_peak": fp64_instr / (apply_ms * 1e-3) / fp64_peak,
                                    "peak_definition": "%d SMs x 64 FP64 lanes/clk x %.0f MHz (median SM clock in the timed region)" % (sms, clk_mhz)}
    if not args.no_cpu and world == 1 and args.variant == "A":
        if args.solver == "CG" and args.prec == "Jacobi":
            n_cpu = args.cpu_n if args.cpu_n > 0 else 48
            line["cpu_baseline"] = cpu_baseline_entry(cpu_port_run(args.order, n_cpu, 2), n_cpu, args.order)
        else:
            r = cpu_oracle_run(args.order, args.solver, args.prec, 12, 2)
            r["threads"] = 1
            line["cpu_baseline"] = cpu_baseline_entry(r, 12, args.order)
    line.update(extra)
    print(json.dumps(line), flush=True)
    del drv, comm
    finish()


if __name__ == "__main__":
    main()
