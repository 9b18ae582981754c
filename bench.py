#!/usr/bin/env python
"""bench.py -- MLUPS of the collide-and-stream hot path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (config.workload): BASELINE.json configs[1] -- 3D lid-driven cavity, D3Q19 BGK fp64, 512^3
per GPU, regularized velocity boundaries on all faces / edges / corners (Palabos cavity3d), synthetic
initial state rho=1, u=0.  With N GPUs the cavity is (512 N) x 512 x 512, one x-slab per rank
(weak scaling; the rim cells of the sweep store the face-crossing populations straight into the neighbouring rank's slab
over NVLink, NCCL send/recv is the fallback).  Multi-rank runs start with an untimed parity check of that transport
("parity_check": small lattices stepped over the N ranks against a one-slab run).

One JSON line on rank 0:
  value     whole-job MLUPS, state resident in HBM, CUDA events, max over ranks
  e2e       the same job through the public C-ABI from HOST buffers: upload of the initial
            populations (pinned host memory), K x {setBoundaryVelocity on the lid, collideAndStream,
            read the stored statistics}, download of the final populations -- all inside the timed
            region
  roofline  sweep kernel: 2*q*sizeof(T) algorithmic bytes per cell / its CUDA-event duration,
            against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline  the reference's own CPU code (oracle/_ref/ref_driver, or the C port when that binary
            is absent) on a bounded sample of the same workload, rank 0, N=1 only
--impl reference times that CPU implementation with as many host processes as there are cores.

--workload runs another BASELINE.json config (channel_trt, obstacle, les_q27_f64, les_q27_f32) or ext_cavity, the time loop
of the fork's own application (external-moment sweep), and prints the same line for it.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

U_MAX, RE = 0.02, 100.0

# BASELINE.json configs.  "cavity" (configs[1]) is the bench line the driver reads; the others are run on
# request (--workload) and print the same line for their own config.
WORKLOADS = {
    "cavity":      dict(q=19, prec="f64", dyn="BGK",   scaling="weak",   shape=lambda N, n: (n * N, n, n)),
    "channel_trt": dict(q=19, prec="f64", dyn="TRT",   scaling="strong", shape=lambda N, n: (2048, 512, 512)),
    "obstacle":    dict(q=19, prec="f64", dyn="BGK",   scaling="strong", shape=lambda N, n: (1024, 384, 384)),
    "les_q27_f64": dict(q=27, prec="f64", dyn="SMAGO", scaling="weak",   shape=lambda N, n: (n * N, n, n)),
    "les_q27_f32": dict(q=27, prec="f32", dyn="SMAGO", scaling="weak",   shape=lambda N, n: (n * N, n, n)),
    # the time loop of the fork's own application (viscosityTest) -- IncBGK collided with the moments of
    # the rhoBar / j fields (ExternalRhoJcollideAndStream3D), the fields recomputed after every step (BoxRhoBarJfunctional3D)
    "ext_cavity":  dict(q=19, prec="f64", dyn="INCBGK", scaling="weak",  shape=lambda N, n: (n * N, n, n)),
}


def kernel_sources_hash():
    """sha256 over the sources of the sweep kernel -- the key profiles/traffic.json is valid for"""
    import hashlib
    h = hashlib.sha256()
    for f in ("lb_core.cuh", "cell_io.cuh", "sweep_kernel.cuh", "kernels.cuh", "sweep_launch.cuh", "Makefile"):
        h.update(open(os.path.join(ROOT, "palabos_b200", "csrc", f), "rb").read())
    return h.hexdigest()


def cavity_omega(ny):
    nu = U_MAX * ny / RE
    return 1.0 / (3 * nu + 0.5)


# ----------------------------------------------------------------------------------- reference arm


def cpu_reference_run(n, steps, procs):
    """MLUPS of the reference's CPU implementation on an n^3 cavity, `procs` independent serial
    processes side by side (the reference has no threads; its MPI build is not available here)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from cases import cavity
    import scenario
    sc = cavity(n, n, n, 0, dumps=False)
    sc.ops = [op for op in sc.ops if op[0] not in ("stats",)]
    if scenario.have_reference_driver("f64"):
        kind = "reference"

        def one(out, k):
            out[k] = scenario.run_reference(sc, timed_steps=steps)["time"]
    else:
        kind = "port"

        def one(out, k):
            eng = scenario.OracleEngine(sc.q, sc.prec, sc.nx, sc.ny, sc.nz, sc.dyn, sc.p1, sc.p2)
            scenario.replay(sc, eng)
            t0 = time.perf_counter()
            eng.collide_and_stream(steps)
            dt = time.perf_counter() - t0
            out[k] = (steps, dt, n ** 3 * steps / dt / 1e6)
            eng.close()
    out = [None] * procs
    t0 = time.perf_counter()
    if kind == "reference":
        th = [threading.Thread(target=one, args=(out, k)) for k in range(procs)]   # each thread waits on a subprocess
        [t.start() for t in th]
        [t.join() for t in th]
    else:
        # the ctypes port holds no GIL while in C, threads are enough
        th = [threading.Thread(target=one, args=(out, k)) for k in range(procs)]
        [t.start() for t in th]
        [t.join() for t in th]
    wall = time.perf_counter() - t0
    slowest = max(o[1] for o in out)
    mlups = procs * n ** 3 * steps / slowest / 1e6
    return {"value": mlups, "unit": "MLUPS", "cores": procs, "kind": kind,
            "sample": "%d^3 lid-driven cavity D3Q19 BGK fp64, %d steps, %d serial process(es) of the reference side by side, "
                      "slowest process %.2f s (wall incl. set-up %.1f s)" % (n, steps, procs, slowest, wall)}


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    procs = max(1, min(cores, 64))
    t0 = time.perf_counter()
    runs = []
    for k in range(args.warmup + args.steps):
        # a "step" of this arm = a bounded sample: 96^3 cavity, 10 time steps per process
        r = cpu_reference_run(96, 10, procs)
        if k >= args.warmup:
            runs.append(r)
        if time.perf_counter() - t0 > 240 and runs:
            break
    v = statistics.median(r["value"] for r in runs)
    base = dict(runs[-1]); base["value"] = v
    line = {"impl": "reference", "metric": "MLUPS", "value": v, "unit": "MLUPS", "n_gpus": args.gpus, "steps": len(runs),
            "warmup": args.warmup,
            # a "step" of this arm is one bounded sample: `procs` cavities of 96^3, 10 time steps each
            "ms_per_step": procs * 96 ** 3 * 10 / (v * 1e6) * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "3D lid-driven cavity D3Q19 BGK fp64, regularized velocity BC (reference CPU code, bounded sample 96^3)"},
            "cpu_baseline": base,
            "e2e": {"value": v, "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------- our arm


class ClockSampler:
    QUERY = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.proc = None
        self.path = tempfile.NamedTemporaryFile(delete=False, suffix=".csv").name
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.QUERY, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if not self.proc:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in open(self.path):
            w = [x.strip() for x in ln.split(",")]
            if len(w) < 7:
                continue
            try:
                sm.append(float(w[0])); mx.append(float(w[1]))
            except ValueError:
                continue
            for nme, val in zip(names, w[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nme)
        os.unlink(self.path)
        if sm:
            out = {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}
        return out


def bind_near_gpu(torch, local):
    """Multi-rank runs: keep this process (and with it the pinned host buffers it allocates -- the e2e leg moves 20 GB per rank
    each way) on the CPU cores, hence the NUMA node, next to its GPU.  Round 1's e2e leg fell from 52 GB/s per GPU at N = 1 to
    14 GB/s at N = 8 with all ranks' buffers wherever the launcher had put the processes.  Best effort: any failure leaves the
    affinity as it was.  Returns what it did, for the JSON line."""
    try:
        p = torch.cuda.get_device_properties(local)
        bdf = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        cpus = open("/sys/bus/pci/devices/%s/local_cpulist" % bdf).read().strip()
        ids = set()
        for part in cpus.split(","):
            a, _, b = part.partition("-")
            ids.update(range(int(a), int(b or a) + 1))
        ids &= os.sched_getaffinity(0)
        if not ids:
            return None
        os.sched_setaffinity(0, ids)
        node = open("/sys/bus/pci/devices/%s/numa_node" % bdf).read().strip()
        return {"gpu": bdf, "numa_node": int(node), "cpus": cpus}
    except Exception as e:          # noqa: BLE001
        return {"error": str(e)[:120]}


def nccl_parity_check(pb, dist, torch, rank, world):
    """Untimed prologue of the N > 1 runs (the driver's test box has one GPU, its bench boxes have N): a small lid-driven cavity
    stepped over the N-rank lattice -- the very exchange path the timed steps use -- must give, bit for bit, the populations of
    the same engine run on ONE slab on rank 0 (which tests/test_gpu_parity.py pins against the oracle and the reference's dumps)."""
    import numpy as np
    nx, ny, nz, steps = max(24, 5 * world + 3), 20, 18, 25          # odd step count: the download reads across the slab faces
    out = {}
    for name, q, prec, dyn, omega in (("cavity D3Q19 BGK f64", 19, "f64", "BGK", cavity_omega(ny)),
                                      ("periodic D3Q27 SMAGO f32 + BB walls", 27, "f32", "SMAGO", 1.9)):
        def run(lat):
            box = lat.box
            if dyn == "BGK":
                lat.set_velocity_condition_on_block_boundaries("local", box, box)
                lat.set_boundary_velocity(box, (0, 0, 0))
                lat.set_boundary_velocity((1, nx - 2, 1, ny - 2, nz - 1, nz - 1), (U_MAX, 0, 0))
                lat.initialize_at_equilibrium(box, 1.0, (0, 0, 0))
            else:
                lat.set_periodicity(True, False, True)
                lat.define_dynamics((0, nx - 1, 0, 0, 0, nz - 1), "BB", 1.0)
                lat.define_dynamics((0, nx - 1, ny - 1, ny - 1, 0, nz - 1), "BB", 1.0)
                lat.initialize_at_equilibrium(box, 1.0, (0.05, 0, 0))
                lat.initialize_at_equilibrium((1, 3, 2, 5, 0, nz - 1), 1.02, (0.02, 0.01, -0.01))
                lat.initialize_at_equilibrium((nx - 2, nx - 1, 2, 9, 4, 9), 0.99, (0.04, -0.01, 0.02))
            lat.initialize()
            lat.collide_and_stream(steps)
            return lat.download()
        lat = pb.Lattice(q, prec, nx, ny, nz, dyn, omega, 0.14 if dyn == "SMAGO" else 0.0, rank=rank, n_ranks=world)
        uid = [pb.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        lat.comm_init(uid[0])
        mine = run(lat)
        halo = lat.halo_transport() if hasattr(lat, "halo_transport") else "nccl"
        lat.close()
        parts = [None] * world
        dist.all_gather_object(parts, mine.reshape(-1, q))
        if rank == 0:
            single = pb.Lattice(q, prec, nx, ny, nz, dyn, omega, 0.14 if dyn == "SMAGO" else 0.0)
            want = run(single).reshape(-1, q)
            single.close()
            got = np.concatenate(parts, axis=0)
            same = got.shape == want.shape and got.tobytes() == want.tobytes() and bool(np.isfinite(want).all()) and float(np.abs(want).max()) > 0
            out[name] = "identical" if same else "DIFFERENT"
            out["transport"] = halo
    if rank == 0:
        res = {"nccl_vs_single": "identical" if all(v == "identical" for k, v in out.items() if k != "transport") else "DIFFERENT",
               "ranks": world, "lattice": "%dx%dx%d, %d steps" % (nx, ny, nz, steps), "cases": out}
        return res
    return None


def our_arm(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import palabos_b200 as pb

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node %d" % args.gpus)
    assert torch.cuda.is_available(), "bench.py needs a CUDA device: the engine has no CPU fallback"
    torch.cuda.set_device(local)
    numa = bind_near_gpu(torch, local) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    parity = nccl_parity_check(pb, dist, torch, rank, world) if world > 1 else None
    if parity and parity["nccl_vs_single"] != "identical":
        raise SystemExit("multi-rank parity check failed: %s" % json.dumps(parity))

    n = args.n
    wl = WORKLOADS[args.workload]
    Q, PREC = wl["q"], wl["prec"]
    BYTES_PER_LU = 2 * Q * (8 if PREC == "f64" else 4)
    nx, ny, nz = wl["shape"](world, n)
    cavity = args.workload == "cavity"
    ext = args.workload == "ext_cavity"
    if ext:
        # nobody touches the fields between the steps of this workload, so the sweep recomputes the moments from the populations
        # it gathers: 2*Q*8 bytes per cell like the plain sweep; PLB_B200_EXT_EXPLICIT=1 stores the
        # fields after every step and reads them back (4 more values per cell in the sweep, and the moment pass)
        BYTES_PER_LU = (2 * Q + (4 if os.environ.get("PLB_B200_EXT_EXPLICIT", "0")[:1] == "1" else 0)) * 8
        omega, p2 = cavity_omega(ny), 0.0
        what = ("3D lid-driven cavity D3Q19 IncBGK fp64 stepped by executeInternalProcessors(): ExternalRhoJcollideAndStream3D (level 0) + "
                "BoxRhoBarJfunctional3D (level 3), %dx%dx%d (%d^3 per GPU) + the envelope layer, regularized velocity BC" % (nx, ny, nz, n))
    elif cavity:
        omega, p2 = cavity_omega(ny), 0.0
        what = "3D lid-driven cavity D3Q19 BGK fp64, %dx%dx%d (%d^3 per GPU), regularized velocity BC on faces/edges/corners" % (nx, ny, nz, n)
    elif args.workload == "channel_trt":
        omega, p2 = 1.2, 0.0
        what = "periodic channel D3Q19 TRT fp64 (s+=1.2, s-=1.1), %dx%dx%d, periodic x,z, BounceBack walls y=0,ny-1, uniform initial flow u=0.05" % (nx, ny, nz)
    elif args.workload == "obstacle":
        omega, p2 = 1.5, 0.0
        what = ("flow past a voxelised obstacle D3Q19 BGK fp64 (omega 1.5), %dx%dx%d, Zou-He velocity inlet / pressure outlet, BounceBack side walls "
                "and on the voxels of ball.stl" % (nx, ny, nz))
    else:
        omega, p2 = 1.99, 0.14
        what = "turbulent channel D3Q27 Smagorinsky LES %s (omega0 1.99, Cs 0.14), %dx%dx%d (%d^3 per GPU), periodic x,z, BounceBack walls" % (PREC, nx, ny, nz, n)
    lat = pb.Lattice(Q, PREC, nx, ny, nz, wl["dyn"], omega, p2, rank=rank, n_ranks=world, slabs_per_rank=1,
                     enveloped_periodic=(0, 0, 0) if ext else None)
    if world > 1:
        uid = [pb.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        lat.comm_init(uid[0])
    box = lat.box
    lid = (1, nx - 2, 1, ny - 2, nz - 1, nz - 1)
    if cavity or ext:
        lat.set_velocity_condition_on_block_boundaries("local", box, box)
        lat.set_boundary_velocity(box, (0, 0, 0))
        lat.set_boundary_velocity(lid, (U_MAX, 0, 0))
        lat.initialize_at_equilibrium(box, 1.0, (0, 0, 0))
    elif args.workload == "obstacle":
        lat.add_velocity_boundary("zouhe", 0, -1, (0, 0, 1, ny - 2, 1, nz - 2))
        lat.add_pressure_boundary("zouhe", 0, +1, (nx - 1, nx - 1, 1, ny - 2, 1, nz - 2))
        for wall in ((0, nx - 1, 0, 0, 0, nz - 1), (0, nx - 1, ny - 1, ny - 1, 0, nz - 1),
                     (0, nx - 1, 0, ny - 1, 0, 0), (0, nx - 1, 0, ny - 1, nz - 1, nz - 1)):
            lat.define_dynamics(wall, "BB", 1.0)
        # the obstacle: the reference's ball.stl (its triangles are kept with the golden voxel matrices), scaled to 40 % of
        # the channel height and voxelised on the device (plb_b200_mesh_to_lattice + plb_b200_voxelize); inside and
        # innerBorder voxels become BounceBack, as defineDynamics(lattice, voxelMatrix, flag, BounceBack) does
        soup = np.load(os.path.join(ROOT, "tests", "golden", "voxel_ball_40.npz"))["soup"]
        tri, vdims, _, _ = pb.mesh_to_lattice(soup, int(0.4 * min(ny, nz)), 0, 4)
        flags, vox_ms = pb.voxelize(tri, vdims, 1, return_ms=True)
        mask = np.zeros((nx, ny, nz), dtype=np.int32)
        ox, oy, oz = nx // 3 - vdims[0] // 2, (ny - vdims[1]) // 2, (nz - vdims[2]) // 2
        mask[ox:ox + vdims[0], oy:oy + vdims[1], oz:oz + vdims[2]] = (flags >= 3)
        what += "; voxelizer: %d triangles -> %dx%dx%d flags in %.3f ms on the device" % ((tri.shape[0],) + tuple(vdims) + (vox_ms,))
        lat.define_dynamics_mask(mask, 1, "BB", 1.0)
        del mask
        lat.set_boundary_velocity(box, (0.03, 0, 0))
        lat.set_boundary_density(box, 1.0)
        lat.initialize_at_equilibrium(box, 1.0, (0.03, 0, 0))
    else:
        lat.set_periodicity(True, False, True)
        lat.define_dynamics((0, nx - 1, 0, 0, 0, nz - 1), "BB", 1.0)
        lat.define_dynamics((0, nx - 1, ny - 1, ny - 1, 0, nz - 1), "BB", 1.0)
        lat.initialize_at_equilibrium(box, 1.0, (0.05, 0, 0))
    if ext:
        lat.apply_rhobar_j()            # src/variables.hh:650; lattice.initialize() would be a whole step in this mode
    else:
        lat.initialize()
    lat.synchronize()
    step = lat.execute_internal_processors if ext else lat.collide_and_stream

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    cells_global = nx * ny * nz
    cells_local = lat.local_cells()

    # ---- value: state resident in HBM ----------------------------------------------------------
    step(args.warmup)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    launches0 = lat.launch_count()
    lat.profile_sweeps(True)
    # one event per step: the whole region is timed e[0] -> e[K] as before, the per-step distribution comes for free
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    ev[0].record()
    for k in range(args.steps):
        step(1)
        ev[k + 1].record()
    barrier()
    ms = ev[0].elapsed_time(ev[-1])
    per_step = sorted(ev[k].elapsed_time(ev[k + 1]) for k in range(args.steps))
    sweep_ms, sweep_n = lat.profile_read()
    lat.profile_sweeps(False)
    launches = lat.launch_count() - launches0
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = cells_global * args.steps / (ms * 1e-3) / 1e6
    stats = lat.stored_statistics()
    if ext:                             # this loop never evaluates the internal statistics: report the computed averages
        stats = (lat.compute_average_density() - 1.0, 2.0 * lat.compute_average_energy(), 0.0, 0)

    # ---- roofline of the sweep kernel ----------------------------------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    # a step sweeps a slab in one launch, or in three (rim planes first, then the interior) when the halo exchange
    # is overlapped: the roofline is taken over the sweep launches of one step
    sweep_ms_step = sweep_ms / args.steps
    achieved = cells_local * BYTES_PER_LU / (sweep_ms_step * 1e-3) / 1e9
    # DRAM bytes per launch from the ncu --set full capture of this very kernel (tools/ncu_traffic.py writes profiles/traffic.json
    # together with a hash of the kernel sources): a capture of other sources, size or workload is not this kernel's traffic
    traffic, traffic_note = None, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        if n == 512 and args.workload == "cavity":
            if tj.get("kernel_sources_sha256") == kernel_sources_hash():
                traffic = tj.get("dram_bytes_per_launch_512")
            else:
                traffic_note = "profiles/traffic.json was captured for other kernel sources (%s...): not reported" % str(tj.get("kernel_sources_sha256"))[:12]
    except Exception:
        pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "kernel": "sweep_kernel<%s,%d,%s> (even/odd AA halves averaged)" % ("double" if PREC == "f64" else "float", Q, wl["dyn"]),
                "peak_source": peak_src,
                "algorithmic_bytes_per_cell": BYTES_PER_LU, "cells_per_launch": cells_local,
                "launches_per_step": sweep_n / args.steps, "kernel_ms": sweep_ms_step, "kernel_share_of_step": sweep_ms / ms,
                "traffic_note": traffic_note, "peak_nominal_gbs": 8000.0, "frac_of_nominal": achieved / 8000.0}

    # ---- e2e: the same job from host buffers through the C-ABI ----------------------------------
    e2e = None
    if not args.no_e2e and cavity:
        nbytes = cells_local * Q * 8
        try:
            host = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
            pinned = True
        except Exception:
            host = torch.empty(nbytes, dtype=torch.uint8)
            pinned = False
        lat.download_ptr(host.data_ptr(), nbytes)       # a valid initial state to upload (outside the timed region)
        k2 = args.e2e_steps or args.steps
        barrier()
        launches_e0 = lat.launch_count()
        t0 = time.perf_counter()
        lat.upload_ptr(host.data_ptr(), nbytes)
        t1 = time.perf_counter()
        for it in range(k2):
            lat.set_boundary_velocity(lid, (U_MAX * min(1.0, (it + 1) / 50.0), 0, 0))
            lat.collide_and_stream(1)
            lat.stored_statistics()
        t2 = time.perf_counter()
        lat.download_ptr(host.data_ptr(), nbytes)
        barrier()
        t3 = time.perf_counter()
        dt = t3 - t0
        tt = torch.tensor([dt, t1 - t0, t2 - t1, t3 - t2], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt, h2d_s, loop_s, d2h_s = [float(v) for v in tt.tolist()]
        e2e = {"value": cells_global * k2 / dt / 1e6, "unit": "MLUPS",
               "h2d_bytes_per_step": (nbytes * world) / k2 + 48 * world,
               "d2h_bytes_per_step": (nbytes * world) / k2 + 32 * world,
               "steps": k2, "seconds": dt, "pinned_host": pinned, "cpu_binding": numa,
               # where the time goes (max over ranks each): the two PCIe transfers of the whole lattice, and the time loop whose
               # every step changes the lid velocity on the device (set_bc_values_kernel), sweeps, and reads the statistics back
               "h2d_s": h2d_s, "d2h_s": d2h_s, "loop_s": loop_s, "step_ms": loop_s / k2 * 1e3,
               "h2d_gbs_per_gpu": nbytes / h2d_s / 1e9, "d2h_gbs_per_gpu": nbytes / d2h_s / 1e9,
               "what": "upload initial populations + K x {setBoundaryVelocity(lid), collideAndStream, getStored statistics} + download final populations",
               "launches": lat.launch_count() - launches_e0}
        del host

    # ---- CPU baseline beside it (rank 0, N=1 only) ------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu and cavity:
        cpu = cpu_reference_run(128, 10, 1)

    if rank == 0:
        line = {"metric": "MLUPS", "value": value, "unit": "MLUPS", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps,
                "ms_per_step_p50": per_step[len(per_step) // 2], "ms_per_step_p95": per_step[min(len(per_step) - 1, int(0.95 * len(per_step)))],
                "ms_per_step_min": per_step[0], "ms_per_step_max": per_step[-1], "higher_is_better": True, "scaling": wl["scaling"], "vs_baseline": None,
                "dtype": PREC, "data": "synthetic",
                "config": {"workload": what,
                           "parallelism": "x-slabs, %d rank(s)" % world, "streaming": "in-place AA pattern",
                           "l2": "working set %.1f GB per GPU >> 126 MB L2, no flush needed" % (cells_local * BYTES_PER_LU / 2 / 1e9),
                           "omega": omega, "stored_avg_energy": 0.5 * stats[1], "stored_avg_density": stats[0] + 1.0},
                "mlups_per_gpu": value / world,
                "parity_check": parity, "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "e2e": e2e, "cpu_baseline": cpu}
        print(json.dumps(line), flush=True)
    lat.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=512, help="cavity edge per GPU (default: the configs[1] size)")
    ap.add_argument("--workload", default="cavity", choices=sorted(WORKLOADS), help="BASELINE.json config (default: configs[1], the bench line)")
    ap.add_argument("--e2e-steps", type=int, default=0)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        reference_arm(args)
    else:
        our_arm(args)


if __name__ == "__main__":
    main()
