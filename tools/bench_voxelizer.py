#!/usr/bin/env python
"""Measurement of the STL -> voxel-flag step on one B200: one JSON line with the device
throughput, its HBM roofline (algorithmic bytes: one int32 flag written per voxel + the bit planes) and the CPU
oracle (numpy port, bounded sample) beside it.  usage: python tools/bench_voxelizer.py [--res 1000]"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--res", type=int, default=1000)
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    import palabos_b200 as pb
    import voxel_oracle as vo
    soup = np.load(os.path.join(ROOT, "tests", "golden", "voxel_ball_40.npz"))["soup"]
    tri, dims, _, _ = pb.mesh_to_lattice(soup, args.res, 0, 4)
    n = int(np.prod(dims))
    for _ in range(3):
        pb.voxelize(tri, dims, 1)
    ms, wall = [], []
    for _ in range(args.reps):
        t0 = time.perf_counter()
        flags, k = pb.voxelize(tri, dims, 1, return_ms=True)
        wall.append(time.perf_counter() - t0)
        ms.append(k)
    kms = float(np.median(ms))
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    bytes_per_voxel = 4 + 3 * (1.0 / 8.0)          # flag written; bit planes: memset, toggles + prefix pass, border pass
    achieved = n * bytes_per_voxel / (kms * 1e-3) / 1e9
    tri_s, dims_s, _, _ = pb.mesh_to_lattice(soup, 96, 0, 4)
    t0 = time.perf_counter()
    want = vo.voxelize(tri_s, dims_s, 1)
    cpu_dt = time.perf_counter() - t0
    assert np.array_equal(pb.voxelize(tri_s, dims_s, 1), want)
    line = {"metric": "voxels/s", "value": n / (kms * 1e-3), "unit": "voxels/s", "n_gpus": 1, "higher_is_better": True,
            "config": {"workload": "voxelize ball.stl (2048 triangles) at resolution %d: %dx%dx%d int32 flags, borderWidth 1" % ((args.res,) + tuple(dims))},
            "kernel_ms": kms, "e2e": {"value": n / float(np.median(wall)), "unit": "voxels/s", "what": "C-ABI call from host triangles to host flags",
                                     "h2d_bytes": int(tri.nbytes), "d2h_bytes": int(flags.nbytes)},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "algorithmic_bytes_per_voxel": bytes_per_voxel},
            "cpu_baseline": {"value": int(np.prod(dims_s)) / cpu_dt, "unit": "voxels/s", "cores": 1, "kind": "port",
                             "sample": "numpy oracle, resolution 96 (%dx%dx%d), %.2f s; flags identical to the device's" % (tuple(dims_s) + (cpu_dt,))}}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
