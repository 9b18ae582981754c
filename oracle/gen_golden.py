#!/usr/bin/env python
"""Generate tests/golden/*.npz from the UNMODIFIED reference.  TEST INFRASTRUCTURE ONLY.

Runs every case of tests/cases.py through oracle/_ref/ref_driver{,_f32} (the real Palabos
code, compiled by `make -C oracle ref` from /root/reference/src) and stores, per case:

  stats                  rows of [iter, stored avg density, stored avg energy, stored max velocity,
                         computeAverageEnergy, computeAverageDensity]
  <dump>_sha256          sha256 of the raw dump bytes ([nx][ny][nz][q] of T; rho_u dumps [..][4])
  <dump>_probe_idx/_probe   a deterministic sample of cells and their values
  <dump>_full            the whole dump, only for the cases marked "full" that are tiny

This can only run where /root/reference exists (this container); the fixtures travel.
Usage: python oracle/gen_golden.py [case ...]      (no argument: every case of CASES; the LARGE_CASES only by name)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from cases import CASES, LARGE_CASES, probe_index  # noqa: E402
from scenario import digest, run_reference  # noqa: E402

FULL_LIMIT = 800  # cells: dumps of lattices up to this size are stored whole


def main():
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    names = sys.argv[1:] or list(CASES)
    for name in names:
        build, how = CASES[name] if name in CASES else LARGE_CASES[name]
        sc = build()
        ref = run_reference(sc)
        rec = {"stats": ref["stats"]}
        for key, arr in ref.items():
            if key in ("stats", "time"):
                continue
            flat = arr.reshape(-1, arr.shape[-1])
            idx = probe_index(arr.shape)
            rec[key + "_sha256"] = np.array(digest(arr))
            rec[key + "_probe_idx"] = idx
            rec[key + "_probe"] = flat[idx]
            if how == "full" and flat.shape[0] <= FULL_LIMIT:
                rec[key + "_full"] = arr
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), **rec)
        print("golden", name, "stats rows", len(ref["stats"]), flush=True)


if __name__ == "__main__":
    main()
