"""development tool: what each kind of boundary face costs the step at 512^3 (D3Q19 BGK fp64)
usage: python tools/face_cost.py [n [cases]]      (PLB_B200_NO_HANDOFF=1: the plane pass writes every cell back into the lattice;
FC_BLOCK=1: the "xyz" case uses setVelocityConditionOnBlockBoundaries, i.e. with edges and corners)"""
import os, sys, json
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import palabos_b200 as pb

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
cases = {"none": [], "x": [(0, -1), (0, 1)], "y": [(1, -1), (1, 1)], "z": [(2, -1), (2, 1)], "zlo": [(2, -1)], "xyz": [(0, -1), (0, 1), (1, -1), (1, 1), (2, -1), (2, 1)]}
sel = sys.argv[2].split(",") if len(sys.argv) > 2 else list(cases)
U = float(os.environ.get("FC_U", "0.01")); OMEGA = float(os.environ.get("FC_OMEGA", "1.6")); BLOCK = os.environ.get("FC_BLOCK", "0") == "1"
for name in sel:
    lat = pb.Lattice(19, "f64", n, n, n, "BGK", OMEGA)
    if BLOCK and name == "xyz":
        lat.set_velocity_condition_on_block_boundaries("local", lat.box, lat.box)
    else:
      for d, o in cases[name]:
          b = [1, n - 2, 1, n - 2, 1, n - 2]
          b[2 * d] = b[2 * d + 1] = 0 if o < 0 else n - 1
          lat.add_velocity_boundary("local", d, o, tuple(b))
    lat.set_boundary_velocity(lat.box, (U, 0, 0))
    lat.initialize_at_equilibrium(lat.box, 1.0, (U, 0, 0))
    lat.initialize()
    lat.collide_and_stream(6)
    torch.cuda.synchronize()
    K = 20
    lat.profile_sweeps(True)
    l0 = lat.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); lat.collide_and_stream(K); e1.record(); torch.cuda.synchronize()
    sw, sn = lat.profile_read()
    print("%-5s no_handoff=%s  ms/step %.3f  sweep %.3f ms/step  launches/step %.1f" % (name, os.environ.get("PLB_B200_NO_HANDOFF", "0"), e0.elapsed_time(e1) / K, sw / K, (lat.launch_count() - l0) / K), flush=True)
    lat.close()
