"""Per-launch profile of one phase on the GPU box (diagnostics, not a pytest file):
    python tests/gpu_plan_profile.py 3d:100 --phase=3 [--order=geo] [--top=40]
phase 0 factorization, 1 selinv, 2 forward sweep, 3 backward sweep; --lanes=16 profiles a lane sweep (phase 0 only)."""
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "gaussianmarkovrandomfields.jl_b200"), os.path.join(ROOT, "tests")]
from gmrf_b200 import spde  # noqa: E402
from gmrf_b200.backend import B200Backend  # noqa: E402
from gpu_perf import build_problem  # noqa: E402

spec = [a for a in sys.argv[1:] if not a.startswith("--")][0]
opt = dict(a[2:].split("=") for a in sys.argv[1:] if a.startswith("--") and "=" in a)
phases = [int(p) for p in opt.get("phase", "3").split(",")]
top = int(opt.get("top", "30"))
from gmrf_b200 import _lib  # noqa: E402
for k, v in opt.items():          # --set:key=value -> library option
    if k.startswith("set:"):
        _lib.set_option(k[4:], float(v))
Q, dims, width, _ = build_problem(spec)
ordering = spde.geometric_nd_perm(dims, leaf=64, width=width) if opt.get("order", "geo") == "geo" else "nd"
lanes = int(opt.get("lanes", "1"))
if lanes > 1:                      # lane handle: profile one sweep of `lanes` value sets (grid.y = lanes on every launch)
    import numpy as np
    _lib.set_option("lanes", lanes)
    b = B200Backend(Q, ordering=ordering, device=0, factorize=False)
    _lib.set_option("lanes", 1)
    Qc = Q.tocsc()
    for _ in range(2):
        ld, st = b.refactorize_lanes(np.tile(Qc.data, (lanes, 1)))
    print(f"lanes {lanes}: sweep {b.timings()['factor_ms']:.3f} ms on the device = {b.timings()['factor_ms'] / lanes:.3f} ms per value set; "
          f"logdets agree: {bool(np.all(ld == ld[0]))}")
else:
    b = B200Backend(Q, ordering=ordering, device=0)
for phase in phases:
    b.profile_plan(phase)                       # warm-up
    rows = b.profile_plan(phase)
    tot = sum(r[2] for r in rows)
    print(f"== {spec} phase {phase}: {len(rows)} launches, {tot:.3f} ms summed")
    agg = collections.OrderedDict()
    for k, g, t in rows:
        c, tt, gg = agg.get(k, (0, 0.0, 0))
        agg[k] = (c + 1, tt + t, gg + g)
    for k, (c, tt, gg) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"   {k:12s} {c:6d} launches {tt:10.3f} ms ({100 * tt / tot:5.1f}%)  avg {1e3 * tt / c:9.1f} us  avg grid {gg / c:9.1f}")
    print("   slowest launches (index kind grid ms):")
    for i in sorted(range(len(rows)), key=lambda i: -rows[i][2])[:top]:
        print(f"     {i:6d} {rows[i][0]:12s} {rows[i][1]:8d} {rows[i][2]:9.4f}")
    if "dump" in opt:
        a, b_ = (int(v) for v in opt["dump"].split(":"))
        print(f"   launches {a}..{b_} (index kind grid ms):")
        for i in range(a, min(b_, len(rows))):
            print(f"     {i:6d} {rows[i][0]:12s} {rows[i][1]:8d} {rows[i][2]:9.4f}")
    if phase in (0, 1):
        # GEMM launches by contraction length: time, algorithmic flops, TFLOP/s
        fl, km, nt = b.plan_launch_info(phase)
        if len(fl) == len(rows):
            classes = [(0, 64), (65, 128), (129, 256), (257, 512), (513, 1024), (1025, 4096), (4097, 10 ** 9)]
            print("   GEMM launches by longest contraction k (launches, ms, Gflop, TFLOP/s, avg tiles):")
            for lo, hi in classes:
                sel = [i for i in range(len(rows)) if rows[i][0].startswith("gemm") and lo <= km[i] <= hi]
                if not sel:
                    continue
                tms = sum(rows[i][2] for i in sel); f = sum(fl[i] for i in sel)
                print(f"     k in [{lo:5d}, {hi if hi < 10 ** 9 else 'inf':>5}] {len(sel):6d} {tms:10.3f} {f / 1e9:12.2f} {f / (tms * 1e-3) / 1e12 if tms > 0 else 0:8.2f} {sum(rows[i][1] for i in sel) / len(sel):10.1f}")
    # cumulative time by position (coarse timeline in 20 buckets)
    nb = 20
    step = max(1, len(rows) // nb)
    line = []
    for a in range(0, len(rows), step):
        line.append(f"{sum(r[2] for r in rows[a:a + step]):.1f}")
    print("   timeline (ms per 1/20 of the launch list):", " ".join(line))
