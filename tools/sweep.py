"""Development sweep: build a workload once, render it under several environment settings, print Mrays/s.

  python tools/sweep.py <config id> <spp> NAME=v1,v2,... [NAME2=...]     (cartesian product)

The kernels read their tuning knobs (RAYLIB_B200_REFILL, RAYLIB_B200_WALK, ...) from the environment at
every render call, so one process can compare them on one uploaded scene.  Not part of the product.
"""
import ctypes as C
import itertools
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "software-raytracing_b200"))
import pyraylib as rl  # noqa: E402

libc = C.CDLL(None)


def setenv(k, v):
    os.environ[k] = v
    libc.setenv(k.encode(), v.encode(), 1)


def main():
    cfg, spp = int(sys.argv[1]), int(sys.argv[2])
    axes = []
    for a in sys.argv[3:]:
        name, vals = a.split("=")
        axes.append([(name, v) for v in vals.split(",")])
    prod = rl.Product()
    prod.require_gpu()
    prod.lib.Raylib_Initialize()
    info = prod.create_demo(cfg, 0)
    s = info.settings.copy(samplesPerPixel=spp) if spp > 0 else info.settings
    img = prod.lib.Raylib_CreateImage(s.viewportWidth, s.viewportHeight)
    # SWEEP_TIME_STAGES=0: no events around the k_extend launches (they keep small frames off the one-launch path)
    prod.lib.RaylibB200_SetTimeStages(0 if os.environ.get("SWEEP_TIME_STAGES") == "0" else 1)
    results = []
    for combo in itertools.product(*axes) if axes else [()]:
        for k, v in combo:
            setenv(k, v)
        prod.lib.RaylibB200_ReloadTuning()      # the library latches its knobs once per process
        best, wall = None, None
        for _ in range(4):
            t0 = time.perf_counter()
            prod.lib.Raylib_Render(C.byref(s), info.scene, info.camera, img)
            w = (time.perf_counter() - t0) * 1e3
            wall = w if wall is None else min(wall, w)
            st = prod.last_stats()
            if best is None or st.deviceMs < best[0]:
                best = (st.deviceMs, st.extendMs, st.rayQueries)
        row = {"cfg": cfg, "env": dict(combo), "device_ms": round(best[0], 3), "extend_ms": round(best[1], 3),
               "mrays_s": round(best[2] / best[0] / 1e3, 1), "wall_ms": round(wall, 3), "launches": st.kernelLaunches}
        results.append(row)
        print(json.dumps(row), flush=True)
    prod.lib.Raylib_DestroyImage(img)
    prod.destroy_demo(info)
    prod.lib.Raylib_Terminate()


if __name__ == "__main__":
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real, "w")
    main()
