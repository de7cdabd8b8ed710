"""Soak (development, GPU box): random small frames through the one-launch path and the kernel-per-stage path, bits compared.

  python tools/soak_fused.py [cases]
"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "software-raytracing_b200"))
import pyraylib as rl


def main():
    cases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    rng = np.random.default_rng(20240)
    prod = rl.Product(); prod.require_gpu(); prod.lib.Raylib_Initialize()
    bad = 0
    for i in range(cases):
        cfg = int(rng.choice([1, 2, 6, 7, 3]))
        size = 40 if cfg == 3 else 0            # a small displaced grid
        w, h = int(rng.integers(8, 400)), int(rng.integers(8, 300))
        spp = int(rng.integers(1, 12)); depth = int(rng.choice([0, 1, 2, 3, 5, 8, 17]))
        info = prod.create_demo(cfg, size)
        try:
            prod.set_viewport(info, w, h)
            s = info.settings.copy(samplesPerPixel=spp, maxPathLength=depth)
            frames = []
            for mode in (1, 2):
                prod.lib.RaylibB200_SetFusedPass(mode)
                frames.append(prod.render(s, info.scene, info.camera))
                st = prod.last_stats()
                frames.append(st.rayQueries)
            same = np.array_equal(frames[0].view(np.uint32), frames[2].view(np.uint32)) and frames[1] == frames[3]
            print("case %2d: cfg %d %3dx%-3d spp %2d depth %2d rays %9d launches(fused) %d -> %s" % (i, cfg, w, h, spp, depth, frames[1], st.kernelLaunches, "same" if same else "DIFFERENT"), flush=True)
            bad += 0 if same else 1
        finally:
            prod.destroy_demo(info)
    prod.lib.RaylibB200_SetFusedPass(0)
    prod.lib.Raylib_Terminate()
    print("soak_fused: %d cases, %d different" % (cases, bad))
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
