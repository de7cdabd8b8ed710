"""Generates tests/golden/*.npz from the COMPILED REFERENCE (oracle/_ref, built from /root/reference by
oracle/Makefile).  Run in the build container:  python tools/make_golden.py
The fixtures pin the oracle restatement and the GPU path on hosts where the reference cannot be rebuilt."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "software-raytracing_b200"))
import pyraylib as rl
sys.path.insert(0, ROOT)
from oracle import bindings as ob

# (config, size parameter, viewport for primary hits, viewport + spp for radiance)
CASES = [(1, 0, (160, 90), (96, 54, 4)), (2, 0, (160, 90), (96, 54, 4)), (3, 48, (160, 90), (96, 54, 4)),
         (4, 24, (160, 90), (96, 54, 2)), (5, 40, (160, 90), (96, 54, 2)), (6, 0, (160, 90), (96, 54, 4))]

def main():
    ref = ob.Reference()
    out = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out, exist_ok=True)
    for cfg, size, (w, h), (rw, rh, spp) in CASES:
        info = ref.create_demo(cfg, size)
        ref.set_viewport(info, w, h)
        rank, t, rays, st = ref.primary_hits(info.settings, info.scene, info.camera, want_rays=True)
        assert st.walkVsHitMismatches == 0
        ref.set_viewport(info, rw, rh)
        s = info.settings.copy(samplesPerPixel=spp)
        img, rst = ref.render_deterministic(s, info.scene, info.camera)
        modes = {}
        for mode in (1, 2, 4, 5):
            m, _ = ref.render_deterministic(info.settings.copy(renderMode=mode), info.scene, info.camera)
            modes["mode%d" % mode] = m
        np.savez_compressed(os.path.join(out, "config%d.npz" % cfg),
                            config=cfg, size=size, primary_wh=np.array([w, h]), rank=rank, t=t, rays=rays,
                            ref_tests=np.array([st.boxTests, st.triTests, st.sphereTests, st.rays]),
                            radiance_whs=np.array([rw, rh, spp]), radiance=img, ray_queries=rst.rayQueries, seed=1337, **modes)
        print("config%d: leaves=%d hit=%.3f rays/sample=%.3f" % (cfg, st.numLeaves, (rank >= 0).mean(), rst.rayQueries / (rw * rh * spp)))
        ref.destroy_demo(info)

if __name__ == "__main__":
    main()
