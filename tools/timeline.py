"""Development: when each k_extend launch of each pipe ran (CUDA events), to see how the two passes in flight overlap.

  python tools/timeline.py <config id> <spp>
"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "software-raytracing_b200"))
import pyraylib as rl  # noqa: E402

libc = C.CDLL(None)


def main():
    cfg, spp = int(sys.argv[1]), int(sys.argv[2])
    prod = rl.Product()
    prod.require_gpu()
    prod.lib.Raylib_Initialize()
    info = prod.create_demo(cfg, 0)
    s = info.settings.copy(samplesPerPixel=spp) if spp > 0 else info.settings
    img = prod.lib.Raylib_CreateImage(s.viewportWidth, s.viewportHeight)
    prod.lib.RaylibB200_SetTimeStages(1)
    prod.lib.Raylib_Render(C.byref(s), info.scene, info.camera, img)        # warm-up
    libc.setenv(b"RAYLIB_B200_DUMP_TIMELINE", b"1", 1)
    prod.lib.Raylib_Render(C.byref(s), info.scene, info.camera, img)
    st = prod.last_stats()
    sys.stderr.write("[frame] %.3f ms on device, k_extend %.3f ms, %.1f Mrays/s\n" % (st.deviceMs, st.extendMs, st.rayQueries / st.deviceMs / 1e3))
    prod.lib.Raylib_DestroyImage(img)
    prod.destroy_demo(info)
    prod.lib.Raylib_Terminate()


if __name__ == "__main__":
    main()
