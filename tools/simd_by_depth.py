import ctypes as C, os, sys
sys.path.insert(0, "software-raytracing_b200")
import pyraylib as rl
prod = rl.Product(); prod.require_gpu(); prod.lib.Raylib_Initialize()
info = prod.create_demo(4, 0)
img = prod.lib.Raylib_CreateImage(info.settings.viewportWidth, info.settings.viewportHeight)
for depth in (1, 2, 8):
    s = info.settings.copy(samplesPerPixel=1, maxPathLength=depth)
    prod.lib.RaylibB200_SetCollectStats(1)
    prod.lib.Raylib_Render(C.byref(s), info.scene, info.camera, img)
    ss = prod.last_stats()
    prod.lib.RaylibB200_SetCollectStats(0)
    n = max(1, ss.statRays)
    sys.stderr.write("[simd] depth %d: rays %d nodes/ray %.2f box/ray %.2f tri/ray %.2f | node phase stepping %.1f alive %.1f, leaf testing %.1f, node iters/ray %.3f leaf iters/ray %.3f\n" % (
        depth, n, ss.nodeVisits / n, ss.boxTests / n, ss.triTests / n, 32.0 * ss.nodeStep / max(1, ss.nodeIters), 32.0 * ss.nodeAlive / max(1, ss.nodeIters),
        32.0 * ss.leafBusy / max(1, ss.leafIters), ss.nodeIters / 32.0 / n, ss.leafIters / 32.0 / n))
