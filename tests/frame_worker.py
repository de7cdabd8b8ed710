"""Second process of test_shared_frame_is_the_gather: maps the parent's frame through its CUDA IPC handle and renders
one shard of the demo scene straight into it (what every rank > 0 does in the one-process-per-GPU launch).

  python frame_worker.py <handle hex> <W> <H> <spp> <shardRank> <shardCount>
"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "software-raytracing_b200"))


def main():
    import pyraylib as rl
    handle = (C.c_ubyte * 64).from_buffer_copy(bytes.fromhex(sys.argv[1]))
    W, H, spp, rank, count = [int(x) for x in sys.argv[2:7]]
    prod = rl.Product()
    prod.require_gpu()
    prod.lib.Raylib_Initialize()
    info = prod.create_demo(6)
    prod.set_viewport(info, W, H)
    s = info.settings.copy(samplesPerPixel=spp)
    frame = prod.lib.RaylibB200_FrameOpen(handle)
    if not frame:
        print("FrameOpen failed: " + prod.last_error())
        return 2
    ok = prod.lib.RaylibB200_RenderShardToFrame(C.byref(s), info.scene, info.camera, rank, count, frame, None)
    if not ok:
        print("RenderShardToFrame failed: " + prod.last_error())
        return 3
    prod.lib.RaylibB200_FrameClose(frame)
    prod.destroy_demo(info)
    prod.lib.Raylib_Terminate()
    return 0


if __name__ == "__main__":
    sys.exit(main())
