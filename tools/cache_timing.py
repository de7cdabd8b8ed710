"""Development: what the flattened-scene cache saves on a large scene (host only, no GPU needed).

  python tools/cache_timing.py <config id> [size]
"""
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "software-raytracing_b200"))
import pyraylib as rl  # noqa: E402


def main():
    cfg = int(sys.argv[1]); size = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    prod = rl.Product()
    prod.lib.Raylib_Initialize()
    t0 = time.time(); info = prod.create_demo(cfg, size); t1 = time.time()
    prod.flat_desc(info.scene); t2 = time.time()
    path = os.path.join(tempfile.gettempdir(), "cache_timing_%d.rtflat" % cfg)
    assert prod.lib.RaylibB200_SaveFlattenedScene(info.scene, path.encode()) == 1, prod.last_error()
    t3 = time.time()
    loaded = prod.lib.RaylibB200_LoadFlattenedScene(path.encode()); t4 = time.time()
    assert loaded, prod.last_error()
    print("config %d: build object graph + reference BVHs %.2f s, flatten (SAH, 4-wide, quantize) %.2f s, save %.2f s (%.1f MB), load %.2f s"
          % (cfg, t1 - t0, t2 - t1, t3 - t2, os.path.getsize(path) / 1e6, t4 - t3))
    prod.lib.Raylib_DestroyScene(loaded)
    prod.destroy_demo(info)
    os.remove(path)


if __name__ == "__main__":
    main()
