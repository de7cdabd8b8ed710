"""Development fuzzer for the host-side media code (image codecs, OBJ/MTL importer): valid files made with PIL and by hand,
a few bytes flipped or the tail cut off, thousands of times, through Raylib_LoadImage / Raylib_LoadOBJModel.  Meant to be run
under AddressSanitizer + UBSan by tools/sanitize_host.sh; on its own it only shows that nothing crashes.

  python tools/fuzz_media.py [repository root]
"""
import io
import os
import sys

import numpy as np

R = sys.argv[1] if len(sys.argv) > 1 else os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R + '/software-raytracing_b200'); sys.path.insert(0, R + '/tests')
import pyraylib as rl  # noqa: E402
from PIL import Image  # noqa: E402
from test_cpu_host import OBJ_TEXT, MTL_TEXT, _write_png  # noqa: E402

prod = rl.Product(); prod.lib.Raylib_Initialize()
rng = np.random.default_rng(3)
TMP = os.environ.get("FUZZ_TMP", "/tmp")


def mutate_and_load(path, data, rounds, cut_every):
    total = ok = 0
    for it in range(rounds):
        d = bytearray(data)
        for _ in range(rng.integers(1, 6)):
            i = rng.integers(min(2, len(d) - 1), len(d)); d[i] = rng.integers(0, 256)
        if it % cut_every == 0:
            d = d[:rng.integers(1, len(d))]
        open(path, 'wb').write(d)
        img = prod.lib.Raylib_LoadImage(path.encode()); total += 1
        if img:
            ok += 1; prod.lib.Raylib_DestroyImage(img)
    return total, ok


pic = rng.integers(0, 256, (23, 31, 4), dtype=np.uint8)
srcs = {}
for fmt, ext, mode in (("PNG", "png", "RGBA"), ("PNG", "png", "P"), ("PNG", "png", "L"), ("BMP", "bmp", "RGB"), ("TGA", "tga", "RGBA"), ("PPM", "ppm", "RGB")):
    b = io.BytesIO(); im = Image.fromarray(pic, "RGBA").convert(mode)
    im.save(b, fmt, **({"compression": "tga_rle"} if fmt == "TGA" else {})); srcs[ext + mode] = (ext, b.getvalue())
for name, opts in (("jpg420", dict(quality=75, subsampling=2)), ("jpgprog", dict(quality=80, subsampling=0, progressive=True)),
                   ("jpgrst", dict(quality=70, subsampling=1, restart_marker_blocks=2)), ("jpgopt", dict(quality=90, subsampling=2, progressive=True, optimize=True))):
    b = io.BytesIO(); Image.fromarray(rng.integers(0, 256, (40, 56, 3), dtype=np.uint8)).save(b, 'JPEG', **opts); srcs[name] = ("jpg", b.getvalue())
srcs["hdr"] = ("hdr", b"#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n\n-Y 4 +X 9\n" + rng.integers(0, 256, (4 * 9 * 4,), dtype=np.uint8).tobytes())


def rle_line(w):
    out = bytes([2, 2, w >> 8, w & 255])
    for _ in range(4):
        n = 0
        while n < w:
            run = min(w - n, int(rng.integers(1, 100)))
            out += bytes([128 + run, int(rng.integers(0, 256))]) if rng.random() < 0.5 else bytes([run]) + rng.integers(0, 256, (run,), dtype=np.uint8).tobytes()
            n += run
    return out


srcs["hdrrle"] = ("hdr", b"#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n\n-Y 3 +X 40\n" + b"".join(rle_line(40) for _ in range(3)))
total = ok = 0
for name, (ext, data) in srcs.items():
    path = os.path.join(TMP, 'fz_media.' + ext)
    open(path, 'wb').write(data)
    img = prod.lib.Raylib_LoadImage(path.encode())
    assert img, name + ": the undamaged file must load"
    prod.lib.Raylib_DestroyImage(img)
    t, k = mutate_and_load(path, data, 1500, 6)
    total += t; ok += k
print('image fuzz: %d files, %d decoded, no crash' % (total, ok))

objdir = os.path.join(TMP, 'fz_obj'); os.makedirs(objdir, exist_ok=True)
_write_png(os.path.join(objdir, 'tiles.png'), np.full((4, 4, 4), 200, dtype=np.uint8))
tot = good = 0
tokens = [b"f", b"v", b"vt", b"vn", b"usemtl", b"mtllib", b"-1", b"1/2/3", b"//", b"1e999", b"nan", b"\n", b" ", b"0", b"99999999999", b"map_Kd", b"illum"]
for it in range(1500):
    o = bytearray(OBJ_TEXT.encode()); m = bytearray(MTL_TEXT.encode())
    for buf in (o, m):
        for _ in range(rng.integers(0, 4)):
            i = rng.integers(0, len(buf)); tok = tokens[rng.integers(0, len(tokens))]
            if rng.random() < 0.5:
                buf[i:i] = tok
            else:
                buf[i] = rng.integers(9, 127)
    open(os.path.join(objdir, 'scene.obj'), 'wb').write(o); open(os.path.join(objdir, 'scene.mtl'), 'wb').write(m)
    model = prod.lib.Raylib_LoadOBJModel(os.path.join(objdir, 'scene.obj').encode()); tot += 1
    if model:
        good += 1
        prod.lib.Raylib_FinalizeOBJModel(model)
        scene = prod.lib.Raylib_CreateScene(); prod.lib.Raylib_AddOBJModelToScene(scene, model); prod.lib.Raylib_FinalizeScene(scene)
        prod.flat_desc(scene)
        prod.lib.RaylibB200_ReleaseInspection(scene); prod.lib.Raylib_DestroyScene(scene); prod.lib.Raylib_UnloadOBJModel(model)
print('obj fuzz: %d files, %d loaded and flattened, no crash' % (tot, good))
