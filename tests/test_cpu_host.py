"""CPU-only: the C-ABI boundary and the host logic (no compute calls that need a GPU)."""
import ctypes as C
import os
import re
import subprocess
import sys
import numpy as np
import pytest
from conftest import bits

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    return sorted(set(re.findall(r"RAYLIB_API[^;{(]*?\b((?:Raylib|RaylibB200)_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol(rl):
    lib = C.CDLL(rl.PRODUCT_LIB, mode=os.RTLD_LOCAL | os.RTLD_NOW)
    ref_api = declared_symbols("raylib.h")
    assert len(ref_api) == 33, ref_api                      # the reference's 33 entry points (raylib/raylib.h:23-149)
    for name in ref_api + declared_symbols("raylib_b200.h") + ["CHECK_IMPL", "CHECKF_IMPL"]:
        assert hasattr(lib, name), "missing export: " + name
    # the thin device ABI (include/rt_device_abi.h): every declared entry point is exported as well
    text = open(os.path.join(ROOT, "include", "rt_device_abi.h")).read()
    device_api = sorted(set(re.findall(r"RT_DEVICE_API[^;(]*?\b(rt_\w+)\s*\(", text)))
    assert len(device_api) == 28, device_api
    for name in device_api:
        assert hasattr(lib, name), "missing export: " + name
    assert lib.rt_shard_tile_capacity(33, 17, 1) == 6                 # 3 x 2 tiles of 16x16 (no GPU needed)
    # every bound name in the python tables is declared in a header
    declared = set(ref_api) | set(declared_symbols("raylib_b200.h")) | {"RaylibB200_SeedHostRandom"}
    for name in list(rl.RAYLIB_C_API) + list(rl.B200_C_API):
        assert name in declared, name + " is bound but not declared in include/*.h"


def test_handles_and_error_conventions(prod):
    lib = prod.lib
    assert lib.Raylib_GetRenderModeString(2) == b"SurfaceNormal"       # raylib.cc:298-312
    assert lib.Raylib_GetRenderModeString(7) is None
    assert lib.Raylib_IsDenoiserSupported() == 0
    cam = lib.Raylib_CreateCamera()
    assert lib.Raylib_DestroyCamera(cam) == 1 and lib.Raylib_DestroyCamera(cam) == 0     # raylib.cc:170-179
    img = lib.Raylib_CreateImage(4, 3)
    out = np.empty((3, 4, 3), dtype=np.float32)
    lib.Raylib_DumpImageData(img, out)
    assert (out == 0).all()
    assert lib.Raylib_DestroyImage(img) == 1 and lib.Raylib_DestroyImage(img) == 0
    scene = lib.Raylib_CreateScene()
    assert lib.Raylib_DestroyScene(scene) == 1 and lib.Raylib_DestroyScene(scene) == 0
    assert lib.Raylib_LoadOBJModel(b"/nonexistent.obj") == 0          # NULL on failure, raylib.cc:56-69
    assert lib.Raylib_WriteImageToDisk(0, b"x.bmp", 0) == 0


def test_render_without_gpu_fails_loudly(prod):
    if prod.device_count() > 0:
        pytest.skip("a GPU is visible here")
    info = prod.create_demo(6)
    try:
        with pytest.raises(RuntimeError, match="GPU only"):
            prod.render(info.settings, info.scene, info.camera)
        img = prod.lib.Raylib_CreateImage(8, 8)
        s = info.settings.copy(viewportWidth=8, viewportHeight=8)
        prod.lib.Raylib_Render(C.byref(s), info.scene, info.camera, img)     # must not crash, must not render
        assert "no CUDA device" in prod.last_error()
        assert (prod.dump_image(img, 8, 8) == 0).all()
        prod.lib.Raylib_DestroyImage(img)
    finally:
        prod.destroy_demo(info)


def test_flatten_is_deterministic_and_consistent(prod):
    snaps = []
    for _ in range(2):
        info = prod.create_demo(4, 12)
        d = prod.flat_desc(info.scene).contents
        nodes = np.ctypeslib.as_array(C.cast(d.nodes, C.POINTER(C.c_uint32)), shape=(d.numNodes * 16,)).copy()
        hot = np.ctypeslib.as_array(C.cast(d.triHot, C.POINTER(C.c_uint32)), shape=(d.numTris * 16,)).copy()
        rank = np.ctypeslib.as_array(C.cast(d.triRank, C.POINTER(C.c_uint32)), shape=(d.numTris,)).copy()
        snaps.append((nodes, hot, rank, d.numLeaves, d.maxStackDepth, d.rootRef))
        assert d.numTris == 12 * 1280 + 2
        assert d.numLeaves == d.numTris + d.numSpheres + d.numCubes
        assert sorted(rank.tolist()) == list(range(d.numTris)), "ranks must be a permutation of the leaves"
        assert (np.diff(rank) > 0).all(), "triangles are stored in in-order (tie-break) order"
        assert d.materialTypeMask == 0b100111                              # lambertian, metal, dielectric, microfacet
        prod.destroy_demo(info)
    for a, b in zip(snaps[0], snaps[1]):
        assert np.array_equal(a, b), "BVH build / flatten must be reproducible"


def test_unsupported_graphs_are_reported(prod):
    lib = prod.lib
    scene = lib.Raylib_CreateScene()
    lib.Raylib_FinalizeScene(scene)                       # empty scene
    assert not lib.RaylibB200_FlattenForInspection(scene)
    assert "no elements" in prod.last_error()
    lib.Raylib_DestroyScene(scene)


@pytest.mark.parametrize("wh", [(1, 1), (16, 16), (17, 33), (640, 360), (1920, 1080)])
@pytest.mark.parametrize("shards", [1, 2, 3, 8])
def test_shard_maps_partition_the_image(prod, wh, shards):
    w, h = wh
    cap = int(prod.lib.RaylibB200_ShardPixelCapacity(w, h, shards))
    assert cap % 256 == 0
    seen = np.zeros(w * h, dtype=np.int32)
    for r in range(shards):
        m = np.empty(cap, dtype=np.int64)
        assert prod.lib.RaylibB200_ShardPixelMap(w, h, r, shards, m)
        valid = m[m >= 0]
        seen[valid] += 1
    assert (seen == 1).all(), "every pixel belongs to exactly one shard slot"


def test_host_assemble_roundtrip(prod):
    w, h, shards = 50, 37, 3
    cap = int(prod.lib.RaylibB200_ShardPixelCapacity(w, h, shards))
    truth = np.arange(w * h * 4, dtype=np.float32).reshape(h * w, 4)
    slabs = np.zeros((shards, cap, 4), dtype=np.float32)
    for r in range(shards):
        m = np.empty(cap, dtype=np.int64)
        prod.lib.RaylibB200_ShardPixelMap(w, h, r, shards, m)
        slabs[r, m >= 0] = truth[m[m >= 0]]
    out = np.zeros((h * w, 4), dtype=np.float32)
    assert prod.lib.RaylibB200_AssembleShardsHost(slabs.reshape(-1), shards, w, h, out.reshape(-1))
    assert np.array_equal(out, truth)


GLOO_WORKER = r'''
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.join(%(root)r, "software-raytracing_b200"))
import pyraylib as rl
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
prod = rl.Product()
w, h = 70, 45
cap = int(prod.lib.RaylibB200_ShardPixelCapacity(w, h, world))
m = np.empty(cap, dtype=np.int64)
prod.lib.RaylibB200_ShardPixelMap(w, h, rank, world, m)
# stand-in for a rendered shard: pixel value = f(global pixel index), exactly what a rank would own
shard = torch.zeros(cap, 4)
idx = torch.from_numpy(m)
shard[idx >= 0] = torch.stack([idx[idx >= 0].float() * k for k in (1, 2, 3, 4)], dim=1)
gathered = [torch.zeros(cap, 4) for _ in range(world)] if rank == 0 else None
dist.gather(shard, gathered, dst=0)
if rank == 0:
    flat = torch.cat(gathered).numpy().astype(np.float32)
    out = np.zeros((h * w, 4), dtype=np.float32)
    assert prod.lib.RaylibB200_AssembleShardsHost(flat.reshape(-1), world, w, h, out.reshape(-1))
    want = np.arange(w * h, dtype=np.float32)[:, None] * np.array([1, 2, 3, 4], dtype=np.float32)
    assert np.array_equal(out, want), "assembled frame differs"
    print("GLOO_OK")
dist.destroy_process_group()
'''


def test_two_rank_gather_path_gloo(tmp_path):
    """The N>1 host logic: per-rank tile ownership, one gather to rank 0, de-interleave. gloo, world_size 2."""
    script = tmp_path / "worker.py"
    script.write_text(GLOO_WORKER % {"root": ROOT})
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29531", str(script)],
                         capture_output=True, text=True, timeout=280, env=env)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert "GLOO_OK" in res.stdout


# ---- host-side media: OBJ import and image codecs (reference loader/obj_loader.cc, render/image.cc) ----

def _write_png(path, rgba):
    """Minimal PNG writer (zlib) so the decoder is tested against an independent encoder."""
    import struct, zlib
    h, w, _ = rgba.shape
    raw = b"".join(b"\x00" + rgba[y].tobytes() for y in range(h))
    def chunk(t, d):
        return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d) & 0xffffffff)
    open(path, "wb").write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 6, 0, 0, 0))
                           + chunk(b"IDAT", zlib.compress(raw)) + chunk(b"IEND", b""))


def test_image_codecs_roundtrip(prod, tmp_path):
    import ctypes as C
    rng = np.random.default_rng(5)
    rgba = rng.integers(0, 256, size=(7, 13, 4), dtype=np.uint8)
    png = str(tmp_path / "t.png")
    _write_png(png, rgba)
    img = prod.lib.Raylib_LoadImage(png.encode())
    assert img, "PNG written by an independent encoder must load"
    px = prod.dump_image(img, 13, 7)
    assert np.array_equal(px, rgba[:, :, :3].astype(np.float32) / np.float32(255.0)), "row 0 = top, byte / 255"
    # write it back as BMP and PNG through the reference entry point, reload both
    for ftype, name in ((0, "o.bmp"), (2, "o.png")):
        out = str(tmp_path / name)
        assert prod.lib.Raylib_WriteImageToDisk(img, out.encode(), ftype) == 1
        back = prod.lib.Raylib_LoadImage(out.encode())
        assert back
        assert np.array_equal(prod.dump_image(back, 13, 7), px)
        assert prod.lib.Raylib_DestroyImage(back) == 1
    assert prod.lib.Raylib_WriteImageToDisk(img, str(tmp_path / "o.xyz").encode(), 7) == 0      # unknown file type: refused, not faked
    assert prod.lib.Raylib_DestroyImage(img) == 1
    # Radiance HDR (flat RGBE) and PFM
    hdr = str(tmp_path / "t.hdr")
    rgbe = np.array([[[128, 64, 32, 129], [0, 0, 0, 0]], [[255, 255, 255, 128], [1, 2, 3, 136]]], dtype=np.uint8)
    open(hdr, "wb").write(b"#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n\n-Y 2 +X 2\n" + rgbe.tobytes())
    img = prod.lib.Raylib_LoadImage(hdr.encode())
    assert img
    got = prod.dump_image(img, 2, 2)
    scale = np.where(rgbe[..., 3:] > 0, np.ldexp(1.0, rgbe[..., 3:].astype(np.int32) - 136), 0.0)
    assert np.allclose(got, rgbe[..., :3] * scale)
    prod.lib.Raylib_DestroyImage(img)
    assert prod.lib.Raylib_LoadImage(str(tmp_path / "missing.png").encode()) == 0


def _jpeg_test_picture(w, h, seed):
    """Smooth colour gradients + a few hard edges + a little noise: exercises DC prediction, long zero runs and clamping."""
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:h, 0:w].astype(np.float64)
    img = np.stack([127 + 120 * np.sin(x / 7.0 + y / 11.0), 127 + 120 * np.cos(x / 5.0 - y / 13.0), (x * 3 + y * 5) % 256], axis=-1)
    img[h // 3: h // 2, w // 4: w // 2] = (255, 0, 255)
    img[: h // 5, -w // 3:] = (0, 255, 10)
    img += rng.normal(0, 6, img.shape)
    return np.clip(img, 0, 255).astype(np.uint8)


JPEG_FLAVOURS = [
    # name, size, save options
    ("baseline_444", (67, 45), dict(quality=90, subsampling=0)),
    ("baseline_422", (67, 45), dict(quality=85, subsampling=1)),
    ("baseline_420", (67, 45), dict(quality=75, subsampling=2)),
    ("baseline_420_even", (64, 48), dict(quality=50, subsampling=2)),
    ("optimized_420", (130, 71), dict(quality=95, subsampling=2, optimize=True)),
    ("progressive_444", (67, 45), dict(quality=80, subsampling=0, progressive=True)),
    ("progressive_420", (131, 77), dict(quality=92, subsampling=2, progressive=True)),
    ("restart_420", (131, 77), dict(quality=70, subsampling=2, restart_marker_blocks=3)),
    ("tiny_420", (3, 2), dict(quality=75, subsampling=2)),
    ("narrow_422", (5, 33), dict(quality=75, subsampling=1)),
    ("low_quality", (67, 45), dict(quality=5, subsampling=2)),
]


def test_jpeg_decoder_matches_libjpeg(prod, tmp_path):
    """Raylib_LoadImage on JPEG files (reference: FreeImage -> libjpeg, render/image.cc:152-230): every flavour an OBJ
    texture comes in decodes to the bytes libjpeg-turbo (inside PIL) produces -- integer IDCT, fancy upsampling and the
    fixed-point colour conversion are restated exactly (csrc/host/jpeg_codec.cc)."""
    Image = pytest.importorskip("PIL.Image")
    for name, (w, h), options in JPEG_FLAVOURS:
        src = _jpeg_test_picture(w, h, len(name))
        path = str(tmp_path / (name + ".jpg"))
        try:
            Image.fromarray(src).save(path, "JPEG", **options)
        except TypeError:
            continue          # an option this PIL does not know
        want = np.asarray(Image.open(path).convert("RGB"))
        img = prod.lib.Raylib_LoadImage(path.encode())
        assert img, name
        got = prod.dump_image(img, w, h)
        assert np.array_equal(got, want.astype(np.float32) / np.float32(255.0)), \
            "%s: %d of %d samples differ from libjpeg" % (name, int((got != want.astype(np.float32) / np.float32(255.0)).sum()), want.size)
        prod.lib.Raylib_DestroyImage(img)
    # grey
    grey = _jpeg_test_picture(50, 37, 3)[..., 0]
    path = str(tmp_path / "grey.jpg")
    Image.fromarray(grey, "L").save(path, "JPEG", quality=80)
    img = prod.lib.Raylib_LoadImage(path.encode())
    assert img
    want = np.asarray(Image.open(path).convert("L")).astype(np.float32) / np.float32(255.0)
    assert np.array_equal(prod.dump_image(img, 50, 37), np.repeat(want[..., None], 3, axis=-1))
    prod.lib.Raylib_DestroyImage(img)
    # truncated and foreign files are refused, not crashed on
    data = open(str(tmp_path / "baseline_420.jpg"), "rb").read()
    open(str(tmp_path / "cut.jpg"), "wb").write(data[:40])
    assert prod.lib.Raylib_LoadImage(str(tmp_path / "cut.jpg").encode()) == 0
    open(str(tmp_path / "noise.jpg"), "wb").write(bytes(range(256)) * 4)
    assert prod.lib.Raylib_LoadImage(str(tmp_path / "noise.jpg").encode()) == 0
    # a scan that ends early still yields a picture of the right size (what libjpeg does too)
    open(str(tmp_path / "short.jpg"), "wb").write(data[: len(data) * 2 // 3])
    img = prod.lib.Raylib_LoadImage(str(tmp_path / "short.jpg").encode())
    assert img
    assert prod.dump_image(img, 67, 45).shape == (45, 67, 3)
    prod.lib.Raylib_DestroyImage(img)


def test_jpeg_golden_fixtures(prod):
    """The same comparison without PIL at test time: files and expected bytes committed by tools/make_jpeg_golden.py."""
    gold = os.path.join(ROOT, "tests", "golden")
    names = sorted(f[:-4] for f in os.listdir(gold) if f.startswith("jpeg_") and f.endswith(".jpg"))
    assert len(names) >= 3
    for name in names:
        want = np.load(os.path.join(gold, name + ".npy"))
        h, w, _ = want.shape
        img = prod.lib.Raylib_LoadImage(os.path.join(gold, name + ".jpg").encode())
        assert img, name
        assert np.array_equal(prod.dump_image(img, w, h), want.astype(np.float32) / np.float32(255.0)), name
        prod.lib.Raylib_DestroyImage(img)


def test_jpeg_writer(prod, tmp_path):
    """Raylib_WriteImageToDisk(..., RAYLIB_IMAGEFILETYPE_Jpg) -- what the reference client calls for every result
    (src/main.cc:478-510) -- writes a baseline JFIF file that this library and libjpeg read back alike and that is close to
    the source picture."""
    src = _jpeg_test_picture(83, 59, 11)
    src[:, :, 2] = src[:, :, 0] // 2 + 60          # keep the chroma smooth enough for 4:2:0
    png = str(tmp_path / "src.png")
    _write_png(png, np.dstack([src, np.full(src.shape[:2], 255, np.uint8)]))
    img = prod.lib.Raylib_LoadImage(png.encode())
    out = str(tmp_path / "o.jpg")
    assert prod.lib.Raylib_WriteImageToDisk(img, out.encode(), 1) == 1
    data = open(out, "rb").read()
    assert data[:4] == b"\xff\xd8\xff\xe0" and data[6:11] == b"JFIF\0" and data[-2:] == b"\xff\xd9"
    back = prod.lib.Raylib_LoadImage(out.encode())
    assert back
    got = prod.dump_image(back, 83, 59)
    err = got.astype(np.float64) * 255.0 - src
    psnr = 10.0 * np.log10(255.0 ** 2 / np.mean(err ** 2))
    assert psnr > 25.0, psnr          # a noisy picture with hard colour edges at quality 75, 4:2:0
    try:
        from PIL import Image
    except ImportError:
        Image = None
    if Image is not None:
        want = np.asarray(Image.open(out).convert("RGB")).astype(np.float32) / np.float32(255.0)
        assert np.array_equal(got, want), "libjpeg reads the written file to the same bytes"
        # same quality and size class as libjpeg's own encoder at the settings FreeImage uses (quality 75, 4:2:0)
        import io
        buf = io.BytesIO()
        Image.fromarray(src).save(buf, "JPEG", quality=75, subsampling=2)
        theirs = np.asarray(Image.open(io.BytesIO(buf.getvalue())).convert("RGB")).astype(np.float64)
        their_psnr = 10.0 * np.log10(255.0 ** 2 / np.mean((theirs - src) ** 2))
        assert psnr > their_psnr - 0.3, (psnr, their_psnr)
        assert len(data) < 1.1 * len(buf.getvalue()) + 64, (len(data), len(buf.getvalue()))
    prod.lib.Raylib_DestroyImage(back)
    prod.lib.Raylib_DestroyImage(img)


OBJ_TEXT = """mtllib scene.mtl
o floor
v -1 0 -1
v 1 0 -1
v 1 0 1
v -1 0 1
vt 0 0
vt 1 0
vt 1 1
vt 0 1
vn 0 1 0
usemtl tiled
f 1/1/1 2/2/1 3/3/1 4/4/1
o glass
v 0 0.2 0
v 0.5 0.2 0
v 0 0.7 0
usemtl glass
f 5 6 7
g mirror_part
usemtl chrome
f -3 -1 -2
f 5 6 7
"""
MTL_TEXT = """newmtl tiled
Kd 1 0.5 0.25
Ks 0.5 0.5 0.5
Ns 96
map_Kd -s 1 1 1 tiles.png
newmtl glass
Kd 0 0 0
Tf 0.9 0.9 1.0
Ni 1.5
illum 4
newmtl chrome
Kd 0.8 0.8 0.8
illum 3
"""


def test_obj_model_import(prod, tmp_path):
    import ctypes as C
    (tmp_path / "scene.obj").write_text(OBJ_TEXT)
    (tmp_path / "scene.mtl").write_text(MTL_TEXT)
    _write_png(str(tmp_path / "tiles.png"), np.full((4, 4, 4), 200, dtype=np.uint8))
    model = prod.lib.Raylib_LoadOBJModel(str(tmp_path / "scene.obj").encode())
    assert model, "OBJ + MTL + PNG must load"
    prod.lib.Raylib_TransformOBJModel(model, 0.0, 1.0, 0.0, 0.0, 0.0, 0.0, 2.0, 2.0, 2.0)
    prod.lib.Raylib_FinalizeOBJModel(model)
    scene = prod.lib.Raylib_CreateScene()
    prod.lib.Raylib_AddOBJModelToScene(scene, model)
    prod.lib.Raylib_FinalizeScene(scene)
    d = prod.flat_desc(scene).contents
    assert d.numTris == 2 + 1 + 2, "the quad is triangulated, three shapes"
    # tiled -> MicrofacetMaterial with an albedo texture, glass -> Dielectric (illum 4, Kd 0), chrome -> Mirror (illum 3)
    assert d.materialTypeMask == (1 << 5) | (1 << 2) | (1 << 3)
    assert d.numTextures == 1 and d.numTexels == 16
    assert d.flags & 1, "an albedo texture switches the alpha cut-out test on"
    lo, hi = np.array(d.rootMin[:]), np.array(d.rootMax[:])
    assert np.all(lo <= np.array([-2.0, 1.0, -2.0]) + 1e-3) and np.all(hi >= np.array([2.0, 1.0, 2.0]) - 1e-3), "scale 2, translate y+1"
    assert hi[1] >= 2.4 - 1e-3, "the glass triangle top (0.7 * 2 + 1)"
    prod.lib.RaylibB200_ReleaseInspection(scene)
    assert prod.lib.Raylib_DestroyScene(scene) == 1
    assert prod.lib.Raylib_UnloadOBJModel(model) == 1
    assert prod.lib.Raylib_UnloadOBJModel(model) == 0
    assert prod.lib.Raylib_LoadOBJModel(str(tmp_path / "nope.obj").encode()) == 0


def test_obj_model_with_jpeg_texture(prod, tmp_path):
    """The usual shape of a downloaded OBJ scene: map_Kd names a .jpg.  The texels that reach the flattened scene are the
    bytes libjpeg decodes (tests/golden/jpeg_baseline_420.npy), not a constant-colour fallback."""
    import shutil
    gold = os.path.join(ROOT, "tests", "golden")
    (tmp_path / "scene.obj").write_text(OBJ_TEXT)
    (tmp_path / "scene.mtl").write_text(MTL_TEXT.replace("tiles.png", "tiles.jpg"))
    shutil.copy(os.path.join(gold, "jpeg_baseline_420.jpg"), str(tmp_path / "tiles.jpg"))
    want = np.load(os.path.join(gold, "jpeg_baseline_420.npy"))
    model = prod.lib.Raylib_LoadOBJModel(str(tmp_path / "scene.obj").encode())
    assert model
    prod.lib.Raylib_FinalizeOBJModel(model)
    scene = prod.lib.Raylib_CreateScene()
    prod.lib.Raylib_AddOBJModelToScene(scene, model)
    prod.lib.Raylib_FinalizeScene(scene)
    d = prod.flat_desc(scene).contents
    assert d.numTextures == 1 and d.numTexels == want.shape[0] * want.shape[1]
    prod.lib.RaylibB200_ReleaseInspection(scene)
    assert prod.lib.Raylib_DestroyScene(scene) == 1
    assert prod.lib.Raylib_UnloadOBJModel(model) == 1


def test_flattened_scene_cache_roundtrip(prod, rl, restate, tmp_path):
    """RaylibB200_SaveFlattenedScene / _LoadFlattenedScene: every uploaded array comes back bit for
    bit, the oracle restatement finds the same primary hits on the loaded scene (reference topology and the quantized
    tree the device walks), damaged or foreign files are refused."""
    lib = prod.lib
    rs = restate
    for cfg, size in ((6, 0), (4, 12)):
        info = prod.create_demo(cfg, size)
        path = str(tmp_path / ("scene%d.rtflat" % cfg)).encode()
        assert lib.RaylibB200_SaveFlattenedScene(info.scene, path) == 1, prod.last_error()
        loaded = lib.RaylibB200_LoadFlattenedScene(path)
        assert loaded, prod.last_error()
        a = prod.flat_desc(info.scene).contents
        b = prod.flat_desc(loaded).contents
        for count, fields in (("numWideNodes", [("quantNodes", 64)]), ("numRefNodes", [("refNodes", 64)]),
                              ("numTris", [("triHot", 64), ("triCold", 64), ("triRank", 4), ("triGate", 4)]),
                              ("numGates", [("gateBoxes", 32)]), ("numSpheres", [("spheres", 16), ("sphereMaterial", 4), ("sphereRank", 4), ("sphereGate", 4)]),
                              ("numCubes", [("cubes", 48), ("cubeRank", 4), ("cubeGate", 4)]), ("numMaterials", [("materials", 64)]),
                              ("numTextures", [("textures", 32)]), ("numTexels", [("texels", 16)])):
            n = getattr(a, count)
            assert n == getattr(b, count), count
            for name, size_b in fields:
                if n:
                    x = np.ctypeslib.as_array(C.cast(getattr(a, name), C.POINTER(C.c_uint8)), shape=(n * size_b,))
                    y = np.ctypeslib.as_array(C.cast(getattr(b, name), C.POINTER(C.c_uint8)), shape=(n * size_b,))
                    assert np.array_equal(x, y), name
        for f in ("wideRootRef", "wideMaxStack", "refRootRef", "refRootBoxTests", "refMaxDepth", "flags", "materialTypeMask", "numLeaves", "skyTexture"):
            assert getattr(a, f) == getattr(b, f), f
        assert list(a.rootMin) == list(b.rootMin) and list(a.rootMax) == list(b.rootMax)
        assert list(a.sunIlluminance) == list(b.sunIlluminance) and list(a.sunDirection) == list(b.sunDirection) and list(a.skyRotation) == list(b.skyRotation)
        cam = prod.camera_block(info.camera)
        W, H = 96, 54
        prod.set_viewport(info, W, H)
        cam = prod.camera_block(info.camera)
        for tree in (0, 3):
            rs.select_tree(tree)
            r0, t0, _ = rs.primary(prod.flat_desc(info.scene), cam, W, H, info.settings.rayTMin)
            r1, t1, _ = rs.primary(prod.flat_desc(loaded), cam, W, H, info.settings.rayTMin)
            assert np.array_equal(r0, r1) and np.array_equal(bits(t0), bits(t1))
        rs.select_tree(0)
        assert (r0 >= 0).mean() > 0.2
        # a loaded scene saves again to the same bytes
        path2 = str(tmp_path / ("again%d.rtflat" % cfg)).encode()
        assert lib.RaylibB200_SaveFlattenedScene(loaded, path2) == 1
        raw = open(path.decode(), "rb").read()
        assert raw == open(path2.decode(), "rb").read()
        assert lib.Raylib_DestroyScene(loaded) == 1
        prod.destroy_demo(info)
        # refused: flipped payload byte, truncation, foreign file, missing file
        bad = bytearray(raw); bad[len(bad) // 2] ^= 0x40
        for name, blob in (("flip", bytes(bad)), ("short", raw[:len(raw) - 17]), ("long", raw + b"\0"), ("foreign", b"P6 1 1 255 abc" * 10)):
            q = str(tmp_path / (name + ".rtflat"))
            open(q, "wb").write(blob)
            assert not lib.RaylibB200_LoadFlattenedScene(q.encode()), name
            assert "flattened-scene" in prod.last_error()
        assert not lib.RaylibB200_LoadFlattenedScene(str(tmp_path / "missing.rtflat").encode())


def _np_struct(ptr, count, words):
    return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint32)), shape=(count, words))


@pytest.mark.parametrize("env", [{}, {"RAYLIB_B200_SAH_AXES": "1", "RAYLIB_B200_SAH_ROTATIONS": "0"},
                                 {"RAYLIB_B200_SAH_AXES": "3", "RAYLIB_B200_SAH_ROTATIONS": "3"}, {"RAYLIB_B200_COLLAPSE": "dp"},
                                 {"RAYLIB_B200_COLLAPSE_PARALLEL_FROM": "1000"}, {"RAYLIB_B200_COLLAPSE_PARALLEL_FROM": "1000", "RAYLIB_B200_COLLAPSE": "dp"},
                                 {"RAYLIB_B200_SAH_TWO_LEVEL": "1"}, {"RAYLIB_B200_SAH_TWO_LEVEL": "1", "RAYLIB_B200_PARALLEL_WALK": "1", "RAYLIB_B200_SAH_ROTATIONS": "2"}])
def test_traversal_tree_structure(prod, env):
    """Whatever the builder options (split policy, tree rotations, collapse): the binary SAH tree and its 4-wide collapse
    hold every leaf exactly once, every record is reachable exactly once, and every box is the exact union of the boxes
    below it (what the equivalence argument of bvh_sah.h rests on)."""
    libc = C.CDLL(None)
    saved = {k: os.environ.get(k) for k in env}
    for k, v in env.items():
        os.environ[k] = v; libc.setenv(k.encode(), v.encode(), 1)
    try:
        info = prod.create_demo(4, 40)                  # 40 meshes + ground: 51,202 triangles
        d = prod.flat_desc(info.scene).contents
        n = d.numNodes
        nodes = _np_struct(d.nodes, n, 16)
        f = nodes.view(np.float32)
        KIND = lambda r: r >> 28
        IDX = lambda r: r & 0x0FFFFFFF
        # binary tree: boxes as (lo, hi) per side, refs in words 3 and 7
        seen_nodes, leaves = np.zeros(n, dtype=np.int32), []
        box_lo, box_hi = np.empty((n, 3), np.float32), np.empty((n, 3), np.float32)
        order, stack = [], [IDX(int(d.rootRef))]
        assert KIND(int(d.rootRef)) == 0
        while stack:
            i = stack.pop(); order.append(i); seen_nodes[i] += 1
            for ref in (int(nodes[i, 3]), int(nodes[i, 7])):
                if KIND(ref) == 0: stack.append(IDX(ref))
                else: leaves.append(ref)
        assert (seen_nodes == 1).all(), "every binary record is reachable exactly once"
        tri_leaves = sorted(IDX(r) for r in leaves if KIND(r) == 1)
        assert tri_leaves == list(range(d.numTris)), "every triangle is a leaf exactly once"
        for i in reversed(order):                        # children before parents
            lo = np.minimum(f[i, 0:3], f[i, 8:11]); hi = np.maximum(f[i, 4:7], f[i, 12:15])
            box_lo[i], box_hi[i] = lo, hi
            for side, ref in ((0, int(nodes[i, 3])), (8, int(nodes[i, 7]))):
                if KIND(ref) == 0:
                    c = IDX(ref)
                    assert np.array_equal(f[i, side:side + 3], box_lo[c]) and np.array_equal(f[i, side + 4:side + 7], box_hi[c]), "child box = exact union"
        r = IDX(int(d.rootRef))
        assert np.array_equal(box_lo[r], np.array(list(d.rootMin), np.float32)) and np.array_equal(box_hi[r], np.array(list(d.rootMax), np.float32))
        # 4-wide collapse: SoA record {lox[4] loy[4] loz[4] hix[4] hiy[4] hiz[4] ref[4] pad[4]}
        w = d.numWideNodes
        wide = _np_struct(d.wideNodes, w, 32); wf = wide.view(np.float32)
        seen_w, wleaves, stack = np.zeros(w, dtype=np.int32), [], [IDX(int(d.wideRootRef))]
        worder = []
        while stack:
            i = stack.pop(); worder.append(i); seen_w[i] += 1
            for k in range(4):
                ref = int(wide[i, 24 + k])
                if ref == 0xFFFFFFFF: continue
                if KIND(ref) == 0: stack.append(IDX(ref))
                else: wleaves.append(ref)
        assert (seen_w == 1).all(), "every wide record is reachable exactly once"
        assert sorted(wleaves) == sorted(leaves), "the collapse keeps the leaves"
        wlo, whi = np.empty((w, 3), np.float32), np.empty((w, 3), np.float32)
        for i in reversed(worder):
            present = [k for k in range(4) if int(wide[i, 24 + k]) != 0xFFFFFFFF]
            lo = np.array([[wf[i, 0 + k], wf[i, 4 + k], wf[i, 8 + k]] for k in present], np.float32)
            hi = np.array([[wf[i, 12 + k], wf[i, 16 + k], wf[i, 20 + k]] for k in present], np.float32)
            wlo[i], whi[i] = lo.min(axis=0), hi.max(axis=0)
            for j, k in enumerate(present):
                ref = int(wide[i, 24 + k])
                if KIND(ref) == 0:
                    assert np.array_equal(lo[j], wlo[IDX(ref)]) and np.array_equal(hi[j], whi[IDX(ref)])
        wr = IDX(int(d.wideRootRef))
        assert np.array_equal(wlo[wr], box_lo[r]) and np.array_equal(whi[wr], box_hi[r])
        prod.destroy_demo(info)
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None); libc.unsetenv(k.encode())
            else:
                os.environ[k] = v; libc.setenv(k.encode(), v.encode(), 1)


def test_threaded_builder_passes_in_a_subprocess(tmp_path):
    """The SAH builder's chunked passes (centroid bounds, binning, stable partition) normally start at 512 K items; the
    threshold is latched per process, so a child process lowers it to 64 and checks that the reference topology and the
    quantized traversal tree still select the same hits (oracle restatement), twice with identical arrays."""
    script = tmp_path / "threaded_builder.py"
    script.write_text(
        "import sys, ctypes as C, hashlib\n"
        "sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "import numpy as np, pyraylib as rl\n"
        "from oracle import bindings as ob\n"
        "p = rl.Product(); p.lib.Raylib_Initialize(); rs = ob.Restatement()\n"
        "hashes = []\n"
        "for rep in range(2):\n"
        "    info = p.create_demo(4, 40)\n"
        "    p.set_viewport(info, 96, 54)\n"
        "    d = p.flat_desc(info.scene); cam = p.camera_block(info.camera)\n"
        "    rs.select_tree(0); r0, t0, _ = rs.primary(d, cam, 96, 54, info.settings.rayTMin)\n"
        "    rs.select_tree(3); r3, t3, _ = rs.primary(d, cam, 96, 54, info.settings.rayTMin)\n"
        "    rs.select_tree(0)\n"
        "    assert np.array_equal(r0, r3) and np.array_equal(t0.view(np.uint32), t3.view(np.uint32))\n"
        "    assert (r0 >= 0).mean() > 0.2 and rs.check_quantization(d) == 0\n"
        "    q = np.ctypeslib.as_array(C.cast(d.contents.quantNodes, C.POINTER(C.c_uint32)), shape=(d.contents.numWideNodes * 16,))\n"
        "    hashes.append(hashlib.md5(q.tobytes()).hexdigest())\n"
        "    p.destroy_demo(info)\n"
        "assert hashes[0] == hashes[1]\n"
        "print('OK', hashes[0])\n" % (os.path.join(ROOT, "software-raytracing_b200"), ROOT))
    env = dict(os.environ, RAYLIB_B200_SAH_PARALLEL_FROM="64", RAYLIB_B200_COLLAPSE_PARALLEL_FROM="1000")
    out = subprocess.run([sys.executable, str(script)], env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


def test_child_box_grids_in_a_subprocess(tmp_path):
    """RtNodeQ4 child boxes (include/rt_scene_format.h): the 256-value grid with a free scale (default) and the earlier
    127-step power-of-two grid (RAYLIB_B200_Q4_GRID=7, latched per process) both contain the exact boxes and select the
    reference's hits; the default pads the boxes less (that is its point: fewer node visits on the device)."""
    script = tmp_path / "grids.py"
    script.write_text(
        "import sys, ctypes as C\n"
        "sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "import numpy as np, pyraylib as rl\n"
        "from oracle import bindings as ob\n"
        "p = rl.Product(); p.lib.Raylib_Initialize(); rs = ob.Restatement()\n"
        "pads = []\n"
        "for cfg, size in ((4, 40), (5, 8), (6, 0)):\n"
        "    info = p.create_demo(cfg, size)\n"
        "    p.set_viewport(info, 96, 54)\n"
        "    d = p.flat_desc(info.scene); cam = p.camera_block(info.camera)\n"
        "    rs.select_tree(0); r0, t0, _ = rs.primary(d, cam, 96, 54, info.settings.rayTMin)\n"
        "    rs.select_tree(3); r3, t3, _ = rs.primary(d, cam, 96, 54, info.settings.rayTMin)\n"
        "    rs.select_tree(0)\n"
        "    assert np.array_equal(r0, r3) and np.array_equal(t0.view(np.uint32), t3.view(np.uint32))\n"
        "    assert rs.check_quantization(d) == 0\n"
        "    n = d.contents.numWideNodes\n"
        "    q = np.ctypeslib.as_array(C.cast(d.contents.quantNodes, C.POINTER(C.c_uint32)), shape=(n, 16))\n"
        "    w = np.ctypeslib.as_array(C.cast(d.contents.wideNodes, C.POINTER(C.c_float)), shape=(n, 32))\n"
        "    # x axis of child 0: decoded extent against the exact one\n"
        "    base, S = q[:, 0].view(np.float32), q[:, 3].view(np.float32)\n"
        "    m = lambda b: ((0x3F000000 | (b.astype(np.uint32) << 16)).astype(np.uint32)).view(np.float32)\n"
        "    lo = m(q[:, 4] & 255) * S + base; hi = m(q[:, 7] & 255) * S + base\n"
        "    ext = w[:, 12] - w[:, 0]\n"
        "    ok = np.isfinite(ext) & (ext > 0) & (np.abs(w[:, 0]) < 1e17) & (np.abs(w[:, 12]) < 1e17)\n"
        "    pads.append(float(np.mean(((hi - lo)[ok] - ext[ok]) / ext[ok])))\n"
        "    p.destroy_demo(info)\n"
        "print('PAD', sum(pads) / len(pads))\n" % (os.path.join(ROOT, "software-raytracing_b200"), ROOT))
    pad = {}
    for grid in ("256", "7"):
        env = dict(os.environ, RAYLIB_B200_Q4_GRID=grid)
        out = subprocess.run([sys.executable, str(script)], env=env, capture_output=True, text=True, timeout=600)
        assert out.returncode == 0 and "PAD" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
        pad[grid] = float(out.stdout.strip().splitlines()[-1].split()[1])
    assert 0.0 <= pad["256"] < 0.5 * pad["7"], pad


def test_bench_contract_without_a_gpu():
    """bench.py on a host without a GPU: the reference arm (the compiled reference's own renderer on the host cores) prints
    one JSON line with the contract's keys; the product arm refuses to run -- there is no CPU fallback behind it."""
    import json
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libraylib_ref.so")):
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is visible here")
    except ImportError:
        pass
    bench = os.path.join(ROOT, "bench.py")
    small = ["--workload", "random_spheres_640x360_16spp_d5", "--steps", "1", "--warmup", "0"]
    out = subprocess.run([sys.executable, bench, "--impl", "reference"] + small, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "Mrays/s" and line["unit"] == "Mrays/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["ms_per_step"] > 0 and line["n_gpus"] == 1 and line["steps"] == 1
    assert line["config"]["workload"] == "random_spheres_640x360_16spp_d5" and line["config"]["spheres"] > 100
    assert line["cpu_baseline"]["kind"] == "reference" and line["cpu_baseline"]["cores"] >= 1 and line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # a non-zero rank of a torchrun launch prints nothing and exits 0
    quiet = subprocess.run([sys.executable, bench, "--impl", "reference"] + small, capture_output=True, text=True, timeout=600,
                           env=dict(os.environ, RANK="1", WORLD_SIZE="2"))
    assert quiet.returncode == 0 and quiet.stdout.strip() == ""
    out = subprocess.run([sys.executable, bench, "--no-cpu-baseline"] + small, capture_output=True, text=True, timeout=600)
    assert out.returncode != 0 and "needs a CUDA device" in out.stderr and "{" not in out.stdout


def _adversarial_rays(lo, hi, n, seed):
    """The ray families of tests/test_gpu_parity.py::test_adversarial_rays_against_restatement: axis-parallel directions
    (both signs of zero), denormal / tiny components, origins on box planes and round coordinates, origins 1e3..1e6 away."""
    lo = np.maximum(lo, -50.0); hi = np.minimum(hi, 50.0)
    rng = np.random.default_rng(seed)
    o = rng.uniform(lo - 0.5, hi + 0.5, size=(n, 3))
    v = rng.normal(size=(n, 3)); d = v / np.linalg.norm(v, axis=1, keepdims=True)
    k = n // 6
    axis = rng.integers(0, 3, size=k); d[:k] = 0.0; d[np.arange(k), axis] = rng.choice([-1.0, 1.0], size=k)
    d[:k // 2][d[:k // 2] == 0.0] = -0.0
    z = rng.integers(0, 3, size=k); d[np.arange(k, 2 * k), z] = 0.0
    t = rng.integers(0, 3, size=k); d[np.arange(2 * k, 3 * k), t] = rng.choice([1e-42, -1e-42, 1e-20, -1e-20, 3e-39], size=k)
    a = rng.integers(0, 3, size=k); o[np.arange(3 * k, 4 * k), a] = np.where(rng.random(k) < 0.5, lo[a], hi[a])
    o[4 * k:5 * k] = np.round(o[4 * k:5 * k] * 2.0) / 2.0
    c = 0.5 * (lo + hi); far = rng.normal(size=(n - 5 * k, 3)); far /= np.linalg.norm(far, axis=1, keepdims=True)
    dist = 10.0 ** rng.uniform(3, 6, size=(n - 5 * k, 1))
    target = rng.uniform(lo, hi, size=(n - 5 * k, 3))
    o[5 * k:] = c + far * dist
    dd = target - o[5 * k:]; d[5 * k:] = dd / np.linalg.norm(dd, axis=1, keepdims=True)
    rays = np.zeros((n, 8), dtype=np.float32)
    rays[:, 0:3] = o; rays[:, 4:7] = d
    return rays


@pytest.mark.parametrize("cfg,size", [(2, 0), (4, 24), (6, 0), (1, 0), (5, 8), (3, 64)])
def test_device_node_test_restated_on_adversarial_rays(prod, restate, cfg, size):
    """The kernels' conservative inner-node test (quantized planes evaluated in ray space with slack, rt_traverse.cuh
    trav_step) restated operation by operation in oracle/rt_oracle.c (tree mode 4): on rays chosen to stress it, a walk that
    culls with it finds exactly what the exhaustive walk of the reference topology finds -- checked here without a GPU (the GPU
    suite runs the same rays through the kernels)."""
    info = prod.create_demo(cfg, size)
    try:
        desc = prod.flat_desc(info.scene)
        lo = np.array(desc.contents.rootMin[:], dtype=np.float64); hi = np.array(desc.contents.rootMax[:], dtype=np.float64)
        rays = _adversarial_rays(lo, hi, 24000, 11 + cfg)
        restate.select_tree(0)
        r0, t0, _ = restate.trace(desc, rays, 1e-4)
        restate.select_tree(4)
        r4, t4, c4 = restate.trace(desc, rays, 1e-4)
        restate.select_tree(3)
        r3, t3, c3 = restate.trace(desc, rays, 1e-4)
        restate.select_tree(0)
        assert (r0 >= 0).mean() > 0.05
        assert int(((r4 < 0) & (r0 >= 0)).sum()) == 0, "the conservative test lost a hit"
        assert np.array_equal(r0, r4) and np.array_equal(t0.view(np.uint32), t4.view(np.uint32))
        # (mode 3 tests the decoded boxes with the reference's exact slab arithmetic: from 1e5..1e6 scene sizes away that is
        # not wide enough for the tight per-triangle boxes of the SAH tree -- the reason the kernels' test carries its slack;
        # the far family is therefore left out of this one comparison)
        near = 5 * (len(rays) // 6)
        assert np.array_equal(r0[:near], r3[:near])
        # conservative means it passes at least what the decoded boxes pass -- and, for rays that start near the scene, not
        # many more (from 1e6 away the slack is several scene units wide: correct, and slow)
        restate.select_tree(4); _, _, n4 = restate.trace(desc, rays[:near], 1e-4)
        restate.select_tree(3); _, _, n3 = restate.trace(desc, rays[:near], 1e-4)
        assert c4[0] >= c3[0] and n4[0] >= n3[0] and n4[0] < 1.1 * n3[0] + 1000, (n3[0], n4[0], c3[0], c4[0])
    finally:
        restate.select_tree(0)
        prod.destroy_demo(info)


def test_quantizer_stress(tmp_path):
    """RtQuantizeWide on synthetic nodes with extreme coordinates (tests/native/quantizer_stress.cc), both grids: no decoded
    box may miss a part of the exact one."""
    exe = str(tmp_path / "quantizer_stress")
    host = os.path.join(ROOT, "software-raytracing_b200", "csrc", "host")
    build = subprocess.run(["g++", "-std=c++17", "-O2", "-pthread", "-I" + os.path.join(ROOT, "include"), "-I" + host,
                            os.path.join(ROOT, "tests", "native", "quantizer_stress.cc"), os.path.join(host, "bvh_sah.cc"), "-o", exe],
                           capture_output=True, text=True, timeout=300)
    assert build.returncode == 0, build.stderr[-2000:]
    for grid in ("256", "7"):
        out = subprocess.run([exe], env=dict(os.environ, RAYLIB_B200_Q4_GRID=grid), capture_output=True, text=True, timeout=300)
        assert out.returncode == 0 and "violations 0 of" in out.stdout, out.stdout[-1000:]


def _flat_arrays(prod, scene):
    d = prod.flat_desc(scene).contents
    arrays = {
        "triHot": _np_struct(d.triHot, d.numTris, 16), "triCold": _np_struct(d.triCold, d.numTris, 16),
        "triRank": _np_struct(d.triRank, d.numTris, 1), "triGate": _np_struct(d.triGate, d.numTris, 1),
        "gateBoxes": _np_struct(d.gateBoxes, d.numGates, 8), "refNodes": _np_struct(d.refNodes, d.numRefNodes, 16),
        "nodes": _np_struct(d.nodes, d.numNodes, 16), "quantNodes": _np_struct(d.quantNodes, d.numWideNodes, 16),
        "materials": _np_struct(d.materials, d.numMaterials, 16),
    }
    scalars = (d.numLeaves, d.refMaxDepth, d.refRootRef, d.rootRef, d.wideRootRef, d.wideMaxStack, d.maxStackDepth, d.materialTypeMask)
    return {k: v.copy() for k, v in arrays.items()}, scalars


def test_parallel_walk_and_two_level_build(prod, restate):
    """Time to first frame: scenes made of many meshes are flattened with one worker task per StaticMesh
    (graph walk) and one SAH subtree per mesh under a small top tree.  The parallel walk must produce the serial walk's
    arrays bit for bit; the two-level tree is another tree over the same leaf groups, so it must select the same hits
    (restatement on the quantized tree == reference topology == golden vectors)."""
    import ctypes
    from conftest import load_golden, bits
    libc = ctypes.CDLL(None)

    def flat(env):
        for k, v in env.items():
            os.environ[k] = v; libc.setenv(k.encode(), v.encode(), 1)
        try:
            info = prod.create_demo(4, 40)
            try:
                return _flat_arrays(prod, info.scene)
            finally:
                prod.lib.RaylibB200_ReleaseInspection(info.scene)
                prod.destroy_demo(info)
        finally:
            for k in env:
                os.environ.pop(k, None); libc.unsetenv(k.encode())

    serial = flat({"RAYLIB_B200_PARALLEL_WALK": "0", "RAYLIB_B200_SAH_TWO_LEVEL": "0"})
    parallel = flat({"RAYLIB_B200_PARALLEL_WALK": "1", "RAYLIB_B200_SAH_TWO_LEVEL": "0"})
    assert serial[1] == parallel[1]
    for name in serial[0]:
        assert np.array_equal(serial[0][name], parallel[0][name]), name + " differs between the serial and the parallel walk"
    two = flat({"RAYLIB_B200_PARALLEL_WALK": "1", "RAYLIB_B200_SAH_TWO_LEVEL": "1"})
    for name in ("triHot", "triCold", "triRank", "triGate", "gateBoxes", "refNodes", "materials"):
        assert np.array_equal(serial[0][name], two[0][name]), name
    assert not np.array_equal(serial[0]["nodes"], two[0]["nodes"]), "the two-level build was not taken"

    # hits through the two-level tree against the golden vectors of config 4
    g = load_golden(4)
    os.environ["RAYLIB_B200_SAH_TWO_LEVEL"] = "1"; libc.setenv(b"RAYLIB_B200_SAH_TWO_LEVEL", b"1", 1)
    os.environ["RAYLIB_B200_PARALLEL_WALK"] = "1"; libc.setenv(b"RAYLIB_B200_PARALLEL_WALK", b"1", 1)
    try:
        info = prod.create_demo(4, int(g["size"]))
        try:
            desc = prod.flat_desc(info.scene)
            assert restate.check_quantization(desc) == 0
            for tree in (1, 2, 3):
                restate.select_tree(tree)
                try:
                    rank, t, _ = restate.trace(desc, g["rays"], info.settings.rayTMin)
                finally:
                    restate.select_tree(0)
                assert np.array_equal(rank, g["rank"]) and np.array_equal(bits(t), bits(g["t"])), "tree %d" % tree
        finally:
            prod.destroy_demo(info)
    finally:
        for k in ("RAYLIB_B200_SAH_TWO_LEVEL", "RAYLIB_B200_PARALLEL_WALK"):
            os.environ.pop(k, None); libc.unsetenv(k.encode())


def test_obj_fast_path_equals_object_path(prod, tmp_path):
    """Raylib_LoadOBJModel keeps the faces as arrays and StaticMesh::Finalize emits the flattened records
    directly (compact_mesh.h) -- no Triangle / BVHNode objects.  RAYLIB_B200_OBJ_OBJECTS=1 forces the reference's object
    representation; both must flatten to the same bits, with and without Raylib_TransformOBJModel, on a generated OBJ of
    64 shapes / 204,800 triangles (many meshes: the parallel placement path) and on the small multi-material file."""
    import ctypes, time
    from test_cpu_oracle import write_test_obj
    libc = ctypes.CDLL(None)
    lib = prod.lib
    small = write_test_obj(str(tmp_path))
    big = str(tmp_path / "big.obj")
    with open(big, "w") as f:
        G = 320
        f.write("mtllib parity.mtl\n")
        xs = np.linspace(-4.0, 4.0, G + 1)
        for j in range(G + 1):
            f.write("".join("v %.9g %.9g %.9g\n" % (x, 0.2 * np.sin(2.0 * x) * np.cos(1.5 * xs[j]), xs[j]) for x in xs))
        rows_per_shape = G // 64
        for shape in range(64):
            f.write("g strip%d\nusemtl %s\n" % (shape, ("clay", "chrome", "glass", "glow")[shape % 4]))
            out = []
            for j in range(shape * rows_per_shape, (shape + 1) * rows_per_shape):
                for i in range(G):
                    a, b, c, d = j * (G + 1) + i + 1, j * (G + 1) + i + 2, (j + 1) * (G + 1) + i + 2, (j + 1) * (G + 1) + i + 1
                    out.append("f %d %d %d\nf %d %d %d\n" % (a, c, b, a, d, c))
            f.write("".join(out))

    def flat(path, objects, transform):
        os.environ["RAYLIB_B200_OBJ_OBJECTS"] = objects; libc.setenv(b"RAYLIB_B200_OBJ_OBJECTS", objects.encode(), 1)
        try:
            t0 = time.time()
            model = lib.Raylib_LoadOBJModel(path.encode())
            assert model
            if transform:
                lib.Raylib_TransformOBJModel(model, 0.5, 1.0, -0.25, 20.0, 5.0, 0.0, 1.5, 1.0, 0.75)
            lib.Raylib_FinalizeOBJModel(model)
            scene = lib.Raylib_CreateScene()
            lib.Raylib_AddOBJModelToScene(scene, model)
            lib.Raylib_FinalizeScene(scene)
            arrays = _flat_arrays(prod, scene)
            seconds = time.time() - t0
            lib.RaylibB200_ReleaseInspection(scene)
            assert lib.Raylib_DestroyScene(scene) == 1 and lib.Raylib_UnloadOBJModel(model) == 1
            return arrays, seconds
        finally:
            os.environ.pop("RAYLIB_B200_OBJ_OBJECTS", None); libc.unsetenv(b"RAYLIB_B200_OBJ_OBJECTS")

    for path, transform in ((small, False), (small, True), (big, False), (big, True)):
        (fast, fast_scalars), t_fast = flat(path, "0", transform)
        (objs, obj_scalars), t_objs = flat(path, "1", transform)
        assert fast_scalars == obj_scalars
        for name in fast:
            assert np.array_equal(fast[name], objs[name]), "%s differs between the fast path and the object path (%s)" % (name, os.path.basename(path))
        print("%s transform=%s: %d triangles, load -> flattened %.2f s (fast path) vs %.2f s (objects)" % (
            os.path.basename(path), transform, len(fast["triHot"]), t_fast, t_objs))
    assert len(fast["triHot"]) == 204800
