"""Writes tests/golden/jpeg_*.jpg + .npy: small JPEG files (made with PIL / libjpeg-turbo) and the RGB bytes libjpeg-turbo
decodes them to, so that tests/test_cpu_host.py::test_jpeg_golden_fixtures can check csrc/host/jpeg_codec.cc without PIL.
Run from the repository root:  python tools/make_jpeg_golden.py"""
import os, sys
import numpy as np
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_cpu_host import _jpeg_test_picture

FIXTURES = [
    ("jpeg_baseline_420", (37, 26), dict(quality=75, subsampling=2)),
    ("jpeg_baseline_444_restart", (35, 21), dict(quality=90, subsampling=0, restart_marker_blocks=2)),
    ("jpeg_progressive_422", (41, 19), dict(quality=85, subsampling=1, progressive=True)),
    ("jpeg_optimized_grey", (30, 30), dict(quality=60, optimize=True)),
]
gold = os.path.join(ROOT, "tests", "golden")
for name, (w, h), options in FIXTURES:
    src = _jpeg_test_picture(w, h, len(name))
    path = os.path.join(gold, name + ".jpg")
    if "grey" in name:
        Image.fromarray(src[..., 0], "L").save(path, "JPEG", **options)
    else:
        Image.fromarray(src).save(path, "JPEG", **options)
    np.save(os.path.join(gold, name + ".npy"), np.asarray(Image.open(path).convert("RGB")))
    print(name, os.path.getsize(path), "bytes")
