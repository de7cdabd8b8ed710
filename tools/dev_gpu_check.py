"""Development smoke on a GPU box: parity of primary hits + radiance vs the compiled reference, quick timings."""
import os, sys, time, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "software-raytracing_b200"))
import pyraylib as rl
sys.path.insert(0, ROOT)
from oracle import bindings as ob

prod = rl.Product(); ref = ob.Reference()
print("devices", prod.device_count(), flush=True)
prod.lib.Raylib_Initialize()
out = {}
cases = [(6, 0, (320, 180), 8), (1, 0, (320, 180), 8), (2, 0, (320, 180), 8), (3, 128, (320, 180), 4), (4, 60, (320, 180), 2), (5, 96, (320, 180), 4)]
if len(sys.argv) > 1:
    cases = [c for c in cases if str(c[0]) in sys.argv[1].split(",")]
for cfg, size, (w, h), spp in cases:
    ri = ref.create_demo(cfg, size); pi = prod.create_demo(cfg, size)
    ref.set_viewport(ri, w, h); prod.set_viewport(pi, w, h)
    rank_r, t_r, rays, st = ref.primary_hits(ri.settings, ri.scene, ri.camera, want_rays=True)
    t0 = time.time()
    rank_g, t_g = prod.primary_hits(pi.settings, pi.scene, pi.camera)
    t1 = time.time()
    rank_g2, t_g2 = prod.trace_rays(pi.scene, rays, pi.settings.rayTMin)
    res = {
        "primary_rank_match": float((rank_r == rank_g).mean()), "primary_t_match": float((t_r.view(np.uint32) == t_g.view(np.uint32)).mean()),
        "trace_rank_match": float((rank_r == rank_g2).mean()), "trace_t_match": float((t_r.view(np.uint32) == t_g2.view(np.uint32)).mean()),
        "hitfrac": float((rank_r >= 0).mean()), "first_primary_s": t1 - t0,
    }
    # radiance parity
    s = pi.settings.copy(samplesPerPixel=spp)
    sr = ri.settings.copy(samplesPerPixel=spp)
    img_r, ost = ref.render_deterministic(sr, ri.scene, ri.camera)
    img_g = prod.render(s, pi.scene, pi.camera)
    stt = prod.last_stats()
    diff = np.abs(img_g.astype(np.float64) - img_r.astype(np.float64))
    rel = diff / (np.abs(img_r) + 1e-3)
    res.update({
        "psnr": float(rl.psnr(img_g, img_r)), "max_abs": float(np.nanmax(diff)), "mean_abs": float(np.nanmean(diff)),
        "frac_bitexact_px": float((img_g.view(np.uint32) == img_r.view(np.uint32)).all(axis=2).mean()),
        "frac_rel_gt_1e-3": float((rel > 1e-3).any(axis=2).mean()),
        "mean_ref": float(np.nanmean(img_r)), "mean_gpu": float(np.nanmean(img_g)), "nan_ref": int(np.isnan(img_r).sum()), "nan_gpu": int(np.isnan(img_g).sum()),
        "rays_ref": int(ost.rayQueries), "rays_gpu": int(stt.rayQueries), "device_ms": stt.deviceMs, "total_ms": stt.totalMs,
        "ref_seconds": ost.seconds, "ref_threads": ost.threads, "launches": stt.kernelLaunches,
    })
    # debug modes
    for mode in (1, 2, 4, 5):
        sm = pi.settings.copy(renderMode=mode); smr = ri.settings.copy(renderMode=mode)
        a, _ = ref.render_deterministic(smr, ri.scene, ri.camera)
        b = prod.render(sm, pi.scene, pi.camera)
        res["mode%d_bitexact" % mode] = float((a.view(np.uint32) == b.view(np.uint32)).all(axis=2).mean())
        res["mode%d_maxabs" % mode] = float(np.nanmax(np.abs(a - b)))
    out["cfg%d" % cfg] = res
    print("cfg", cfg, json.dumps(res), flush=True)
    ref.destroy_demo(ri); prod.destroy_demo(pi)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "dev_gpu_check.json"), "w"), indent=1)
prod.lib.Raylib_Terminate()
