"""Soak test (development, GPU box): many scenes / sizes / seeds, device traversal vs the COMPILED REFERENCE on random
and adversarial rays.  Prints one line per case and a final tally; exits 1 on any lost hit or mismatch above the budget.

  python tools/soak_parity.py [cases]
"""
import os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "software-raytracing_b200"))
import pyraylib as rl
sys.path.insert(0, ROOT)
from oracle import bindings as ob


def make_rays(rng, lo, hi, n):
    lo = np.maximum(lo, -60.0); hi = np.minimum(hi, 60.0)
    o = rng.uniform(lo - 0.5, hi + 0.5, size=(n, 3))
    v = rng.normal(size=(n, 3)); d = v / np.linalg.norm(v, axis=1, keepdims=True)
    k = n // 8
    ax = rng.integers(0, 3, size=k); d[:k] = 0.0; d[np.arange(k), ax] = rng.choice([-1.0, 1.0], size=k)
    z = rng.integers(0, 3, size=k); d[np.arange(k, 2 * k), z] = rng.choice([0.0, -0.0, 1e-42, -1e-20, 3e-39], size=k)
    a = rng.integers(0, 3, size=k); o[np.arange(2 * k, 3 * k), a] = np.where(rng.random(k) < 0.5, lo[a], hi[a])
    o[3 * k:4 * k] = np.round(o[3 * k:4 * k] * 4.0) / 4.0
    c = 0.5 * (lo + hi); far = rng.normal(size=(k, 3)); far /= np.linalg.norm(far, axis=1, keepdims=True)
    o[4 * k:5 * k] = c + far * 10.0 ** rng.uniform(2, 6, size=(k, 1))
    dd = rng.uniform(lo, hi, size=(k, 3)) - o[4 * k:5 * k]; d[4 * k:5 * k] = dd / np.linalg.norm(dd, axis=1, keepdims=True)
    rays = np.zeros((n, 8), dtype=np.float32)
    rays[:, 0:3] = o; rays[:, 4:7] = d; rays[:, 3] = rng.uniform(0, 5, size=n)
    return rays


def main():
    cases = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    prod, ref = rl.Product(), ob.Reference()
    prod.require_gpu()
    prod.lib.Raylib_Initialize()
    rng = np.random.default_rng(2026)
    total = lost = mism = tdiff = 0
    sizes = {1: [0], 2: [0], 3: [16, 64, 200, 400], 4: [8, 24, 60, 120], 5: [16, 48, 96], 6: [0]}
    for case in range(cases):
        cfg = [3, 4, 5, 6, 1, 2][case % 6]
        size = int(rng.choice(sizes[cfg]))
        pinfo, rinfo = prod.create_demo(cfg, size), ref.create_demo(cfg, size)
        d = prod.flat_desc(pinfo.scene).contents
        lo, hi = np.array(d.rootMin[:], dtype=np.float64), np.array(d.rootMax[:], dtype=np.float64)
        rays = make_rays(rng, lo, hi, 160000)
        tmin = float(rng.choice([1e-4, 1e-3, 0.0, 1e-6]))
        gr, gt = prod.trace_rays(pinfo.scene, rays, tmin)
        rr, rt, _ = ref.trace_rays(rinfo.scene, rays, tmin)
        l = int(((gr < 0) & (rr >= 0)).sum()); m = int((gr != rr).sum())
        same = gr == rr
        td = int((gt[same].view(np.uint32) != rt[same].view(np.uint32)).sum())
        total += len(rays); lost += l; mism += m; tdiff += td
        for i in np.nonzero(gr != rr)[0][:4]:
            print(json.dumps({"detail": {"cfg": cfg, "size": size, "tmin": tmin, "ray": [float(x) for x in rays[i]], "device": [int(gr[i]), float(gt[i])],
                                         "reference": [int(rr[i]), float(rt[i])]}}), flush=True)
        print(json.dumps({"case": case, "cfg": cfg, "size": size, "tmin": tmin, "tris": int(d.numTris), "spheres": int(d.numSpheres),
                          "hit_frac": round(float((rr >= 0).mean()), 3), "lost": l, "id_mismatch": m, "t_bit_diff": td}), flush=True)
        prod.lib.RaylibB200_ReleaseInspection(pinfo.scene)
        prod.destroy_demo(pinfo); ref.destroy_demo(rinfo)
    print(json.dumps({"rays": total, "lost": lost, "id_mismatch": mism, "t_bit_diff": tdiff, "mismatch_rate": mism / max(1, total)}))
    prod.lib.Raylib_Terminate()
    sys.exit(1 if lost or mism > 1e-5 * total or tdiff else 0)


if __name__ == "__main__":
    sys.stdout.flush()
    real = os.dup(1); os.dup2(2, 1); sys.stdout = os.fdopen(real, "w")
    main()
