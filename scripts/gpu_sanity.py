"""Ad-hoc first-contact check on a B200: hits vs the reference library, a small render, timings."""
import sys, time, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ptb200 as ptb
from ptb200 import procedural as P
import reflib

def cmp_hits(a, b, name):
    same_id = (a["instance"] == b["instance"]) & (a["surface"] == b["surface"]) & (a["triangle"] == b["triangle"])
    hit = b["instance"] != 0xFFFFFFFF
    tb = np.array_equal(a["t"].view(np.uint32), b["t"].view(np.uint32))
    bb = np.array_equal(a["bary"].view(np.uint32), b["bary"].view(np.uint32))
    print(f"{name}: n={len(a)} hit_frac={hit.mean():.3f} id_mismatch={int((~same_id).sum())} t_bitexact={tb} bary_bitexact={bb}")
    if (~same_id).any():
        i = np.nonzero(~same_id)[0][:5]
        print(" first mismatches:", i, a[i], b[i])

G = P.cornell_gltf_path()
print("devices", ptb.device_count(), "extend regs", ptb.lib().ptb_extend_registers())
ref = reflib.RefScene.from_gltf(G)
sc = ptb.Scene.load_gltf(G)
print(sc.info())
W = H = 128
ys, xs = np.mgrid[0:H, 0:W]
aa = np.full((W * H, 2), 0.5, np.float32)
od_ref = ref.camera_rays(W, H, xs.ravel(), ys.ravel(), aa)
od = sc.camera_rays(W, H, xs.ravel(), ys.ravel(), aa)
print("camera rays bit-exact:", np.array_equal(od.view(np.uint32), od_ref.view(np.uint32)))
h_ref, at_ref = ref.trace_rays(od_ref, attrs=True)
h, at = sc.trace_rays(od_ref, attrs=True)
cmp_hits(h, h_ref, "cornell primary")
print(" attrs max abs diff", np.abs(at - at_ref).max())
rng = np.random.default_rng(7)
n = 200000
o = rng.uniform(-3, 3, (n, 3)).astype(np.float32); o[:, 1] = rng.uniform(0, 5, n)
d = rng.normal(size=(n, 3)).astype(np.float32)
od2 = np.concatenate([o, d], 1)
h2_ref = ref.trace_rays(od2); h2 = sc.trace_rays(od2)
cmp_hits(h2, h2_ref, "cornell random")
# bounce rays from primary hits
pos = at_ref[:, 0:3]; dirs = rng.normal(size=pos.shape).astype(np.float32)
od3 = np.concatenate([pos + dirs / np.linalg.norm(dirs, axis=1, keepdims=True) * 1e-4, dirs], 1).astype(np.float32)
cmp_hits(sc.trace_rays(od3), ref.trace_rays(od3), "cornell bounce")

for mode in (0, 1):
    t = time.time(); rgb, alpha, st = sc.render_tile(64, 64, 256, 4 if mode == 0 else 8, integrator=mode); dt = time.time() - t
    r_rgb, r_a, r_rays, r_s = ref.render_linear(64, 64, 64, 4 if mode == 0 else 8, mode=mode)
    print(f"mode {mode}: gpu mean {rgb.mean((0,1))} ref mean {r_rgb.mean((0,1))} gpu rays/path {st['rays']/st['paths']:.3f} gpu_s {st['gpu_seconds']:.4f} wall {dt:.3f} ref_s {r_s:.2f}")
    print("   alpha", alpha.mean(), r_a.mean(), "nan:", np.isnan(rgb).sum())

# heightfield
for nn in (64, 250):
    d = P.heightfield_scene(nn)
    t = time.time(); hs = ptb.Scene.create(d); print(f"heightfield n={nn} create {time.time()-t:.2f}s", hs.info())
    fs = reflib.FlatScene(d.meshes, d.surfaces, d.instances, d.materials, d.camera)
    rs = reflib.RefScene.from_flat(fs)
    W, H = 256, 144
    ys, xs = np.mgrid[0:H, 0:W]
    od = rs.camera_rays(W, H, xs.ravel(), ys.ravel(), np.full((W * H, 2), 0.5, np.float32))
    hr, atr = rs.trace_rays(od, attrs=True); hg = hs.trace_rays(od)
    cmp_hits(hg, hr, f"heightfield{nn} primary")
    dirs = rng.normal(size=(len(od), 3)).astype(np.float32)
    od3 = np.concatenate([atr[:, 0:3] + dirs * 1e-4, dirs], 1).astype(np.float32)
    cmp_hits(hs.trace_rays(od3), rs.trace_rays(od3), f"heightfield{nn} bounce")
    rgb, alpha, st = hs.render_tile(256, 144, 64, 4)
    r_rgb, _, r_rays, r_s = rs.render_linear(256, 144, 8, 4, mode=2)
    print(f" render gpu mean {rgb.mean((0,1))} ref mean {r_rgb.mean((0,1))} rays/path gpu {st['rays']/st['paths']:.3f} ref {r_rays/(256*144*8):.3f} gpu_s {st['gpu_seconds']:.4f} Mrays/s {st['rays']/st['gpu_seconds']/1e6:.1f}")
print("DONE")
