"""Analysis tool (CPU, host simulator): what a coarser slab test would cost - every child box of the wide BVH widened by k quanta per
side, nodes and primitive tests per ray on the C4 scene.  python tests/cpu_plane_pad_experiment.py"""
import sys, time
sys.path.insert(0, ".")
import numpy as np
from ilgpu_raytracing_b200 import layouts as L, scenes
from tests.hostsim_binding import HostSimScene, set_plane_pad
from tests.util import oracle_camera, oracle_scene_from_spec
sc = oracle_scene_from_spec(scenes.terrain_scene(n_quads=708, n_spheres=256))
hs = HostSimScene(sc.arrays())
W, H = 192, 108
cam = oracle_camera("C3", W, H)
base = None
for pad in (0.0, 0.5, 1.0, 1.5, 2.0, 3.0):
    set_plane_pad(pad)
    t = time.time()
    h = hs.render(cam, L.make_render_config(W, H, spp=4, max_depth=8), aovs=False)
    c = h["counters"]
    rays = c["raysPrimary"] + c["raysBounce"] + c["raysShadow"]
    if base is None: base = (c["nodes"], c["tris"] + c["spheres"], h["rgba8"].copy())
    print(f"pad {pad}: nodes/ray {c['nodes']/rays:.2f} (x{c['nodes']/base[0]:.3f}), prims/ray {(c['tris']+c['spheres'])/rays:.2f} (x{(c['tris']+c['spheres'])/base[1]:.3f}), image equal {np.array_equal(h['rgba8'], base[2])}, {time.time()-t:.0f}s", flush=True)
set_plane_pad(0.0)
