"""Manual tuning tool (under gpurun): builds the core with RT_PHASE_STATS=1 (plus any -D given) and prints the lane participation of
the extend kernel's phases for one C4 frame at 8 spp."""
import os
import subprocess
import sys

sys.path.insert(0, ".")
from ilgpu_raytracing_b200 import build  # noqa: E402

so = "/tmp/librtcore_phase.so"
subprocess.run(["nvcc"] + build.NVCC_FLAGS + ["-DRT_PHASE_STATS=1"] + [f"-D{d}" for d in sys.argv[1:]] + ["-o", so] + build.CORE_SRCS, check=True)
code = """
import sys; sys.path.insert(0, '.')
from ilgpu_raytracing_b200 import layouts as L, native, scenes
from tests.util import oracle_camera, oracle_scene_from_spec
W, H = 3840, 2160
sc = oracle_scene_from_spec(scenes.terrain_scene(n_quads=708, n_spheres=256))
ctx = native.Context(0); ctx.scene_upload(sc.arrays()); cam = oracle_camera('C3', W, H)
cfg = L.make_render_config(W, H, spp=8, max_depth=8, flags=L.RT_FLAG_COUNTERS)
ctx.render(cam, cfg); ctx.sync(); s = ctx.stats()
print({k: s[k] for k in ('raysPrimary', 'raysBounce', 'raysShadow', 'raysAnyHitTraced', 'wideNodes', 'trisTested', 'spheresTested')})
"""
subprocess.run([sys.executable, "-c", code], env=dict(os.environ, RTCORE_B200_LIB=so), check=True)
