import sys
sys.path.insert(0, ".")
import numpy as np
from ilgpu_raytracing_b200 import layouts as L, native
from oracle import orc
from tests.util import oracle_camera
sc = orc.Scene(); sc.build_default()
ctx = native.Context(0); ctx.scene_upload(sc.arrays())
W, H = 320, 180
cam = oracle_camera("C1B", W, H)
for flags in (0, L.RT_FLAG_FRAME_GRAPH, L.RT_FLAG_FRAME_GRAPH):
    try:
        ctx.render(cam, L.make_render_config(W, H, spp=2, max_depth=3, flags=flags)); ctx.sync()
        print(flags, "ok", ctx.stats()["lastRenderMs"], int(ctx.download(L.RT_BUF_RGBA8).astype(np.int64).sum()))
    except Exception as e:
        print(flags, "FAILED", e)
