"""Analysis tool (CPU, host simulator): k_extend's warp loop on simulated 32-lane warps over the rays of a C4 frame - lane
participation and warp instructions per ray of the shipped schedule (to compare with the measured RT_PHASE_STATS numbers) and of
alternatives.  python tests/cpu_simt_schedule.py"""
import sys

sys.path.insert(0, ".")
from ilgpu_raytracing_b200 import layouts as L, scenes  # noqa: E402
from tests.hostsim_binding import HostSimScene, capture_rays, simulate  # noqa: E402
from tests.util import oracle_camera, oracle_scene_from_spec  # noqa: E402

W, H, spp = 384, 216, 8
hs = HostSimScene(oracle_scene_from_spec(scenes.terrain_scene(n_quads=708, n_spheres=256)).arrays())
capture_rays(True)
hs.render(oracle_camera("C3", W, H), L.make_render_config(W, H, spp=spp, max_depth=8), aovs=False)
capture_rays(False)
for any_hit in (False, True):
    print("any-hit" if any_hit else "closest")
    for label, kw in (("shipped: 2 node steps + voted primitive step", dict(policy=0, node_steps=2, prim_vote=1)),
                      ("1 node step", dict(policy=0, node_steps=1)), ("3 node steps", dict(policy=0, node_steps=3)),
                      ("2 node steps, vote 12", dict(policy=0, node_steps=2, prim_vote=12)),
                      ("one phase per iteration, the fuller one", dict(policy=1)),
                      ("shipped + 50 % of the failing candidates culled for free", dict(pre_cull=0.5)),
                      ("shipped + 80 % of the failing candidates culled for free", dict(pre_cull=0.8)),
                      ("shipped + 80 % culled, node step 20 instructions longer", dict(pre_cull=0.8, c_node=270.0)),
                      ("shipped + candidates culled by their OWN box, free", dict(pre_cull=-1.0)),
                      ("shipped + own-box cull, node step 30 instructions longer", dict(pre_cull=-1.0, c_node=280.0)),
                      ("two-stage: 45-instruction test, exact test at >= 1 survivor", dict(policy=2, pre_cull=-45.0, prim_vote=1)),
                      ("two-stage: 45-instruction test, exact test at >= 4 survivors", dict(policy=2, pre_cull=-45.0, prim_vote=4)),
                      ("two-stage: 45-instruction test, exact test at >= 8 survivors", dict(policy=2, pre_cull=-45.0, prim_vote=8)),
                      ("two-stage: 70-instruction test, exact test at >= 4 survivors", dict(policy=2, pre_cull=-70.0, prim_vote=4))):
        r = simulate(hs, any_hit, warps=128, **kw)
        print(f"  {label:48s} {r['warp_instr_per_ray']:7.1f} warp instr / ray, {r['iterations'] / r['rays']:.3f} iterations / ray, "
              f"{r['lanes_per_node_phase']:.1f} lanes / node phase, {r['lanes_per_prim_phase']:.1f} lanes / primitive phase, "
              f"{r['prim_steps'] / r['rays']:.2f} exact tests / ray of which {r['accepted'] / max(1.0, r['prim_steps']) * 100:.0f} % accept", flush=True)
