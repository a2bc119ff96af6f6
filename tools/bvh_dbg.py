import time, os
os.environ["POLARIS_BVH_DEBUG"]="1"
from polaris_b200 import scenes
from polaris_b200.scene import compile_scene
raw = scenes.raw_scene("c4_terrain")
for b in ("cuda", "cuda"):
    t = time.time(); sc = compile_scene(raw, 16 / 9, builder=b); dt = time.time() - t
    print("c4", b, "compile_scene %.2fs" % dt, {k: (round(v, 3) if isinstance(v, float) else v) for k, v in sc.compile_timing.items()}, flush=True)
import cProfile, pstats
pr=cProfile.Profile(); pr.enable(); sc = compile_scene(raw, 16 / 9, builder="cuda"); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(12)
