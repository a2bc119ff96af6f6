import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from polaris_b200 import _lib, scenes, tracer as T
w, h = 3840, 2160
bh = int(sys.argv[1]) if len(sys.argv) > 1 else 1
sc, _, _, _ = scenes.build("c5_cornell_4k", w, h)
tr = T.CudaTracer("cuda:0", 0); tr.init()
tr.set_option(_lib.OPT_SAMPLE_CHAINS, 1)
tr.set_option(_lib.OPT_USE_GRAPH, 0)
tr.update_state(T.SYNCHRONOUS, T.FRAME_DIMENSIONS, (w, h))
tr.update_state(T.SYNCHRONOUS, T.SCENE_DATA, sc)
tr.update_state(T.SYNCHRONOUS, T.CAMERA_DATA, sc.camera)
for rep in range(3):
    tr.trace(T.make_block_request(w, h, block_y=1000, block_h=bh, spp=2), T.splitmix_seeds(5, 12))
print(tr.stats().device["query_rays"])
