"""polaris_b200 -- host side of the B200 (sm_100a) tracer backend for polaris.

Only what the `tracer.Tracer` hot path needs: the ctypes binding of libpolaris_cuda.so and a
Python mirror of the reference's tracer / scheduler / renderer interfaces (tracer.py,
scheduler.py, renderer.py), plus the input producers the reference implements in Go
(scene.py, material.py, gotypes.py, scenes.py).
"""
__version__ = "0.1.0"
