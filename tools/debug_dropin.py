import sys, os, faulthandler, importlib.util
faulthandler.enable()
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import refpath as R, blah2_oracle as O
from blah2_b200.scene import make_scene, Target
spec = importlib.util.spec_from_file_location("dropin_binding", R.__file__)
D = importlib.util.module_from_spec(spec); spec.loader.exec_module(D)
D.LIB_PATH = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests/native/_build/libdropin_harness.so")
fs, n = 2000000, 200000
sc = make_scene(n, fs, seed=5, targets=[Target(37, 3000.0, -30.0), Target(92, -2000.0, -35.0)])
g = O.ambiguity_geometry(-10, 120, -5000, 5000, fs, n, True)
ok, yf = D.wienerhopf_process(sc.x, sc.y, -10, 60); print("wh ok", ok, flush=True)
a = D.ambiguity_process(sc.x, yf, -10, 120, -5000, 5000, fs, n, True); print("amb ok", a["noisePower"], flush=True)
d1 = D.cfar_1d(a["map"], a["delay"], a["doppler"], a["noisePower"], 1e-5, 2, 6, 5, 15.0); print("cfar ok", len(d1[0]), flush=True)
d2 = D.centroid(*d1, 6, 6, 1.0 / (n / fs)); print("centroid ok", len(d2[0]), flush=True)
d3 = D.interpolate(*d2, a["map"], a["delay"], a["doppler"], a["noisePower"]); print("interp ok", len(d3[0]), flush=True)
ch = D.Chain(-10, 120, -5000, 5000, fs, n, True, clutter=(-10, 60)); print("chain created", flush=True)
r = ch.run(sc.x, sc.y); print("chain ok", r["stage_ms"], len(r["detections"][0]), flush=True)
