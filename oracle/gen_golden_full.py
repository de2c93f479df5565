"""oracle/gen_golden_full.py -- TEST INFRASTRUCTURE ONLY.

Full-size fixtures of the BASELINE.json configurations, produced by the reference's UNMODIFIED sources
(oracle/_ref/libblah2ref.so) on seeded inputs that the tests regenerate (blah2_b200/scene.py):

  full_cfg3.npz   Ambiguity::process, 2 s CPI @ 10 MS/s, 512 x 1025          (BASELINE configs[2])
  full_cfg4.npz   Ambiguity::process, 1 s CPI @ 10 MS/s, 512 x 1025          (configs[3], one CPI of the stream)
  full_cfg5.npz   Ambiguity::process, 4 s CPI @ 20 MS/s, 512 x 4097          (configs[4]; --with-cfg5, ~4 GB of RAM)
  full_cfg2.npz   WienerHopf + Ambiguity + set_metrics + CFAR/Centroid/Interpolate at N = 2e6  (configs[1])

A whole map would be megabytes per fixture, so each file stores a strided subsample of the complex128 map (every
STEP-th row and column), its Frobenius norm, and a seeded random projection of ALL cells (one complex number that
moves if any cell does); the tests check the CUDA map's subsample, norm and projection to the north star's 1e-5.

    python oracle/gen_golden_full.py [--with-cfg5]
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import refpath as R  # noqa: E402
from blah2_b200.scene import make_scene, random_iq  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
STEP = 8

CAF_CASES = {
    "full_cfg3": ((0, 511, -256, 256, 10000000, 20000000, True), 303),
    "full_cfg4": ((0, 511, -512, 512, 10000000, 10000000, True), 304),
    "full_cfg5": ((0, 511, -512, 512, 20000000, 80000000, True), 305),
}
CHAIN_CFG2 = dict(geom=(0, 299, -128, 128, 2000000, 2000000, True), clutter=(-10, 400), seed=20260923,
                  det=dict(pfa=1e-5, nGuard=2, nTrain=6, minDelay=5, minDoppler=15.0, nCentroid=6))


def projection(shape, seed=7):
    """Unit-modulus pseudo-random weights over all cells (seeded; the tests rebuild them)."""
    rng = np.random.default_rng(seed)
    return np.exp(2j * np.pi * rng.random(shape))


def summarise(m):
    return dict(sub=m[::STEP, ::STEP].copy(), fro=float(np.linalg.norm(m)), amax=float(np.max(np.abs(m))),
                proj=np.sum(m * projection(m.shape)))


def main():
    os.makedirs(OUT, exist_ok=True)
    for name, (geom, seed) in CAF_CASES.items():
        if name == "full_cfg5" and "--with-cfg5" not in sys.argv:
            continue
        t0 = time.time()
        x, y = random_iq(geom[5], seed)
        r = R.ambiguity_process(x, y, *geom)
        s = summarise(r["map"])
        np.savez_compressed(os.path.join(OUT, name + ".npz"), geom=np.array(geom, dtype=np.int64), seed=seed, step=STEP,
                            metrics=np.array([r["noisePower"], r["maxPower"]]), leftover=np.array(r["leftover"]), **s)
        print(name, "done in %.1f s" % (time.time() - t0), flush=True)
        del x, y, r
    c = CHAIN_CFG2
    geom, det = c["geom"], c["det"]
    sc = make_scene(geom[5], geom[4], seed=c["seed"])
    ch = R.Chain(*geom[:6], roundHamming=True, clutter=c["clutter"], **det)
    r = ch.run(sc.x, sc.y)
    assert not r["skipped"]
    s = summarise(r["map"])
    np.savez_compressed(os.path.join(OUT, "full_cfg2.npz"), geom=np.array(geom, dtype=np.int64), seed=c["seed"], step=STEP,
                        clutter=np.array(c["clutter"]),
                        det=np.array([det["pfa"], det["nGuard"], det["nTrain"], det["minDelay"], det["minDoppler"],
                                      det["nCentroid"]]),
                        metrics=np.array([r["noisePower"], r["maxPower"]]), detections=np.array(r["detections"]), **s)
    print("full_cfg2 done:", len(r["detections"][0]), "detections")


if __name__ == "__main__":
    main()
