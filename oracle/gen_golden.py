"""oracle/gen_golden.py -- TEST INFRASTRUCTURE ONLY.

Generates tests/golden/*.npz by running the reference's UNMODIFIED sources
(oracle/_ref/libblah2ref.so, built by `make -C oracle` where /root/reference exists) on
small seeded inputs.  The fixtures are committed; tests compare both the numpy oracle and
the CUDA path against them, so parity stays pinned on machines without /root/reference.

    python oracle/gen_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import refpath as R  # noqa: E402
from blah2_b200.scene import make_scene, random_iq, Target  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def main():
    os.makedirs(OUT, exist_ok=True)
    # 1. next_hamming table
    v = np.array(sorted(set(list(range(0, 260)) + [3322, 6643, 15563, 19043, 19511, 39023, 39051, 65534])),
                 dtype=np.uint32)
    np.savez_compressed(os.path.join(OUT, "hamming.npz"), value=v,
                        next=np.array([R.next_hamming(int(k)) for k in v], dtype=np.uint32))

    # 2. constructor geometry
    geoms = [(-10, 300, -300, 300, 2000000, 1000000, 0), (-10, 300, -300, 300, 2000000, 1000000, 1),
             (0, 299, -128, 128, 2000000, 2000000, 1), (0, 511, -256, 256, 10000000, 20000000, 1),
             (0, 511, -512, 512, 10000000, 10000000, 1), (0, 511, -512, 512, 20000000, 80000000, 1),
             (-10, 400, -200, 200, 2000000, 1500000, 1), (0, 31, -20, 60, 10000, 5000, 1),
             (-3, 20, -50, 50, 10000, 4000, 0), (2, 40, -30, 30, 10000, 3000, 1)]
    rows = []
    for g in geoms:
        r = R.ambiguity_geometry(*g)
        rows.append(list(g) + [r["nDelayBins"], r["nDopplerBins"], r["nCorr"], r["nfft"], r["cpi"],
                               r["dopplerMiddle"]])
    np.savez_compressed(os.path.join(OUT, "geometry.npz"), rows=np.array(rows, dtype=np.float64))

    # 3. Ambiguity::process on small seeded inputs (random IQ like TestAmbiguity.cpp:24-32, seeded)
    for name, geom, seed in [("caf_a", (-3, 20, -50, 50, 10000, 4000, False), 1),
                             ("caf_b", (0, 31, -20, 60, 10000, 5000, True), 2),   # dopplerMiddle = 20 -> A2
                             ("caf_c", (1, 40, -30, 30, 10000, 3000, True), 3)]:  # delayMin = 1: largest the reference reads in bounds
        x, y = random_iq(geom[5], seed)
        r = R.ambiguity_process(x, y, *geom)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), geom=np.array(geom, dtype=np.int64), seed=seed,
                            map=r["map"], delay=r["delay"], doppler=r["doppler"],
                            metrics=np.array([r["noisePower"], r["maxPower"]]), leftover=np.array(r["leftover"]))

    # 4. WienerHopf::process
    for name, (n, dm, dM, seed) in [("wh_a", (5000, -3, 20, 1)), ("wh_b", (6007, 0, 33, 2))]:
        sc = make_scene(n, 2e6, seed=seed, targets=[Target(25, 300.0, -40.0)])
        ok, y = R.wienerhopf_process(sc.x, sc.y, dm, dM)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), params=np.array([n, dm, dM, seed]), ok=ok, y=y)

    # 5. whole chain (blah2.cpp:268-287) + per-stage detection lists on the reference map
    fs, n = 200000, 20000
    geom = (-5, 60, -500, 500, fs, n, True)
    det = dict(pfa=1e-4, nGuard=2, nTrain=6, minDelay=3, minDoppler=15.0, nCentroid=4)
    sc = make_scene(n, fs, seed=11, targets=[Target(17, 300.0, -25.0), Target(41, -200.0, -28.0)])
    ch = R.Chain(*geom[:6], roundHamming=True, clutter=(-5, 30), pfa=det["pfa"], nGuard=det["nGuard"],
                 nTrain=det["nTrain"], minDelay=det["minDelay"], minDoppler=det["minDoppler"],
                 nCentroid=det["nCentroid"])
    r = ch.run(sc.x, sc.y)
    g = R.ambiguity_geometry(*geom)
    ra = R.ambiguity_process(sc.x, R.wienerhopf_process(sc.x, sc.y, -5, 30)[1], *geom)
    d1 = R.cfar_1d(r["map"], ra["delay"], ra["doppler"], r["noisePower"], det["pfa"], det["nGuard"], det["nTrain"],
                   det["minDelay"], det["minDoppler"])
    d2 = R.centroid(*d1, det["nCentroid"], det["nCentroid"], 1.0 / (n / fs))
    d3 = R.interpolate(*d2, r["map"], ra["delay"], ra["doppler"], r["noisePower"], True, True)
    assert all(np.array_equal(a, b) for a, b in zip(d3, r["detections"]))
    np.savez_compressed(os.path.join(OUT, "chain_a.npz"), geom=np.array(geom, dtype=np.int64),
                        clutter=np.array([-5, 30]), seed=11,
                        det=np.array([det["pfa"], det["nGuard"], det["nTrain"], det["minDelay"], det["minDoppler"],
                                      det["nCentroid"]]),
                        map=r["map"], delay=ra["delay"], doppler=ra["doppler"],
                        metrics=np.array([r["noisePower"], r["maxPower"]]),
                        cfar=np.array(d1), centroid=np.array(d2), interp=np.array(d3))
    # 6. SpectrumAnalyser::process (blah2.cpp:263-265).  bandwidth 2000 is what blah2.cpp:198 hard-codes; the
    #    other cases exercise nfft < n, odd decimation, decimation 1 and a non-integer bandwidth.
    for name, (n, bw, seed) in [("spectrum_a", (6000, 2000.0, 1)), ("spectrum_b", (5003, 97.0, 2)),
                                ("spectrum_c", (3999, 2000.0, 3)), ("spectrum_d", (20000, 333.3, 4))]:
        x, _ = random_iq(n, seed)
        spec, freq, left = R.spectrum_process(x, n, bw)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), params=np.array([n, bw, seed], dtype=np.float64),
                            spectrum=spec, frequency=freq, leftover=left)
    print("golden fixtures written to", OUT, g)


if __name__ == "__main__":
    main()
