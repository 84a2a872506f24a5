"""Tumbler(500): time-averaged pile statistics (steps 700..1000, every 10th) of the GPU path against the reference;
a snapshot of an avalanching pile is one sample of a chaotic system, the average is what can be compared."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from box2d_optimized_b200 import GpuScene
from oracle.bindings import RefScene


def stats(scene):
    scene.step(700)
    my, mx, hist, pen = [], [], np.zeros(10), []
    bins = np.linspace(0.0, 20.0, 11)
    for _ in range(30):
        scene.step(10)
        b = scene.bodies()[2:]
        my.append(b[:, 5].mean()); mx.append(b[:, 4].mean())
        hist += np.histogram(b[:, 5], bins)[0]
    return dict(mean_y=float(np.mean(my)), mean_x=float(np.mean(mx)), std_mean_y=float(np.std(my)), hist=(hist / 30).round(1).tolist(),
                snapshot_mean_y=float(my[-1]), contacts=int(scene.contact_count))


r = stats(RefScene("tumbler", 500, 0))
g = stats(GpuScene("tumbler", 500, 0))
print("ref", json.dumps(r)); print("gpu", json.dumps(g))
print("L1(hist)/500 = %.3f  d(mean_y) = %.3f  d(mean_x) = %.3f" % (np.abs(np.array(r["hist"]) - np.array(g["hist"])).sum() / 500,
                                                                  g["mean_y"] - r["mean_y"], g["mean_x"] - r["mean_x"]))
