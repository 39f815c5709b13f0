#!/usr/bin/env python
"""Development check (GPU box): CUDA path vs reference-kernel goldens, all samplers."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deepdrr_b200 import Projector, phantoms, geo
from deepdrr_b200.scene import SceneTables

class FixedProj:
    """A camera projection given directly by the kernel matrices stored in a golden file."""
    def __init__(self, w2i, src, ijk, W, H):
        self._w2i, self._src, self._ijk, self.W, self.H = w2i, src, ijk, W, H
        class K: pass
        self.intrinsic = K(); self.intrinsic.sensor_size = (W, H)

def rel(a, b, floor=1e-30):
    return np.abs(a - b) / np.maximum(np.abs(b), floor)

def check(name, volumes, spectrum, priorities=None, samplers=("alu", "tex", "hybrid")):
    g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    W, H, sub = int(g["W"]), int(g["H"]), int(g["sub"])
    for sampler in samplers:
        with Projector(volumes, priorities=priorities, spectrum=spectrum, step=float(g["step"]), neglog=False,
                       camera_intrinsics=geo.CameraIntrinsicTransform.from_sizes((W, H), 1.0, 1000.0), sampler=sampler) as p:
            i = 0
            while f"w2i_{i}" in g:
                arrs = (g[f"w2i_{i}"].reshape(1, 9), g[f"src_{i}"].reshape(1, -1, 3), g[f"ijk_{i}"].reshape(1, -1, 12))
                area = p.project_arrays(*arrs, (W, H), float(g["max_ray_length"]), want="area")[0]
                tm = p.last_timing_ms()
                img = p.project_arrays(*arrs, (W, H), float(g["max_ray_length"]), want="intensity", raw=True)[0]
                gl, gi = g[f"lineint_{i}"], g[f"intensity_{i}"]
                a = area[:, ::sub, ::sub]; im = img[::sub, ::sub]
                msg = f"{name} v{i} {sampler:6s} march {tm['march']:.2f} ms S={p.last_sample_count():.3e} I rel {rel(im, gi).max():.2e} |"
                for m in range(a.shape[0]):
                    strict = rel(a[m], gl[m])[gl[m] > 0]
                    msg += f" L{m} {strict.max() if strict.size else 0:.2e} eq {np.mean(a[m] == gl[m]):.2f}"
                print(msg, flush=True)
                i += 1

if __name__ == "__main__":
    v1 = phantoms.c1_volume()
    check("c1", [v1], "90KV_AL40")
    vs = phantoms.thorax_volume((128, 128, 100), (3.2, 3.2, 4.0))
    check("thorax_small", [vs], "120KV_AL43")
    w = phantoms.kwire_volume(length_mm=60.0, spacing=0.25, half_width=6); phantoms.place_kwire(w, (-10.0, -20.0, 0.0), (0.3, 1.0, 0.1))
    w2 = phantoms.kwire_volume(length_mm=60.0, spacing=0.25, half_width=6); phantoms.place_kwire(w2, (10.0, -20.0, 5.0), (-0.3, 1.0, 0.0))
    check("multivol3", [vs, w, w2], "90KV_AL40", samplers=("alu",))
    vs2 = phantoms.thorax_volume((96, 96, 80), (3.2, 3.2, 4.0), seed=3); vs2.translate((15.0, -10.0, 20.0))
    check("multivol2_sameprio", [vs, vs2], "60KV_AL35", priorities=[0, 0], samplers=("alu",))
    if "--c2" in sys.argv:
        v2 = phantoms.thorax_volume()
        check("c2", [v2], "120KV_AL43")
