#!/usr/bin/env python
"""Extract the ISS-corner geometry used by the reference's ISSCorner() environment into JSON.

Reads /root/reference/src/environment/iss_corner.mat (MATLAB v5 data blob; only readable in the
build container) and writes gusto.jl_b200/data/iss_corner.json.  Geometry is rounded through float32
exactly as the reference does when it stores corners as Vec3f0
(/root/reference/src/environment/iss_corner.jl:16,22,56,61): HyperRectangle(origin=f32(c1),
widths=f32(c2-c1)); max corner = f32(origin)+f32(widths) evaluated in float32 (GeometryTypes
`maximum`), then widened to float64.
"""
import json, os, sys
import numpy as np
import scipy.io as sio

SRC = "/root/reference/src/environment/iss_corner.mat"
DST = os.path.join(os.path.dirname(__file__), "..", "gusto.jl_b200", "data", "iss_corner.json")

def box(z):
    c1 = np.asarray(z.corner1, dtype=np.float64)
    c2 = np.asarray(z.corner2, dtype=np.float64)
    origin = c1.astype(np.float32)
    widths = (c2 - c1).astype(np.float32)
    hi = (origin + widths).astype(np.float32)
    lo = np.minimum(origin, hi).astype(np.float64)
    hi = np.maximum(origin, hi).astype(np.float64)
    return {"lo": lo.tolist(), "hi": hi.tolist()}

def main():
    m = sio.loadmat(SRC, squeeze_me=True, struct_as_record=False)
    out = {
        "source": "StanfordASL/GuSTO.jl src/environment/iss_corner.mat (float32-rounded as in iss_corner.jl)",
        "keepin_zones": [box(z) for z in np.atleast_1d(m["keepin_zones"])],
        "keepout_zones": [box(z) for z in np.atleast_1d(m["keepout_zones"])],
        "obstacle_rectangles": [box(z) for z in np.atleast_1d(m["rectangles"])],
        "obstacle_spheres": [
            {"center": np.asarray(z.center, dtype=np.float32).astype(np.float64).tolist(),
             "radius": float(np.float32(z.radius))}
            for z in np.atleast_1d(m["spheres"])],
    }
    with open(DST, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", os.path.abspath(DST), {k: len(v) for k, v in out.items() if isinstance(v, list)})

if __name__ == "__main__":
    main()
