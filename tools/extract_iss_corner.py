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

def main():
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
    import __graft_entry__ as entry
    out = {"source": "StanfordASL/GuSTO.jl src/environment/iss_corner.mat (float32-rounded as in iss_corner.jl)"}
    out.update(entry.load_package().trajio.load_iss_corner_mat(SRC))       # the run-time reader is the one code path
    with open(DST, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", os.path.abspath(DST), {k: len(v) for k, v in out.items() if isinstance(v, list)})

if __name__ == "__main__":
    main()
