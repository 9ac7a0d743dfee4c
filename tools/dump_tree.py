#!/usr/bin/env python3
"""Print the flattened tree of a scene blob (development aid)."""
import sys
from pathlib import Path
REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
import fast_ray_tracer_b200 as frt
d = frt.SceneDesc.load(sys.argv[1])
T = ["cone", "cube", "cyl", "plane", "stri", "sphere", "torus", "tri", "CSG", "GROUP"]
for i in range(d.c.n_nodes):
    n = d.c.nodes[i]
    print(i, T[n.type], "skip", n.skip, "parent", n.parent, "xf", n.xform, "mat", n.material, "op", n.csg_op, "right", n.right,
          [round(x, 3) for x in n.bbox_min], [round(x, 3) for x in n.bbox_max])
