import sys, json
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/oracle'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, fast_ray_tracer_b200 as frt
from compare import parity_report
from conftest import golden_names, GOLDEN
for n in golden_names():
    z=np.load(GOLDEN/f'{n}.npz'); d=frt.SceneDesc.load(GOLDEN/f'{n}.frt')
    c,st=frt.render_multi(d)
    r=parity_report(c[...,:3], z['rgb'].astype(float))
    print(f"{n:28s} within1={r['within_1lsb']:.5f} exact={r['exact']:.4f} max={r['max_lsb']} bad={r['bad_pixels']}")
