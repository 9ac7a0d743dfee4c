import sys, time, json
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/oracle'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
import fast_ray_tracer_b200 as frt
from compare import parity_report
G='/root/repo/tests/golden/'
for name in ['cornell_gi_64','cornell_gi_caustics_48']:
    z=np.load(G+name+'.npz'); a=z['rgb'].astype(np.float64); b=z['rgb_b'].astype(np.float64)
    desc=frt.SceneDesc.load(G+name+'.frt')
    with frt.Scene(desc) as sc:
        t=time.time(); st=sc.trace_photons(3, bool(desc.config.gi_include_caustics), True, seed=7); tp=time.time()-t
        print(name,'photon pass',tp,'s',st.extra, sc.photons_count(0), sc.photons_count(1))
        t=time.time(); canvas,stats=sc.render(seed=3); print('render',time.time()-t, stats.frame_ms, stats.rays_gather, stats.kernel_launches)
    img=canvas[...,:3]
    print(' ours vs A',parity_report(img,a)['rmse_lsb'],' A vs B',parity_report(a,b)['rmse_lsb'],' means',img.mean(),a.mean(),b.mean())
    np.save('/root/repo/gpurun_out/'+name+'_gpu.npy', img.astype(np.float32))
