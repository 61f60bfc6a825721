import sys, time, numpy as np
import os; sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'nubomedia-vca_b200', 'python'))
import nubovca as nv
c=nv.Context(0,1920,1080)
casc=nv.Cascade('haarcascade_eye.xml')
rng=np.random.default_rng(0)
a=rng.integers(0,256,(42,60),dtype=np.uint8); b=rng.integers(0,256,(50,72),dtype=np.uint8)
for name,seq in [('same size',[a,a]),('alternating',[a,b])]:
    for _ in range(20): 
        for im in seq: c.detect_multiscale(casc,im,1.1,2,(20,20))
    t=time.perf_counter(); n=300
    for _ in range(n):
        for im in seq: c.detect_multiscale(casc,im,1.1,2,(20,20))
    print(name, (time.perf_counter()-t)/(2*n)*1e6,'us per detect')
