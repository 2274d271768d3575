import sys, time
sys.path.insert(0,'/root/repo')
import numpy as np
from oracle import xy_oracle as oc
from scpn_quantum_control_b200.engine import XYKEngine
for n in (3, 9, 14, 20, 30):
    K,w=oc.random_problem(n, 60+n)
    with XYKEngine(n) as eng:
        eng.set_problem(K,w); eng.init_product_state()
        if n>=12: eng.set_engine("tiled")
        eng.apply_layers(0.05,1,1)
        z=eng.z_expectations(); eng.synchronize()
        t=time.perf_counter(); z=eng.z_expectations(); dt=time.perf_counter()-t
        nrm=eng.norm2()
        if n<=20:
            psi=eng.get_state(); ref=oc.z_expectations(psi,n); err=np.abs(z-ref).max()
        else:
            err=abs(z.sum()-np.cos(np.mod(w,2*np.pi)).sum())
        print(n, 'z err', err, 'norm', nrm, 'ms', round(1e3*dt,3))
