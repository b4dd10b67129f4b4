import sys,time,ctypes as C
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np, mmo_b200
from mmo_b200 import workloads, ScanParams, ScanResult
mmo_b200.init(0)
L=mmo_b200.lib()
c2=workloads.load_c2(); rl=workloads.lig_radius(c2['centered'])
rec_m=workloads.carve(c2['rec'],c2['roi'][:3],c2['roi'][3]+rl+12)
rec=mmo_b200.Receptor.from_mol(rec_m); lig=mmo_b200.Ligand.from_mol(c2['lig'])
rot=mmo_b200.SO3.rotations(100000)
P=ScanParams(); P.rec=rec.h; P.lig=lig.h; P.variant=1; P.prec=0
P.roi_c=(C.c_double*3)(*c2['roi'][:3]); P.roi_r=10.0; P.trans_step=1.0; P.n_rot=100000
P.rot9=rot.ctypes.data_as(C.POINTER(C.c_double)); P.topk=1000; P.first_point=3000; P.n_points=12
for it in range(3):
    t0=time.perf_counter(); job=C.c_void_p(); rc=L.mmo_scan_create(C.byref(P),0,C.byref(job)); t1=time.perf_counter()
    n=C.c_int64(); L.mmo_scan_num_points(job,C.byref(n))
    rc=L.mmo_scan_run(job,0,-1); t2=time.perf_counter()
    L.mmo_scan_destroy(job); t3=time.perf_counter()
    print(f'it{it}: points {n.value} create {1e3*(t1-t0):.1f} ms run {1e3*(t2-t1):.1f} ms destroy {1e3*(t3-t2):.1f} ms')
