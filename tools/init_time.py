import time, ctypes, os
t0=time.time(); L=ctypes.CDLL("/root/repo/faldoi-ipol_b200/libfaldoi_gpu.so")
rt=ctypes.CDLL("libcudart.so.12") if False else None
t=time.time(); n=L.faldoi_device_count(); print("device_count %.3f"%(time.time()-t))
L.faldoi_solver_create.argtypes=[ctypes.POINTER(ctypes.c_void_p)]+[ctypes.c_int]*5
for k in range(3):
    h=ctypes.c_void_p(); t=time.time(); rc=L.faldoi_solver_create(ctypes.byref(h),0,1024,436,0,1); print("create#%d rc=%d %.3f"%(k,rc,time.time()-t))
    L.faldoi_solver_destroy.argtypes=[ctypes.c_void_p]; t=time.time(); L.faldoi_solver_destroy(h); print("destroy %.3f"%(time.time()-t))
