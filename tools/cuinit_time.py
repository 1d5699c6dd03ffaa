import ctypes, time, os, sys
t=time.time(); l=ctypes.CDLL("libcuda.so.1"); rc=l.cuInit(0); n=ctypes.c_int(); l.cuDeviceGetCount(ctypes.byref(n)); print(os.environ.get("CUDA_VISIBLE_DEVICES"), "cuInit rc", rc, "devices", n.value, "%.2f s" % (time.time()-t))
