"""Where the start-up of a drop-in invocation goes: CUDA primary context vs hmmcu_create vs the first kernels."""
import ctypes as C, time, os, sys
t0 = time.perf_counter()
rt = C.CDLL("libcudart.so.12")
t1 = time.perf_counter()
rt.cudaSetDevice(0); rt.cudaFree(None)
t2 = time.perf_counter()
lib = C.CDLL(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "speech_recognition_hmm_continuous_b200", "libhmmcu.so"))
t3 = time.perf_counter()
h = C.c_void_p()
lib.hmmcu_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
rc = lib.hmmcu_create(0, C.byref(h))
t4 = time.perf_counter()
lib.hmmcu_destroy.argtypes = [C.c_void_p]
lib.hmmcu_destroy(h)
t5 = time.perf_counter()
print("dlopen cudart %.0f ms | primary context %.0f ms | dlopen libhmmcu %.0f ms | hmmcu_create %.0f ms (rc %d) | destroy %.0f ms" %
      ((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, (t4 - t3) * 1e3, rc, (t5 - t4) * 1e3))
