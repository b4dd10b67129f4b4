import ctypes as C, time, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import mmo_b200
mmo_b200.init(0)
L = mmo_b200.lib()
d = C.c_void_p()
L.mmo_dev_alloc(C.c_size_t(1 << 20), C.byref(d))
h = np.zeros(1 << 17, np.float64)
hp = C.c_void_p()
L.mmo_host_alloc(C.c_size_t(1 << 20), C.byref(hp))
def t(fn, n=2000):
    for _ in range(50): fn()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    return 1e6 * (time.perf_counter() - t0) / n
print("mmo_sync (idle stream)        %.1f us" % t(lambda: L.mmo_sync()))
for nb in (8, 4096, 51200):
    print("d2h pageable %6d B + sync   %.1f us" % (nb, t(lambda: L.mmo_d2h(h.ctypes.data_as(C.c_void_p), d, C.c_size_t(nb)))))
    print("d2h pinned   %6d B + sync   %.1f us" % (nb, t(lambda: L.mmo_d2h(hp, d, C.c_size_t(nb)))))
    print("h2d pageable %6d B + sync   %.1f us" % (nb, t(lambda: L.mmo_h2d(d, h.ctypes.data_as(C.c_void_p), C.c_size_t(nb)))))
    print("h2d pinned   %6d B + sync   %.1f us" % (nb, t(lambda: L.mmo_h2d(d, hp, C.c_size_t(nb)))))
ms = C.c_float()
def timed_empty():
    L.mmo_timer_start(); L.mmo_timer_stop(C.byref(ms))
print("timer start/stop              %.1f us (device %.1f us)" % (t(timed_empty), ms.value * 1e3))
