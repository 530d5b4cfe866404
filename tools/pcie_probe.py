import sys, time, ctypes as C
sys.path.insert(0, "/root/repo")
import numpy as np
from cortenn_b200 import capi
ctx = capi.Context(0)
n = 1 << 30
h = C.c_void_p(); capi.check(ctx.lib.llo_cuda_host_alloc(ctx.handle, n, C.byref(h)))
d = ctx.alloc(n)
C.memset(h, 1, n)
for name, fn in (("h2d", lambda: capi.check(ctx.lib.llo_cuda_h2d(ctx.handle, d, h, n))),
                 ("d2h", lambda: capi.check(ctx.lib.llo_cuda_d2h(ctx.handle, h, d, n)))):
    fn()
    t0 = time.perf_counter()
    for _ in range(5): fn()
    dt = (time.perf_counter() - t0) / 5
    print(name, "%.1f ms  %.1f GB/s" % (dt * 1e3, n / dt / 1e9))
pg = np.ones(n // 4, dtype=np.float32)
t0 = time.perf_counter(); capi.check(ctx.lib.llo_cuda_h2d(ctx.handle, d, pg.ctypes.data, n)); print("pageable h2d %.1f ms" % ((time.perf_counter() - t0) * 1e3))
