"""Per-kernel-class device profile (CUDA events) of single primitives: qr / svd on given shapes."""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tensor4all-rs_b200", "python"))
import t4b
ctx = t4b.Context(0)
rng = np.random.default_rng(0)
def prof(fn):
    fn(); ctx.sync()
    t4b._check(t4b.lib().t4b_ctx_profile_begin(ctx.h))
    fn()
    need = C.c_size_t()
    t4b._check(t4b.lib().t4b_ctx_profile_end(ctx.h, None, C.c_size_t(0), C.byref(need)))
    buf = C.create_string_buffer(need.value)
    t4b._check(t4b.lib().t4b_ctx_profile_end(ctx.h, buf, need, None))
    return buf.value.decode()
for arg in sys.argv[1:]:
    op, shp = arg.split(":")
    m, n = [int(x) for x in shp.split("x")]
    a = ctx.upload(rng.standard_normal((m, n)))
    if op == "qr":
        f = lambda: ctx.qr_thin(ctx.permute(a, [0, 1]))
    elif op == "qrr":
        f = lambda: ctx.qr_thin(ctx.permute(a, [0, 1]), want_q=False)
    else:
        f = lambda: ctx.svd_thin(ctx.permute(a, [0, 1]), want_u=True, want_vh=False)
    print("==", arg)
    for ln in prof(f).splitlines():
        nm, cnt, ms, work = ln.split()
        print(f"   {nm:20s} n={int(cnt):5d} total={float(ms):9.3f} ms  avg={1e3*float(ms)/int(cnt):8.1f} us")
