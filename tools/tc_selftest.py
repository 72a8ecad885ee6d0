"""Dev/test helper (GPU): tcgen05 linear + wgrad kernels against torch matmul on the same fp16 operands."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from moda_b200._lib import call, ptr, stream

dev = "cuda"
torch.manual_seed(0)


def tc_linear(A1, A2, B, bias=None, rowbias=None, rep=1, relu=0, mask=None, rv=None, cv=None, rscale=None,
              y16=None, acc16=0, y32=None, oscale=None):
    M, K1 = A1.shape
    K2 = A2.shape[1] if A2 is not None else 0
    N = B.shape[0]
    call("moda_tc_linear", ptr(A1), A1.stride(0), K1, ptr(A2) if A2 is not None else None,
         A2.stride(0) if A2 is not None else 0, K2, ptr(B), B.stride(0), M, N, ptr(bias), ptr(rowbias), rep, relu,
         ptr(mask), mask.stride(0) if mask is not None else 0, ptr(rv), ptr(cv), ptr(rscale),
         ptr(y16), y16.stride(0) if y16 is not None else 0, acc16, ptr(y32), y32.stride(0) if y32 is not None else 0,
         ptr(oscale), stream())


def check(name, got, ref, tol):
    err = float((got.double() - ref.double()).abs().max())
    scale = float(ref.abs().max())
    ok = err <= tol * max(scale, 1.0)
    print("%-46s max|err| %.3e (scale %.2e) %s" % (name, err, scale, "ok" if ok else "FAIL"))
    return ok


ok = True
for (M, N, K1, K2) in [(1000, 256, 256, 0), (333, 256, 64, 256), (4096, 128, 256, 0), (129, 64, 256, 0), (1000, 256, 64, 0), (640, 256, 128, 0)]:
    A1 = (torch.randn(M, K1, device=dev) * 0.5).half()
    A2 = (torch.randn(M, K2, device=dev) * 0.5).half() if K2 else None
    B = (torch.randn(N, K1 + K2, device=dev) / (K1 + K2) ** 0.5).half()
    A = torch.cat([A1, A2], 1) if K2 else A1
    ref = A.float() @ B.float().t()
    y32 = torch.full((M, N), float("nan"), device=dev)
    y16 = torch.zeros(M, N, device=dev, dtype=torch.float16)
    tc_linear(A1, A2, B, y16=y16, y32=y32)
    torch.cuda.synchronize()
    ok &= check("linear M%d N%d K%d+%d fp32 out" % (M, N, K1, K2), y32, ref, 2e-5)
    ok &= check("linear M%d N%d K%d+%d fp16 out" % (M, N, K1, K2), y16, ref, 2e-3)

# fused epilogue options
M, N, K = 777, 256, 256
A1 = (torch.randn(M, K, device=dev) * 0.5).half()
B = (torch.randn(N, K, device=dev) / K ** 0.5).half()
bias = torch.randn(N, device=dev)
rep = 7
rowbias = torch.randn((M + rep - 1) // rep, N, device=dev)
ref = A1.float() @ B.float().t() + bias + rowbias.repeat_interleave(rep, 0)[:M]
y32 = torch.zeros(M, N, device=dev)
tc_linear(A1, None, B, bias=bias, rowbias=rowbias, rep=rep, relu=1, y32=y32)
ok &= check("bias + rowbias + relu", y32, torch.relu(ref), 2e-5)
mask = torch.randn(M, N, device=dev).half()
rv, cv = torch.randn(M, device=dev), torch.randn(N, device=dev)
rs, osc = torch.tensor([4.0], device=dev), torch.tensor([0.25], device=dev)
ref = (A1.float() @ B.float().t() + 4.0 * rv[:, None] * cv[None]) * (mask.float() > 0)
y32 = torch.zeros(M, N, device=dev)
y16 = torch.ones(M, N, device=dev, dtype=torch.float16)
tc_linear(A1, None, B, mask=mask, rv=rv, cv=cv, rscale=rs, y16=y16, acc16=1, y32=y32, oscale=osc)
ok &= check("rank-1 + mask + oscale", y32, 0.25 * ref, 2e-5)
ok &= check("acc16", y16, ref + 1.0, 3e-3)

# wgrad
for (M, N, K) in [(1000, 256, 256), (64, 256, 64), (5000, 128, 256), (100000, 256, 256), (777, 256, 128)]:
    dY = (torch.randn(M, N, device=dev) * 0.1).half()
    X = (torch.randn(M, K, device=dev) * 0.5).half()
    dW = torch.zeros(N, K, device=dev)
    db = torch.zeros(N, device=dev)
    osc = torch.tensor([0.5], device=dev)
    call("moda_tc_wgrad", ptr(dY), N, N, ptr(X), K, K, M, ptr(dW), K, N, K, ptr(osc), ptr(db), stream())
    ref = 0.5 * (dY.float().t() @ X.float())
    ok &= check("wgrad M%d N%d K%d" % (M, N, K), dW, ref, 1e-4)
    ok &= check("wgrad bias M%d N%d K%d" % (M, N, K), db, 0.5 * dY.float().sum(0), 1e-4)
for (M, N, K, nv, kv) in [(1000, 64, 64, 64, 64), (5000, 64, 64, 25, 32), (3000, 64, 128, 64, 128)]:
    dY = (torch.randn(M, N, device=dev) * 0.1).half()
    X = (torch.randn(M, K, device=dev) * 0.5).half()
    dW = torch.zeros(N, K, device=dev)
    db = torch.zeros(N, device=dev)
    call("moda_tc_wgrad", ptr(dY), N, N, ptr(X), K, K, M, ptr(dW), K, nv, kv, None, ptr(db), stream())
    ref = dY.float().t() @ X.float()
    ref[nv:] = 0
    ref[:, kv:] = 0
    rb = dY.float().sum(0)
    rb[nv:] = 0
    ok &= check("wgrad64 M%d N%d K%d valid %dx%d" % (M, N, K, nv, kv), dW, ref, 1e-4)
    ok &= check("wgrad64 bias", db, rb, 1e-4)

# split precision (hi/lo fp16 pairs): fp32-class accuracy
def split(x):
    hi = x.half()
    lo = (x - hi.float()).half()
    return hi.contiguous(), lo.contiguous()

for (M, N, K1, K2) in [(1000, 64, 64, 0), (777, 64, 64, 64), (300, 32, 64, 0)]:
    X1 = torch.randn(M, K1, device=dev) * 0.5
    X2 = torch.randn(M, K2, device=dev) * 0.5 if K2 else None
    W = torch.randn(N, K1 + K2, device=dev) / (K1 + K2) ** 0.5
    x1h, x1l = split(X1)
    x2h, x2l = split(X2) if K2 else (None, None)
    wh, wl = split(W)
    B3 = torch.cat([wh, wh, wl], 1).contiguous()
    bias = torch.randn(N, device=dev)
    X = torch.cat([X1, X2], 1) if K2 else X1
    ref = torch.relu(X.double() @ W.double().t() + bias.double())
    yh = torch.zeros(M, N, device=dev, dtype=torch.float16)
    yl = torch.zeros(M, N, device=dev, dtype=torch.float16)
    y32 = torch.zeros(M, N, device=dev)
    call("moda_tc_linear_split", ptr(x1h), ptr(x1l), K1, K1, ptr(x2h) if K2 else None, ptr(x2l) if K2 else None, K2, K2,
         ptr(B3), B3.stride(0), M, N, ptr(bias), None, 1, 1, None, 0, ptr(yh), ptr(yl), N, 0, ptr(y32), N, None, stream())
    ok &= check("split linear M%d N%d K%d+%d fp32 out" % (M, N, K1, K2), y32.double(), ref, 3e-6)
    ok &= check("split linear M%d N%d K%d+%d hi+lo out" % (M, N, K1, K2), yh.double() + yl.double(), ref, 3e-6)
for (M, N, K, nv, kv) in [(1000, 64, 64, 64, 64), (5000, 64, 64, 25, 32), (3000, 64, 128, 64, 128)]:
    dY = torch.randn(M, N, device=dev) * 0.1
    X = torch.randn(M, K, device=dev) * 0.5
    yh, yl = split(dY)
    xh, xl = split(X)
    dW = torch.zeros(N, K, device=dev)
    db = torch.zeros(N, device=dev)
    call("moda_tc_wgrad_split", ptr(yh), ptr(yl), N, N, ptr(xh), ptr(xl), K, K, M, ptr(dW), K, nv, kv, None, ptr(db), stream())
    ref = (dY.double().t() @ X.double())
    ref[nv:] = 0
    ref[:, kv:] = 0
    rb = dY.double().sum(0)
    rb[nv:] = 0
    ok &= check("split wgrad M%d N%d K%d valid %dx%d" % (M, N, K, nv, kv), dW.double(), ref, 3e-6)
    ok &= check("split wgrad bias", db.double(), rb, 3e-6)

print("ALL OK" if ok else "SOME FAILED")

# timing at the training size
M = 8192 * 128
A1 = (torch.randn(M, 256, device=dev) * 0.5).half()
B = (torch.randn(256, 256, device=dev) / 16).half()
y16 = torch.empty(M, 256, device=dev, dtype=torch.float16)
bias = torch.zeros(256, device=dev)
dW = torch.zeros(256, 256, device=dev)
def t(fn, n=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
ms = t(lambda: tc_linear(A1, None, B, bias=bias, relu=1, y16=y16))
print("tc_linear 1M x 256 x 256: %.3f ms  %.1f TFLOP/s  %.1f GB/s" % (ms, 2 * M * 256 * 256 / ms / 1e9, 2 * M * 512 / ms / 1e6))
ms = t(lambda: tc_linear(A1, None, B, mask=A1, y16=y16))
print("tc_linear (dgrad+mask)  : %.3f ms  %.1f TFLOP/s" % (ms, 2 * M * 256 * 256 / ms / 1e9))
ms = t(lambda: call("moda_tc_wgrad", ptr(y16), 256, 256, ptr(A1), 256, 256, M, ptr(dW), 256, 256, 256, None, ptr(bias), stream()))
print("tc_wgrad  1M x 256 x 256: %.3f ms  %.1f TFLOP/s" % (ms, 2 * M * 256 * 256 / ms / 1e9))
ms = t(lambda: torch.matmul(A1, B.t()))
print("torch fp16 matmul       : %.3f ms  %.1f TFLOP/s" % (ms, 2 * M * 256 * 256 / ms / 1e9))
