"""indices beyond 2^31 in forall + reducers: ReduceSum / MinLoc / MaxLoc over 2^31 + 1029 int32 elements with the extrema planted
past index 2^31 (first-index tie rule included), and triad over 2^31 + 5 float32 elements checked in chunks."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import raja_b200 as rb
n = (1 << 31) + 1029
g = torch.Generator(device="cuda").manual_seed(11)
x = (torch.randint(0, 4, (n,), dtype=torch.int8, device="cuda", generator=g) == 0).to(torch.int32)
x[(1 << 31) + 5] = -7; x[(1 << 31) + 900] = -7      # duplicate minimum: the first one wins
x[(1 << 31) + 17] = 9; x[(1 << 31) + 18] = 9
total = int(x.sum(dtype=torch.int64).item())
s = rb.reduce_sum(x)
mn = rb.reduce_minloc(x)
mx = rb.reduce_maxloc(x)
ok = int(s) == total and tuple(int(t) for t in mn) == (-7, (1 << 31) + 5) and tuple(int(t) for t in mx) == (9, (1 << 31) + 17)
print("sum", int(s), total, "minloc", mn, "maxloc", mx, "->", ok)
del x
m = (1 << 31) + 5
b = torch.rand(m, dtype=torch.float32, device="cuda", generator=g); c = torch.rand(m, dtype=torch.float32, device="cuda", generator=g)
a = torch.empty_like(b)
rb.forall_triad(a, b, c, 3.0)
torch.cuda.synchronize()
ok2 = True
CH = 1 << 28
for lo in range(0, m, CH):
    hi = min(m, lo + CH)
    ok2 = ok2 and bool((a[lo:hi] == torch.addcmul(b[lo:hi], c[lo:hi], torch.tensor(3.0, device="cuda"))).all().item() or
                       torch.allclose(a[lo:hi], b[lo:hi] + 3.0 * c[lo:hi], rtol=1e-6, atol=0))
print("triad:", ok2)
print("BIG REDUCE OK" if ok and ok2 else "BIG REDUCE FAILED")
