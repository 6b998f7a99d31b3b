"""indices beyond 2^31: inclusive_scan / exclusive_scan of 2^31 + 7 int32 elements; checks out[i] - out[i-1] == in[i] everywhere,
out[0], and the last element against an independent 64-bit sum."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import raja_b200 as rb
n = (1 << 31) + 7
g = torch.Generator(device="cuda").manual_seed(9)
x = (torch.randint(0, 4, (n,), dtype=torch.int8, device="cuda", generator=g) == 0).to(torch.int32)  # ~2^29 ones
total = int(x.sum(dtype=torch.int64).item())
out = torch.empty_like(x)
rb.inclusive_scan(x, out)
torch.cuda.synchronize()
ok = int(out[-1].item()) == total and int(out[0].item()) == int(x[0].item())
CH = 1 << 28
for a in range(1, n, CH):
    b = min(n, a + CH)
    ok = ok and bool((out[a:b] - out[a - 1:b - 1] == x[a:b]).all().item())
print("inclusive:", ok)
rb.exclusive_scan(x, out)
torch.cuda.synchronize()
ok2 = int(out[0].item()) == 0 and int(out[-1].item()) == total - int(x[-1].item())
for a in range(1, n, CH):
    b = min(n, a + CH)
    ok2 = ok2 and bool((out[a:b] - out[a - 1:b - 1] == x[a - 1:b - 1]).all().item())
print("exclusive:", ok2)
print("BIG SCAN OK" if ok and ok2 else "BIG SCAN FAILED")
