"""n >= 2^30 takes the 64-bit look-back words (C = unsigned long long): sort 2^30 + 12345 u32 keys (keys only, then pairs) and
check sortedness, the key checksum and that the values are the matching permutation."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import raja_b200 as rb
n = (1 << 30) + 12345
g = torch.Generator(device="cuda").manual_seed(5)
k = torch.randint(-2**31, 2**31, (n,), dtype=torch.int64, device="cuda", generator=g).to(torch.int32)
ref_sum = int(k.to(torch.int64).sum().item())
ku = k.view(torch.uint32)
rb.sort(ku)
torch.cuda.synchronize()
s = (k ^ -2**31)  # unsigned order as signed order
ok_sorted = bool((s[1:] >= s[:-1]).all().item())
print("keys only: sorted", ok_sorted, "checksum", int(k.to(torch.int64).sum().item()) == ref_sum)
del s
k2 = torch.randint(-2**31, 2**31, (n,), dtype=torch.int64, device="cuda", generator=g).to(torch.int32)
orig = k2.clone()
v = torch.arange(n, dtype=torch.int32, device="cuda")
rb.sort_pairs(k2.view(torch.uint32), v.view(torch.uint32))
torch.cuda.synchronize()
s = (k2 ^ -2**31)
ok_sorted2 = bool((s[1:] >= s[:-1]).all().item())
perm_ok = bool((orig[v.to(torch.int64)] == k2).all().item())
# stability: equal keys keep ascending original positions
eq = s[1:] == s[:-1]
stable = bool((v[1:][eq] > v[:-1][eq]).all().item())
print("pairs: sorted", ok_sorted2, "values are the permutation", perm_ok, "stable", stable)
print("BIG SORT OK" if (ok_sorted and ok_sorted2 and perm_ok and stable) else "BIG SORT FAILED")
