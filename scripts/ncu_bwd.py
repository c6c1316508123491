"""one launch of the tcgen05 backward at the 64^2 level for `ncu --set full` (argv: [H N d M])"""
import sys
import torch
sys.path.insert(0, ".")
from geodiffuser_b200 import _lib
from geodiffuser_b200._lib import call, stream, ptr
H, N, d, M = (int(a) for a in sys.argv[1:5]) if len(sys.argv) > 4 else (8, 4096, 40, 0)
g = torch.Generator(device="cuda").manual_seed(1)
mk = lambda: (torch.randn(H, N, d, device="cuda", generator=g) * 1.5).bfloat16()
q, k, v, do = mk(), mk(), mk(), mk()
O = torch.empty(1, H, N, d, device="cuda"); L = torch.empty(1, H, N, device="cuda")
call("gd_attn_fwd_sm100", _lib.ptr_array([q]), _lib.ptr_array([k]), _lib.ptr_array([v]), _lib.ptr_array([O[0]]), _lib.ptr_array([L[0]]), None, 1, H, N, N, d,
     d ** -0.5, None, 0, stream())
delta = (do.float() * O[0]).sum(-1).contiguous()
dq = torch.empty(H, N, d, device="cuda")
for _ in range(3):
    call("gd_attn_bwd_sm100", ptr(q), ptr(k), ptr(v), ptr(do), ptr(L[0]), ptr(delta), None, None, None, (N + 7) // 8 * 8, 0, ptr(dq), H, N, d, d ** -0.5, None, 0, 0, stream())
torch.cuda.synchronize()
