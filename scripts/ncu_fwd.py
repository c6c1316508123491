"""one launch of the tcgen05 forward at the 64^2 level for `ncu --set full` (argv: np [G N d]; np = pairs of 8 on the polynomial, -1 = round-1 arithmetic)"""
import sys
import torch
sys.path.insert(0, ".")
from geodiffuser_b200 import _lib
from geodiffuser_b200._lib import call, stream
np_ = int(sys.argv[1])
G, N, d = (int(a) for a in sys.argv[2:5]) if len(sys.argv) > 4 else (3, 4096, 40)
H = 8
call("gd_attn_sm100_config", 0, np_)
g = torch.Generator(device="cuda").manual_seed(1)
mk = lambda: (torch.randn(H, N, d, device="cuda", generator=g) * 1.5).bfloat16()
qs = [mk() for _ in range(G)]
k, v = mk(), mk()
O = torch.empty(G, H, N, d, device="cuda"); L = torch.empty(G, H, N, device="cuda")
for _ in range(3):
    call("gd_attn_fwd_sm100", _lib.ptr_array(qs), _lib.ptr_array([k] * G), _lib.ptr_array([v] * G), _lib.ptr_array([O[i] for i in range(G)]),
         _lib.ptr_array([L[i] for i in range(G)]), None, G, H, N, N, d, d ** -0.5, None, 0, stream())
torch.cuda.synchronize()
