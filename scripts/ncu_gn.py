"""a few launches of the GroupNorm forward (argv: scheme B C S) for `ncu --set full`"""
import sys
import torch
sys.path.insert(0, ".")
from geodiffuser_b200 import _lib
from geodiffuser_b200._lib import call, ptr, stream
from geodiffuser_b200.body_ops import ptr_cl
scheme, B, C, S = (int(a) for a in sys.argv[1:5])
x = torch.randn(B, C, S, S, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last)
y = torch.empty_like(x)
w = torch.ones(C, device="cuda").bfloat16(); b = torch.zeros(C, device="cuda").bfloat16()
stats = torch.empty(B, 32, 2, device="cuda")
n = _lib.lib().gd_group_norm_nhwc_workspace(B, S * S, C, 32)
ws = torch.empty(n, device="cuda"); cnt = torch.zeros(64, device="cuda", dtype=torch.int32)
call("gd_group_norm_config", scheme)
for _ in range(4):
    call("gd_group_norm_nhwc_fwd", ptr_cl(x), None, ptr(w), ptr(b), 1, B, S * S, C, 32, 1e-5, 1, ptr(ws), n, ptr(cnt), ptr(stats), ptr_cl(y), stream())
torch.cuda.synchronize()
