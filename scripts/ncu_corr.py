"""one launch sequence of the tcgen05 removal-loss correlation at the 64^2 level for `ncu --set full` (argv: [H N d M])"""
import sys
import torch
sys.path.insert(0, ".")
from geodiffuser_b200._lib import call, stream, ptr
H, N, d, M = (int(a) for a in sys.argv[1:5]) if len(sys.argv) > 4 else (8, 4096, 40, 410)
g = torch.Generator(device="cuda").manual_seed(1)
mk = lambda: (torch.randn(H, N, d, device="cuda", generator=g) * 1.5).bfloat16()
q_b, k_b = mk(), mk()
scale = d ** -0.5
lse_b = torch.logsumexp(torch.einsum("hnd,hkd->hnk", q_b.float(), k_b.float()) * scale, -1).contiguous()
a_e = torch.softmax(torch.randn(H, M, N, device="cuda", generator=g), -1).bfloat16().contiguous()
m_in = torch.zeros(N, device="cuda"); m_in[N // 3:N // 3 + M] = 1.0
m_bg = (1 - m_in).contiguous()
part = torch.empty(H, N // 32, M, 4, device="cuda")
for _ in range(3):
    call("gd_removal_corr_sm100", ptr(q_b), ptr(k_b), ptr(lse_b), ptr(a_e), H, M, N, d, float(scale), N, None, ptr(m_in), ptr(m_bg), ptr(part), stream())
torch.cuda.synchronize()
