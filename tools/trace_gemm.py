import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from www2023tiger_b200 import ops
m, n, k = (int(x) for x in sys.argv[1:4])
a, w, b = torch.randn(m, k, device='cuda'), torch.randn(n, k, device='cuda'), torch.randn(n, device='cuda')
c = torch.empty(m, n, device='cuda')
pk = ops.WeightPack(w, m_rows_hint=m)
for i in range(3):
    print('--- launch', i, flush=True)
    ops.sgemm_nt_packed(a, pk, b, c)
    torch.cuda.synchronize()
