"""In-kernel timeline of the GRU update (build with TIGER_EXTRA_NVCC_FLAGS=-DTIGER_TRACE)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from www2023tiger_b200 import ops
rows, M, d = (int(x) for x in sys.argv[1:4]) if len(sys.argv) > 3 else (1900, 688, 172)
N = 11000
g = torch.Generator().manual_seed(0)
x = torch.randn(N, M, generator=g).cuda(); h = torch.randn(N, d, generator=g).cuda()
w_ih = (torch.randn(3 * d, M, generator=g) / M ** 0.5).cuda(); w_hh = (torch.randn(3 * d, d, generator=g) / d ** 0.5).cuda()
b_ih = torch.randn(3 * d, generator=g).cuda(); b_hh = torch.randn(3 * d, generator=g).cuda()
ids = torch.randperm(N, generator=g)[:rows].cuda()
pack = ops.GruPack(w_ih, w_hh, b_ih, b_hh)
out = torch.empty(rows, d, device='cuda')
for i in range(3):
    print('--- launch', i, flush=True)
    ops.gru_update(pack, node_ids=ids, x_table=x, h_table=h, n_rows=rows, out=out)
    torch.cuda.synchronize()
