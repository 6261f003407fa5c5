"""Small driver for ncu captures of the tensor-core kernels (GRU gate GEMM and the NT GEMM)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from www2023tiger_b200 import ops
torch.manual_seed(0)
dev = 'cuda'
rows, M, d, N = 1900, 688, 172, 11000
cell = torch.nn.GRUCell(M, d)
pack = ops.GruPack(cell.weight_ih.to(dev), cell.weight_hh.to(dev), cell.bias_ih.to(dev), cell.bias_hh.to(dev))
xt, ht = torch.randn(N, M, device=dev), torch.randn(N, d, device=dev)
ids = torch.randint(0, N, (rows,), device=dev)
out = torch.empty(rows, d, device=dev)
a, w, b = torch.randn(600, 344, device=dev), torch.randn(344, 344, device=dev), torch.randn(344, device=dev)
c = torch.empty(600, 344, device=dev)
for _ in range(6):
    ops.gru_update(pack, node_ids=ids, x_table=xt, h_table=ht, n_rows=rows, out=out)
    ops.sgemm_nt(a, w, b, c)
torch.cuda.synchronize()
print('done')
