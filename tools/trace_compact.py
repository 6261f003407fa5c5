"""Cold vs warm instruction cache of the single-CTA compaction kernel (build with -DTIGER_TRACE): the kernel is
launched twice back to back on identical inputs; between the pairs a large unrelated kernel evicts the caches."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from www2023tiger_b200 import ops
N, cap = 10984, 6656
g = torch.Generator().manual_seed(0)
ids = torch.randint(1, N, (2300,), generator=g).cuda()
bitmap = torch.zeros(ops.bitmap_words(N), dtype=torch.int32, device='cuda')
has_msg = (torch.rand(N, generator=g) < 0.8).to(torch.uint8).cuda()
upto = torch.ones(N, dtype=torch.uint8, device='cuda')
involved = torch.zeros(cap, dtype=torch.int64, device='cuda'); outdated = torch.zeros_like(involved); rst = torch.zeros_like(involved)
gru_row = torch.zeros(N, dtype=torch.int32, device='cuda'); counts = torch.zeros(4, dtype=torch.int32, device='cuda')
a = torch.randn(4096, 4096, device='cuda')
for rep in range(3):
    (a @ a).sum().item()          # evict
    for k in range(2):
        ops.mark_nodes(ids, bitmap, N)
        hm = has_msg.clone()
        torch.cuda.synchronize()
        print(f'--- rep {rep} launch {k}', flush=True)
        ops.compact_involved(bitmap, N, involved, counts, has_msg=hm, uptodate=upto, outdated=outdated, gru_row=gru_row,
                             restart_nodes=rst)
        torch.cuda.synchronize()
