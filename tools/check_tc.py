"""GPU check of the tensor-core kernels (run on the B200 box): accuracy against float64 and timing
against the FFMA baseline."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import ctypes, subprocess
from www2023tiger_b200 import ops

torch.manual_seed(0)
dev = 'cuda'

# The CUDA-core GEMM the tensor-core kernels are compared with lives outside the product library
# (tools/baseline/gemm_ffma.cu); it is compiled here, on the GPU box, into its own shared object.
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_so = os.path.join(ROOT, 'tools', 'baseline', 'libtiger_ffma_baseline.so')
if not os.path.exists(_so):
    subprocess.check_call(['nvcc', '-O3', '-std=c++17', '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-shared',
                           '-Xcompiler', '-fPIC', '-I', os.path.join(ROOT, 'include'),
                           '-I', os.path.join(ROOT, 'www2023tiger_b200', 'csrc'),
                           os.path.join(ROOT, 'tools', 'baseline', 'gemm_ffma.cu'), '-o', _so])
_ffma = ctypes.CDLL(_so)
_P, _L, _I = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int
_ffma.tiger_sgemm_ffma.argtypes = [_P, _L, _P, _L, _P, _P, _L, _L, _P, _L, _I, _I, _I, _P]
_ffma.tiger_sgemm_ffma.restype = _I


def sgemm_ffma(a, w, bias, out):
    rc = _ffma.tiger_sgemm_ffma(a.data_ptr(), a.stride(0), w.data_ptr(), w.stride(0), bias.data_ptr(), out.data_ptr(),
                                out.stride(0), a.shape[0], None, 1, w.shape[0], a.shape[1], 0,
                                torch.cuda.current_stream().cuda_stream)
    assert rc == 0, rc

def timeit(fn, n=20, reps=5):
    """us per call, n calls captured in one CUDA graph (no host launch overhead in the number)."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            for _ in range(n):
                fn()
        g.replay()
        st.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(reps):
            g.replay()
        e1.record(st)
        st.synchronize()
    return e0.elapsed_time(e1) / (n * reps) * 1e3

ok = True
for (m, n, k) in [(128, 32, 32), (128, 64, 8), (1, 7, 5), (33, 65, 17), (54, 172, 860), (600, 344, 344), (600, 172, 516),
                  (600, 517, 172), (300, 1720, 860), (2200, 130, 54), (12000, 1720, 860)]:
    a = torch.randn(m, k)
    w = torch.randn(n, k) / k ** 0.5
    b = torch.randn(n)
    ref = (a.double() @ w.double().t() + b.double())
    ad, wd, bd = a.to(dev), w.to(dev), b.to(dev)
    out = torch.full((m, n), float('nan'), device=dev)
    ops.sgemm_nt(ad, wd, bd, out)
    torch.cuda.synchronize()
    err = (out.cpu().double() - ref).abs().max().item() / ref.abs().max().item()
    out2 = torch.empty(m, n, device=dev)
    sgemm_ffma(ad, wd, bd, out2)
    err2 = (out2.cpu().double() - ref).abs().max().item() / ref.abs().max().item()
    pk = ops.WeightPack(wd, m_rows_hint=m)
    out3 = torch.full((m, n), float('nan'), device=dev)
    ops.sgemm_nt_packed(ad, pk, bd, out3)
    torch.cuda.synchronize()
    err3 = (out3.cpu().double() - ref).abs().max().item() / ref.abs().max().item()
    t_pk = timeit(lambda: ops.sgemm_nt_packed(ad, pk, bd, out3))
    t_tc = timeit(lambda: ops.sgemm_nt(ad, wd, bd, out))
    t_ff = timeit(lambda: sgemm_ffma(ad, wd, bd, out2))
    flag = 'OK ' if max(err, err3) < 2e-6 else 'BAD'
    ok &= max(err, err3) < 2e-6
    print(f'{flag} gemm {m}x{n}x{k}: tc err {err:.2e} ({t_tc:.1f} us)  packed(bn={pk.bn}) err {err3:.2e} ({t_pk:.1f} us, {2*m*n*k/t_pk*1e-6:.1f} TF/s)   ffma err {err2:.2e} ({t_ff:.1f} us)', flush=True)

# GRU
for (rows, M, d, N) in [(100, 64, 16, 300), (1900, 688, 172, 11000), (1426, 304, 100, 7000), (6600, 688, 172, 11000)]:
    cell = torch.nn.GRUCell(M, d)
    x_table = torch.randn(N, M)
    h_table = torch.randn(N, d)
    ids = torch.randint(0, N, (rows,))
    with torch.no_grad():
        ref = cell.double()(x_table[ids].double(), h_table[ids].double())
    cell = cell.float()
    pack = ops.GruPack(cell.weight_ih.to(dev), cell.weight_hh.to(dev), cell.bias_ih.to(dev), cell.bias_hh.to(dev))
    xt, ht, idd = x_table.to(dev), h_table.to(dev), ids.to(dev)
    out = torch.full((rows, d), float('nan'), device=dev)
    ops.gru_update(pack, node_ids=idd, x_table=xt, h_table=ht, n_rows=rows, out=out)
    torch.cuda.synchronize()
    err = (out.cpu().double() - ref).abs().max().item() / ref.abs().max().item()
    t = timeit(lambda: ops.gru_update(pack, node_ids=idd, x_table=xt, h_table=ht, n_rows=rows, out=out))
    flag = 'OK ' if err < 2e-6 else 'BAD'
    ok &= err < 2e-6
    print(f'{flag} gru rows={rows} M={M} d={d}: err {err:.2e}  {t:.1f} us  ({2*rows*(M+d)*3*d/t*1e-6:.1f} TF/s)', flush=True)
print('ALL OK' if ok else 'FAILURES')
sys.exit(0 if ok else 1)
