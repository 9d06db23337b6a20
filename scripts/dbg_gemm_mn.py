"""Diagnostic: which smem element does the MN-major UMMA operand read?  (GPU only)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from plankassembly_b200 import ops
torch.set_printoptions(linewidth=250, edgeitems=8, precision=1, sci_mode=False)
M, N, K = 128, 128, 32
a = torch.zeros(M, K)
for m in range(M):
    a[m, m % K] = 1.0                       # row m selects contraction index k = m % 32
for name, fn in (('n', lambda k, n: float(n)), ('k', lambda k, n: float(k))):
    b = torch.tensor([[fn(k, n) for n in range(N)] for k in range(K)])      # stored [K, N]  (MN-major B)
    c = torch.full((M, N), float('nan'), device='cuda')
    ops.gemm_tf32(a.cuda(), b.cuda(), c, M, N, K, lda=K, ldb=N, ldc=N, b_mn=True)
    torch.cuda.synchronize()
    c = c.cpu()
    print(f'--- B[k][n] = {name}; expected C[m][n] = {name} with k = m % 32')
    print('rows 0..3, cols 0..40:\n', c[:4, :40])
    print('col 0, rows 0..63:\n', c[:64, 0])
    print('col 33, rows 0..40:\n', c[:40, 33])
    exp = torch.tensor([[fn(m % K, n) for n in range(N)] for m in range(M)])
    print('exact match:', torch.equal(c, exp), ' nonzero frac', (c != 0).float().mean().item(), 'nan', torch.isnan(c).any().item())
# A MN-major: A stored [K, M]; B K-major identity-selector
b = torch.zeros(N, K)
for n in range(N):
    b[n, n % K] = 1.0
for name, fn in (('m', lambda k, m: float(m)), ('k', lambda k, m: float(k))):
    a2 = torch.tensor([[fn(k, m) for m in range(M)] for k in range(K)])     # stored [K, M]
    c = torch.full((M, N), float('nan'), device='cuda')
    # need (A_MN, B_K) -> not built; use (MN, MN): B stored [K, N] selector
    b2 = b.T.contiguous()
    ops.gemm_tf32(a2.cuda(), b2.cuda(), c, M, N, K, lda=M, ldb=N, ldc=N, a_mn=True, b_mn=True)
    torch.cuda.synchronize()
    c = c.cpu()
    exp = torch.tensor([[fn(n % K, m) for n in range(N)] for m in range(M)])
    print(f'--- A[k][m] = {name} (MN,MN): exact match:', torch.equal(c, exp), ' nonzero frac', (c != 0).float().mean().item())
    print(c[:4, :40])
