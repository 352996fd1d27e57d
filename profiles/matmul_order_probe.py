"""Probe: fp32 summation order of torch.matmul for the three broadcast 3x3 @ 3x1 products of get_lidar_coor."""
import torch
torch.manual_seed(0)
dev = 'cuda:0'
B, N, D, H, W = 2, 6, 88, 16, 44
def f32(x): return x.to(torch.float32)
def fma(a, b, c): return f32(a.double() * b.double() + c.double())
def mul(a, b): return a * b
def variants(A, p):
    a0, a1, a2 = A[..., 0], A[..., 1], A[..., 2]        # rows broadcast: A[..., i, k]
    x0, x1, x2 = p[..., None, 0], p[..., None, 1], p[..., None, 2]
    t0, t1, t2 = a0 * x0, a1 * x1, a2 * x2
    return {
        '(t0+t1)+t2': (t0 + t1) + t2,
        'fma(a2,x2,fma(a1,x1,t0))': fma(a2, x2, fma(a1, x1, t0)),
        'fma(a2,x2,t0+t1)': fma(a2, x2, t0 + t1),
        'fma(a1,x1,t0)+t2': fma(a1, x1, t0) + t2,
        't0+(t1+t2)': t0 + (t1 + t2),
        'fma(a0,x0,fma(a1,x1,t2))': fma(a0, x0, fma(a1, x1, t2)),
        'fma(a0,x0,t1+t2)': fma(a0, x0, t1 + t2),
        't0+fma(a1,x1,t2)': t0 + fma(a1, x1, t2),
        '(t0+t2)+t1': (t0 + t2) + t1,
        'fma(a1,x1,fma(a2,x2,t0))': fma(a1, x1, fma(a2, x2, t0)),
        'fma(a2,x2,fma(a0,x0,t1))': fma(a2, x2, fma(a0, x0, t1)),
    }
p = torch.randn(B, N, D, H, W, 3, device=dev)
cases = {
    'stage1/2: A.view(B,N,1,1,1,3,3)': torch.randn(B, N, 1, 1, 1, 3, 3, device=dev),
    'stage3:   A.view(B,1,1,1,1,3,3)': torch.randn(B, 1, 1, 1, 1, 3, 3, device=dev),
}
for name, A in cases.items():
    T = A.matmul(p.unsqueeze(-1)).squeeze(-1)
    print(name, 'result', tuple(T.shape))
    for vn, v in variants(A, p).items():
        nd = int((v.view(torch.int32) != T.view(torch.int32)).sum())
        print(f'    {vn:32s} mismatching floats: {nd} of {T.numel()}')
# the cat() input of stage 2 (non-contiguous slices) and the in-place add
q = torch.cat((p[..., :2, None] * p[..., 2:3, None], p[..., 2:3, None]), 5)
A = cases['stage1/2: A.view(B,N,1,1,1,3,3)']
T = A.matmul(q).squeeze(-1)
for vn, v in variants(A, q.squeeze(-1)).items():
    nd = int((v.view(torch.int32) != T.view(torch.int32)).sum())
    if nd < T.numel() // 100: print(f'  stage2 input via cat: {vn:32s} mismatching floats: {nd}')
print(torch.__version__, torch.backends.cuda.matmul.allow_tf32)
