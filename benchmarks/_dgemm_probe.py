import torch, time
torch.backends.cuda.matmul.allow_tf32 = False
n, w, k = 1306127, 70, 52
for layout in ('row', 'col'):
    if layout == 'row':
        V = torch.randn(n, w, dtype=torch.float64, device='cuda')
    else:
        V = torch.randn(w, n, dtype=torch.float64, device='cuda').t()      # column-major n x w
    M = torch.randn(w, k, dtype=torch.float64, device='cuda')
    for _ in range(3): out = V @ M
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): out = V @ M
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(layout, 'ms', round(ms, 3), 'TFLOP/s', round(2 * n * w * k / ms / 1e9, 2), 'GB/s', round((n * w + n * k) * 8 / ms / 1e6, 1))
# square-ish DGEMM peak
A = torch.randn(8192, 8192, dtype=torch.float64, device='cuda'); B = torch.randn(8192, 8192, dtype=torch.float64, device='cuda')
for _ in range(2): C = A @ B
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); C = A @ B; C = A @ B; e1.record(); torch.cuda.synchronize()
print('dgemm 8192^3 TFLOP/s', round(2 * 2 * 8192 ** 3 / e0.elapsed_time(e1) / 1e9, 2))
