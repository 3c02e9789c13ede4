# cuBLAS DGEMM denominators (FP64 roofline): square and the layer shape (100x100 times 100xN)
import torch, time
torch.backends.cuda.matmul.allow_tf32 = False
def bench(m, k, n, reps=5):
    a = torch.randn(m, k, dtype=torch.float64, device="cuda"); b = torch.randn(k, n, dtype=torch.float64, device="cuda")
    c = a @ b; torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); c = a @ b; e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(f"dgemm {m}x{k}x{n}: {best:.3f} ms {2.0*m*k*n/best/1e9:.2f} TFLOP/s", flush=True)
bench(8192, 8192, 8192)
bench(4096, 4096, 4096)
bench(100, 100, 4_000_000)
bench(104, 104, 4_000_000)
bench(128, 128, 4_000_000)
