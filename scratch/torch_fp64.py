# Scratch: library fp64 numbers on B200 for context (cuBLAS DGEMM peak, batched bmm, batched cholesky/cholesky_inverse).
import torch, time, json
torch.backends.cuda.matmul.allow_tf32 = False
dev = "cuda"
def t(f, reps=5, warm=2):
    for _ in range(warm): f()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
out = {}
for n in (4096, 8192):
    a = torch.randn(n, n, dtype=torch.float64, device=dev); b = torch.randn(n, n, dtype=torch.float64, device=dev)
    ms = t(lambda: a @ b)
    out[f"dgemm_{n}_tflops"] = 2 * n**3 / ms * 1e-9
    print(f"DGEMM {n}: {ms:.2f} ms  {2*n**3/ms*1e-9:.2f} TFLOP/s", flush=True)
# sustained
n = 8192
t0 = time.time(); cnt = 0
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
while time.time() - t0 < 4.0:
    for _ in range(5): c = a @ b
    cnt += 5; torch.cuda.synchronize()
e1.record(); torch.cuda.synchronize()
out["dgemm_8192_sustained_tflops"] = cnt * 2 * n**3 / e0.elapsed_time(e1) * 1e-9
print("DGEMM 8192 sustained", out["dgemm_8192_sustained_tflops"], flush=True)
for (B, n) in ((2000, 600), (500, 600), (2000, 152)):
    a = torch.randn(B, n, n, dtype=torch.float64, device=dev); b = torch.randn(B, n, n, dtype=torch.float64, device=dev)
    ms = t(lambda: torch.bmm(a, b.transpose(1, 2)))
    print(f"bmm NT B={B} n={n}: {ms:.2f} ms {B*2*n**3/ms*1e-9:.2f} TFLOP/s", flush=True)
    out[f"bmm_{B}_{n}_tflops"] = B * 2 * n**3 / ms * 1e-9
    spd = torch.bmm(a, a.transpose(1, 2)) + n * torch.eye(n, dtype=torch.float64, device=dev)
    ms = t(lambda: torch.linalg.cholesky(spd), reps=3, warm=1)
    print(f"cholesky B={B} n={n}: {ms:.2f} ms {B*n**3/3/ms*1e-9:.2f} TFLOP/s", flush=True)
    out[f"chol_{B}_{n}_ms"] = ms
    L = torch.linalg.cholesky(spd)
    ms = t(lambda: torch.cholesky_inverse(L), reps=3, warm=1)
    print(f"cholesky_inverse B={B} n={n}: {ms:.2f} ms {B*2*n**3/3/ms*1e-9:.2f} TFLOP/s", flush=True)
    out[f"cholinv_{B}_{n}_ms"] = ms
n = 8192
a = torch.randn(n, n, dtype=torch.float64, device=dev); spd = a @ a.T + n * torch.eye(n, dtype=torch.float64, device=dev)
ms = t(lambda: torch.linalg.cholesky(spd), reps=3, warm=1)
print(f"cholesky n=8192: {ms:.2f} ms {n**3/3/ms*1e-9:.2f} TFLOP/s")
out["chol_8192_ms"] = ms
json.dump(out, open("gpurun_out/torch_fp64.json", "w"), indent=1)
