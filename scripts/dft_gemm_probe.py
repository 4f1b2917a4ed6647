"""Evidence for the FFT-vs-DFT-GEMM choice (BASELINE config 3): how fast could the transform alone be as a
tensor-core GEMM?  Times cuBLAS on frames(BT x L) @ DFT(L x 2F') with TF32 (single pass: NOT accurate enough for the
1e-5 log-mel budget; a 3xTF32 split costs 3 passes) and with FP32 FFMA, next to our whole fused forward
(framing + window + FFT + |.|^2 + mel + log).  A GEMM DFT would still need the framing, |.|^2, mel and log around it."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import dmel_b200 as dm

dev = torch.device("cuda:0")

def timeit(fn, n=30):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3   # us

for n_fft in (128, 256, 512, 1024, 2048):
    w = bench.workload(f"sweep_{n_fft}")
    B, N, hop, M, sr, lam = w["B"], w["N"], w["hop"], w["M"], w["sr"], w["lam"]
    T = N // hop + 1
    F2 = 2 * (n_fft // 2 + 1)
    frames = torch.randn(B * T, n_fft, device=dev)
    dft = torch.randn(n_fft, F2, device=dev)
    out = torch.empty(B * T, F2, device=dev)
    torch.backends.cuda.matmul.allow_tf32 = True
    t_tf32 = timeit(lambda: torch.matmul(frames, dft, out=out))
    torch.backends.cuda.matmul.allow_tf32 = False
    t_fp32 = timeit(lambda: torch.matmul(frames, dft, out=out))
    x = torch.randn(B, N, device=dev)
    ll = torch.tensor(lam, device=dev)
    kw = dict(hop_length=hop, optimized=True, n_mels=M, sample_rate=sr, log_eps=1e-10, lambd_value=lam)
    t_ours = timeit(lambda: dm.dmel(x, ll, **kw))
    flops = 2.0 * B * T * n_fft * F2
    print(json.dumps({"n_fft": n_fft, "frames": B * T, "gemm_tf32_1pass_us": round(t_tf32, 1), "gemm_3xtf32_us_est": round(3 * t_tf32, 1),
                      "gemm_fp32_us": round(t_fp32, 1), "gemm_tf32_tflops": round(flops / t_tf32 / 1e6, 1),
                      "our_fused_forward_us (prep+fwd, eager)": round(t_ours, 1)}), flush=True)
