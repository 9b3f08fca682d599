"""cuBLAS DGEMM peak via torch (library call, measurement only): the FP64 roofline denominator."""
import json, time, torch
def main():
    dev = torch.device("cuda:0")
    out = {}
    for n in (4096, 8192):
        a = torch.randn(n, n, dtype=torch.float64, device=dev)
        b = torch.randn(n, n, dtype=torch.float64, device=dev)
        for _ in range(2):
            torch.matmul(a, b)
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(5):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); torch.matmul(a, b); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        out[f"dgemm_tflops_{n}"] = 2.0 * n ** 3 / best * 1e-9
        # sustained: back-to-back for ~3 s
        t0 = time.time(); cnt = 0
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        while time.time() - t0 < 3.0:
            for _ in range(4):
                torch.matmul(a, b); cnt += 1
            torch.cuda.synchronize()
        e1.record(); torch.cuda.synchronize()
        out[f"dgemm_tflops_sustained_{n}"] = 2.0 * n ** 3 * cnt / e0.elapsed_time(e1) * 1e-9
    # batched potrf via torch (cuSOLVER/MAGMA) as a library yardstick for the batched Cholesky
    for (bs, n) in ((64, 1600),):
        a = torch.randn(bs, n, n, dtype=torch.float64, device=dev)
        h = a @ a.transpose(1, 2) + n * torch.eye(n, dtype=torch.float64, device=dev)
        torch.linalg.cholesky(h); torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); torch.linalg.cholesky(h); e1.record(); torch.cuda.synchronize()
        out[f"lib_potrf_batched_{bs}x{n}_tflops"] = bs * n ** 3 / 3.0 / e0.elapsed_time(e1) * 1e-9
    print(json.dumps(out))
if __name__ == "__main__":
    main()
