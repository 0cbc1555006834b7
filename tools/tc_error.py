"""GPU experiment: error of the tcgen05 3xTF32 building block vs fp64 for split / accumulator variants."""
import os, subprocess, sys, json
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from stc_gnn_b200 import _lib
    lib = _lib.load()
    torch.backends.cuda.matmul.allow_tf32 = False
    res = {}
    for dist in ("pos", "randn"):
        for K in (32, 128, 512, 2048):
            g = torch.Generator().manual_seed(K)
            M, N = 4096, 32
            A = (torch.rand(M, K, generator=g) if dist == "pos" else torch.randn(M, K, generator=g)).cuda()
            B = (torch.rand(K, N, generator=g) if dist == "pos" else torch.randn(K, N, generator=g)).cuda()
            ref = A.double() @ B.double()
            D = torch.empty(M, N, device="cuda")
            _lib.check(lib.stc_tf32x3_gemm(A.data_ptr(), B.data_ptr(), D.data_ptr(), M, N, K, torch.cuda.current_stream().cuda_stream), "gemm")
            torch.cuda.synchronize()
            scale = ref.abs().mean().item()
            e = (D.double() - ref) / scale
            f = ((A @ B).double() - ref) / scale
            res[f"{dist}_K{K}"] = dict(tc_mean=e.mean().item(), tc_rms=e.pow(2).mean().sqrt().item(), tc_max=e.abs().max().item(),
                                      fp32_mean=f.mean().item(), fp32_rms=f.pow(2).mean().sqrt().item(), fp32_max=f.abs().max().item())
    print(json.dumps(res))
    sys.exit(0)
for mode, nmain, small in [(1, 1, 0), (0, 1, 0), (0, 1, 1), (0, 4, 1), (0, 15, 1)]:
    env = dict(os.environ, STC_TC_TEST_MODE=str(mode), STC_TC_TEST_NMAIN=str(nmain), STC_TC_TEST_SMALL=str(small))
    out = subprocess.run([sys.executable, __file__, "child"], env=env, capture_output=True, text=True)
    try:
        r = json.loads(out.stdout.strip().split("\n")[-1])
    except Exception:
        print("FAILED", mode, nmain, small, out.stderr[-500:]); continue
    print(f"--- split={'trunc' if mode & 1 else 'rne'} nmain={nmain} small_separate={small}")
    for k, v in r.items():
        print(f"  {k:10s} tc: mean {v['tc_mean']:+.2e} rms {v['tc_rms']:.2e} max {v['tc_max']:.2e} | fp32 cublas: mean {v['fp32_mean']:+.2e} rms {v['fp32_rms']:.2e} max {v['fp32_max']:.2e}")
