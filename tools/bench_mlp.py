"""Time the fused tcgen05 up-projection (csrc/mlp.cu) on one 640x480 feature map against torch's two Linear
layers (fp32 SIMT GEMM and TF32 tensor-core GEMM), with the error of each against fp64.
usage: python tools/bench_mlp.py [H W]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch


def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def main():
    H, W = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (480, 640)
    from gaussiangrasper_b200.losses import UpProjection, up_project
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    mlp = UpProjection(32).to(dev)
    img = torch.randn((H, W, 40), device=dev)
    x = img[..., 7:39]                       # the feature channels of a blended image, rows 40 floats apart
    out = torch.empty((H, W, 512), device=dev)
    t_ours = timed(lambda: up_project(x, mlp, out=out))
    xc = x.contiguous()
    torch.backends.cuda.matmul.allow_tf32 = False
    with torch.no_grad():
        t_fp32 = timed(lambda: mlp(xc))
        y32 = mlp(xc)
        torch.backends.cuda.matmul.allow_tf32 = True
        t_tf32 = timed(lambda: mlp(xc))
        ytf = mlp(xc)
        torch.backends.cuda.matmul.allow_tf32 = False
        ref = mlp.double()(xc.double())
        mlp.float()
    scale = float(ref.abs().max())
    err = lambda y: float((y.double() - ref).abs().max()) / scale
    P = H * W
    flops = 2.0 * P * (32 * 128 + 128 * 512)
    out_bytes = P * 512 * 4 + P * 32 * 4
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    rep = {"rows": P, "ms": {"gg_mlp_up (tcgen05, 3xTF32)": t_ours, "torch fp32 (allow_tf32=False)": t_fp32,
                             "torch TF32 (allow_tf32=True)": t_tf32},
           "max_rel_err_vs_fp64": {"gg_mlp_up": err(out), "torch fp32": err(y32), "torch TF32": err(ytf)},
           "gg_mlp_up": {"algorithmic_GFLOP": flops / 1e9, "TFLOP/s (x3 executed)": 3 * flops / t_ours / 1e9,
                         "output+input GB/s": out_bytes / t_ours / 1e6, "frac_of_hbm_peak": out_bytes / t_ours / 1e6 / hbm,
                         "hbm_peak_GBs": hbm}}
    print(json.dumps(rep))


if __name__ == "__main__":
    main()
