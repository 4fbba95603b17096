"""torchrun --nproc-per-node N tools/check_nvls_exchange.py : NvlsExchange (one kernel over symmetric memory) against
FactoredExchange (NCCL all-gather + all-reduce) on the gradients of a real render, then a timing of both."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    from gaussiangrasper_b200 import scenes, ops
    from gaussiangrasper_b200.distributed import FactoredExchange, NvlsExchange
    from gaussiangrasper_b200.render import ViewBatch, render_views
    n, W, H, D = int(os.environ.get("GG_N", "500000")), 640, 480, 16
    sc = scenes.random_scene(n, feature_dim=D, seed=1235)
    names = ("means", "log_scales", "quats", "opacity_logit", "sh_coeffs", "features")
    P = {k: sc[k].to(dev).requires_grad_(True) for k in names}
    cams = scenes.orbit_cameras(1, W, H, first=rank, total=max(world, 8))
    vb = ViewBatch.from_cameras(cams, dev)
    all_pos = torch.stack([scenes.orbit_cameras(1, W, H, first=r, total=max(world, 8))[0].position for r in range(world)]).float().to(dev)
    v_img = torch.randn((1, H, W, 24), generator=torch.Generator().manual_seed(100 + rank)).to(dev)
    results = {}
    for name, cls in (("nccl", FactoredExchange), ("nvls", NvlsExchange)):
        ex = cls(P, 1)
        if name == "nvls" and rank == 0:
            print("multicast mapping:", ex.multicast, flush=True)
        for p in P.values():
            p.grad = None
        holder = ex.holder()
        out = render_views(*(P[k] for k in names), vb, holder=holder)
        out["image"].backward(v_img)
        g = ex.exchange(P["means"], vb.positions, 4, 4, holder, all_pos)
        torch.cuda.synchronize()
        results[name] = {k: v.clone() for k, v in g.items()}
        # timing of the exchange alone (buffers keep their contents: the cost does not depend on them)
        dist.barrier(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20):
            ex.exchange(P["means"], vb.positions, 4, 4, holder, all_pos)
        b.record(); torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b) / 20], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(f"{name}: exchange {float(t):.3f} ms per step (max over ranks, includes the SH rebuild)", flush=True)
        del ex
    worst = 0.0
    for k in names:
        a, b = results["nccl"][k], results["nvls"][k]
        scale = float(a.abs().max()) + 1e-30
        err = float((a - b).abs().max()) / scale
        worst = max(worst, err)
        if rank == 0:
            print(f"{k}: max |nvls - nccl| / max|nccl| = {err:.2e}", flush=True)
    ok = torch.tensor([1.0 if worst < 1e-5 else 0.0], device=dev)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("NVLS_EXCHANGE_OK" if float(ok) == 1.0 else "NVLS_EXCHANGE_MISMATCH", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
