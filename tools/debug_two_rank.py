"""Diagnostic: where do two view-sharded ranks (two processes on one GPU, gloo) first diverge?"""
import os, sys, socket
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch


def worker(rank, world, port, out_dir, reset_grad=True):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gaussiangrasper_b200 import scenes
    from gaussiangrasper_b200.render import ViewBatch, render_views
    from gaussiangrasper_b200.training import DensifyStats, FusedAdam
    dev = torch.device("cuda:0")
    n, W, H, D = 6000, 96, 64, 4
    sc = scenes.random_scene(n, feature_dim=D, seed=5)
    sc["log_scales"] = sc["log_scales"] + 0.8
    names = ("means", "log_scales", "quats", "opacity_logit", "sh_coeffs", "features")
    P = {k: sc[k].to(dev).requires_grad_(True) for k in names}
    opt = FusedAdam(P)
    stats = DensifyStats(n, dev)
    cams = scenes.orbit_cameras(2 * world, W, H, total=2 * world)
    dump = {}
    for it in range(2):
        if reset_grad:
            for p in P.values():
                p.grad = None
        v = it * world + rank
        holder = {"grad_out": opt.bucket.unpack()}
        out = render_views(*(P[k] for k in names), ViewBatch.from_cameras([cams[v]], dev), holder=holder)
        out["image"].backward(torch.randn((1, H, W, out["image"].shape[-1]), generator=torch.Generator().manual_seed(v)).to(dev) * 1e-3)
        torch.cuda.synchronize()
        dump[f"local{it}"] = opt.bucket.flat.cpu().clone()
        stats.update(holder["v_geo"], holder["radii"], H, W)
        opt.bucket.all_reduce()
        torch.cuda.synchronize()
        dump[f"reduced{it}"] = opt.bucket.flat.cpu().clone()
        with torch.no_grad():
            opt.step()
        torch.cuda.synchronize()
        for k in names:
            dump[f"P{it}.{k}"] = P[k].detach().cpu().clone()
        dump[f"m{it}"] = opt.exp_avg.cpu().clone()
        dump[f"v{it}"] = opt.exp_avg_sq.cpu().clone()
    from gaussiangrasper_b200.training import refine_gaussians
    rules = dict(max_dim=float(max(W, H)), densify_grad_thresh=2e-7, densify_size_thresh=0.03, split_screen_size=0.05,
                 cull_alpha_thresh=0.1, cull_scale_thresh=0.5, cull_screen_size=0.15, do_densify=1, split_by_screen=1,
                 do_cull=1, cull_by_scale=0, cull_by_screen=0)
    new_p, new_m, info = refine_gaussians({k: P[k].detach() for k in names}, opt.moments(), stats, rules, seed=77, step=2)
    torch.cuda.synchronize()
    dump["stats.gn"] = stats.xys_grad_norm.cpu().clone(); dump["stats.vc"] = stats.vis_counts.cpu().clone()
    dump["stats.m2d"] = stats.max_2Dsize.cpu().clone()
    for k in names:
        dump[f"Pin.{k}"] = P[k].detach().cpu().clone()
        dump[f"new.{k}"] = new_p[k].cpu().clone()
        dump[f"newm.{k}"] = new_m[k][0].cpu().clone()
    print(rank, info, flush=True)
    dump["offsets"] = dict(opt.bucket.offsets)
    torch.save(dump, os.path.join(out_dir, f"dbg{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    import tempfile
    import torch.multiprocessing as mp
    d = tempfile.mkdtemp()
    with socket.socket() as s_:
        s_.bind(("127.0.0.1", 0))
        port = s_.getsockname()[1]
    reset = "--no-reset" not in sys.argv
    mp.spawn(worker, args=(2, port, d, reset), nprocs=2, join=True)
    a, b = (torch.load(os.path.join(d, f"dbg{r}.pt")) for r in range(2))
    print("offsets", a["offsets"])
    for k in a:
        if k == "offsets":
            continue
        same = torch.equal(a[k], b[k])
        if k.startswith("local"):
            print(k, "differs as expected" if not same else "IDENTICAL (unexpected)")
            continue
        if not same:
            diff = (a[k] != b[k]).reshape(-1)
            idx = torch.nonzero(diff).reshape(-1)
            print(k, "DIFFERS in", int(diff.sum()), "of", diff.numel(), "first idx", idx[:5].tolist(),
                  "vals", a[k].reshape(-1)[idx[:3]].tolist(), b[k].reshape(-1)[idx[:3]].tolist())
        else:
            print(k, "same")
