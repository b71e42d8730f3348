"""Texture fwd/bwd time for constant-uv (background-like), smooth and mixed coordinate images: checks the uniform-group
cache and pending-sum paths of the fast texture kernels.  Run on the GPU box."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import nvdiffrast_b200.torch as dr
from nvdiffrast_b200 import _lib

N, R = 8, 2048
dev = "cuda"
tex = torch.rand(1, 1024, 1024, 3, device=dev, requires_grad=True)
ys, xs = torch.meshgrid(torch.arange(R, device=dev), torch.arange(R, device=dev), indexing="ij")
smooth = torch.stack([(xs + 0.5) / R * 0.9, (ys + 0.5) / R * 0.9], -1)[None].expand(N, -1, -1, -1).contiguous().float()
da_s = torch.tensor([0.9 / R, 0, 0, 0.9 / R], device=dev).expand(N, R, R, 4).contiguous()
cases = {}
cases["constant"] = (torch.zeros(N, R, R, 2, device=dev), torch.zeros(N, R, R, 4, device=dev))
cases["smooth"] = (smooth, da_s)
mask = ((xs - R / 2) ** 2 + (ys - R / 2) ** 2 < (0.39 * R) ** 2)[None, ..., None]
cases["disc47%"] = (torch.where(mask, smooth, torch.zeros_like(smooth)), torch.where(mask, da_s, torch.zeros_like(da_s)))
dy = torch.randn(N, R, R, 3, device=dev)
for name, (uv, da) in cases.items():
    uv = uv.clone().requires_grad_(True); da = da.clone().requires_grad_(True)
    for mode in ("linear-mipmap-linear", "linear"):
        tf = tb = 0.0
        for it in range(6):
            trace = []
            _lib.set_trace(trace)
            out = dr.texture(tex, uv, da if "mip" in mode else None, filter_mode=mode, max_mip_level=9)
            out.backward(dy)
            torch.cuda.synchronize()
            _lib.set_trace(None)
            if it:                                      # device time of the C-ABI calls only (events around each call)
                tf += sum(a.elapsed_time(b) for n, a, b in trace if n == "b200dr_texture_fwd") / 5
                tb += sum(a.elapsed_time(b) for n, a, b in trace if n == "b200dr_texture_grad") / 5
            tex.grad = None; uv.grad = None; da.grad = None
        px = N * R * R
        if hasattr(_lib.load(), "b200dr_debug_tex_counters"):
            import ctypes
            c = (ctypes.c_ulonglong * 8)(); _lib.load().b200dr_debug_tex_counters(c, 1)
            print("   counters/iter fwd hit/miss/fill %d %d %d  grad hit/miss/fill/uniform %d %d %d %d  groups %d" % tuple([v // 6 for v in c[:7]] + [N * R * R // 32]))
        print("%-10s %-22s fwd %.3f ms (%.1f Gpx/s)  bwd %.3f ms (%.1f Gpx/s)" % (name, mode, tf, px / tf / 1e6, tb, px / tb / 1e6))
