#!/usr/bin/env python
"""bench.py -- Mpixels/s forward+backward of the differentiable-rasterizer hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c3|c2|c4|c5]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

Workload (BASELINE.json): c3 = "Spot-like" bumped sphere (5 856 triangles, 2 930 vertices), 2048x2048, 16 images per GPU,
full chain rasterize -> interpolate(uv, all pixel differentials) -> texture(trilinear, 1024^2 x3) -> antialias, forward +
backward, seeded synthetic data (workloads.py).  A step = one such pass over one batch; per-GPU work is fixed (weak scaling).
With N > 1 the default workload is c5 (the same chain, 32 images per GPU) and every step all-reduces the gradients of ALL
shared parameters (texture 12.6 MB + uv + vertices) through nvdiffrast_b200.distributed.SharedGradReducer: the texture
gradient leaves while interpolate / rasterize backward still run, the small rest in one call at the end.

One JSON line on rank 0.  `value` = whole-job Mpixel/s with inputs resident in HBM, averaged over `timed_rounds` rounds of
exactly K steps each (the rounds run back to back until the timed region is >= 1 s, so that the clock record holds tens
of samples); `e2e` = the same metric through the public API with the step's inputs copied from pinned host memory and the
loss + vertex gradient read back every step; `roofline` = algorithmic HBM bytes of the dominant op / its CUDA-event time,
against the measured copy peak (MEASURED_PEAKS.json); `other_configs` (N = 1 only) = c2, c4 and c5's per-GPU batch with the
same timing rules and their own clock records; `target_kernels` = interpolate / antialias forward + backward at A = C = 3,
2048^2 (the north-star's >= 70 % kernels); `cpu_baseline` = the CPU oracle (a port of the reference math; the reference has
no CPU path) timed on a bounded sample of the same workload.
"""
import argparse
import contextlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import workloads as wl  # noqa: E402
from nvdiffrast_b200.distributed import SharedGradReducer  # noqa: E402

FALLBACK_HBM_GBS = 6650.0   # /opt/skills/guides/B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent


def measured_traffic():
    """DRAM bytes per pixel per op from THIS round's ncu --set full capture (profiles/r02_traffic.json, written by
    scripts/ncu_summary.py from the .ncu-rep files of the same build), or {} -> `traffic: null`."""
    p = os.path.join(ROOT, "profiles", "r02_traffic.json")
    try:
        return json.load(open(p))
    except Exception:
        return {}


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------------------------
# clocks


class ClockSampler:
    FIELDS = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index
        self.samples = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits", "-lms", "25"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        if not self.samples:                             # the polling loop produced nothing in time: one direct query
            try:
                r = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits"],
                                   capture_output=True, text=True, timeout=20)
                self.samples = [l.strip() for l in r.stdout.splitlines() if l.strip()]
            except Exception:
                pass
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            f = [x.strip() for x in s.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(mx)) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------
# the pipeline (same code drives our ops and the reference's)


def op_bytes(workload, P, A, D, C):
    """Algorithmic HBM bytes per C-ABI op call for P pixels (every per-pixel input read once, every
    output written once; L2-resident mesh / texture gathers excluded)."""
    b = {
        "b200dr_rasterize_fwd": 32 * P,
        "b200dr_rasterize_grad": (32 + (16 if workload == "c3" else 0)) * P,
        "b200dr_interpolate_fwd": (16 + 4 * A + (16 + 8 * D if D else 0)) * P,
        "b200dr_interpolate_grad": (32 + 4 * A + (32 + 8 * D if D else 0)) * P,
        "b200dr_antialias_fwd": (8 * C + 16) * P,
        "b200dr_antialias_grad": 8 * C * P,
        "b200dr_texture_fwd": (8 + 4 * C + 16) * P,
        "b200dr_texture_grad": (16 + 4 * C + 32) * P,
        # fused texture -> interpolate -> rasterize backward: dy + rast + rast_db in, nothing per-pixel out (the three ops it
        # replaces move (16+4C+32) + (32+4A+32+8D) + 48 = 196 B/px through their tensor boundaries)
        "b200dr_texture_chain_grad": (4 * C + 32) * P,
        # fused interpolate -> rasterize backward: rast + dy (+ rast_db + dda) in, nothing per-pixel out
        "b200dr_interpolate_chain_grad": (16 + 4 * A + (16 + 8 * D if D else 0)) * P,
    }
    return b


def replaced_bytes(nbytes):
    """Bytes the op-by-op kernels move through their tensor boundaries for the work of each fused backward op."""
    ig, rg = nbytes["b200dr_interpolate_grad"], nbytes["b200dr_rasterize_grad"]
    return {"b200dr_texture_chain_grad": nbytes["b200dr_texture_grad"] + ig + rg, "b200dr_interpolate_chain_grad": ig + rg}


def op_by_op_bytes(nbytes, ops, K_total):
    """Algorithmic bytes per step counted at the reference's op boundaries (fused ops replaced by what they stand for)."""
    rep = replaced_bytes(nbytes)
    return sum((rep.get(n, nbytes.get(n, 0))) * ops[n][1] / K_total for n in ops if n in nbytes)


class Pipeline:
    """One workload on one GPU: pinned host inputs, two device input slots (the end-to-end loop double-buffers), the op chain
    through `mod` (our front end or the reference's, same code), and the shared-gradient exchange."""

    def __init__(self, mod, workload, n_img, res, device, seed, tex_res=1024):
        self.mod, self.workload, self.device = mod, workload, device
        self.full = workload in ("c3", "c5")                   # full chain (texture + antialias); c2 / c4: rasterize + interpolate
        if self.full:
            c = wl.config_c3(n=1, res=res, tex_res=tex_res)
            v = wl.sphere_mesh()[0]
        elif workload == "c2":
            c = wl.config_c2(n=1, res=res)
            v = wl.cube_mesh()[0]
        else:
            c = wl.config_c4(n=1, res=res)
            v = wl.sphere_mesh(1000, 501)[0]
        far = 50.0 if workload == "c2" else 200.0
        self.res = (res, res)
        self.n_img = n_img
        self.host = {                                   # pinned host copies of everything a step consumes
            "verts": torch.from_numpy(np.concatenate([v, np.ones((v.shape[0], 1), np.float32)], 1)).pin_memory(),
            "mvp": torch.from_numpy(wl.camera_batch(n_img, seed, far=far)).pin_memory(),
        }
        if self.full:
            self.host["uv"] = torch.from_numpy(c["uv"]).pin_memory()
            self.host["tex"] = torch.from_numpy(c["tex"]).pin_memory()
        else:
            self.host["attr"] = torch.from_numpy(c["attr"]).pin_memory()
        self.shared = [k for k in ("tex", "uv", "attr", "verts") if k in self.host]      # parameters shared by every image
        self.tri = torch.from_numpy(c["tri"]).to(device)
        self.uv_idx = torch.from_numpy(c["uv_idx"]).to(device) if self.full else None
        self.ctx = mod.RasterizeCudaContext(device)
        self.topo = mod.antialias_construct_topology_hash(self.tri) if self.full else None
        self.bufs = [{}, {}]
        self.upload(0); self.upload(1)
        self.use(0)
        C = 3
        g = torch.Generator(device=device); g.manual_seed(2 + seed)
        self.dy = torch.randn((n_img, res, res, C), device=device, generator=g)
        self.grad_host = [torch.empty((v.shape[0], 4), dtype=torch.float32).pin_memory() for _ in range(2)]
        self.loss_host = [torch.empty((1,), dtype=torch.float32).pin_memory() for _ in range(2)]
        self.reducer = SharedGradReducer([(k, self.host[k].shape) for k in self.shared], device)

    def h2d_bytes(self):
        return int(sum(t.numel() * t.element_size() for t in self.host.values()))

    def d2h_bytes(self):
        return int(self.grad_host[0].numel() * 4 + 4)

    def upload(self, slot=0, stream=None):
        """Copy every per-step input from pinned host memory into device buffer set `slot` (two sets: the end-to-end loop
        uploads step i+1 on a copy stream while step i computes)."""
        bufs = self.bufs[slot]
        with torch.cuda.stream(stream) if stream is not None else contextlib.nullcontext():
            for k, t in self.host.items():
                d = bufs.get(k)
                if d is None:
                    bufs[k] = t.to(self.device, non_blocking=True)
                else:
                    with torch.no_grad():
                        d.copy_(t, non_blocking=True)

    def use(self, slot):
        self.dev = dict(self.bufs[slot])

    def step(self, with_loss=False):
        """forward + backward + all-reduce of every shared-parameter gradient.  Returns (loss, dict of reduced gradients).
        with_loss: form the scalar loss sum(out * dy) explicitly (end-to-end runs read it back); otherwise dy is fed
        straight into backward, so the timed region contains the hot path and nothing else."""
        mod = self.mod
        leaves = {k: self.dev[k].detach().requires_grad_(True) for k in self.shared}
        self.reducer.begin(leaves)
        pos = torch.matmul(leaves["verts"][None], self.dev["mvp"].transpose(1, 2)).contiguous()      # [n, V, 4] clip space
        rast, db = mod.rasterize(self.ctx, pos, self.tri, self.res)
        if os.environ.get("BENCH_DEBUG"):
            sys.stderr.write("coverage %.4f\n" % float((rast[..., 3] > 0).float().mean()))
        if self.full:
            texc, texd = mod.interpolate(leaves["uv"][None], rast, self.uv_idx, rast_db=db, diff_attrs="all")
            color = mod.texture(leaves["tex"], texc, texd, filter_mode="linear-mipmap-linear", max_mip_level=9)
            out = mod.antialias(color, rast, pos, self.tri, topology_hash=self.topo)
        else:
            out, _ = mod.interpolate(leaves["attr"][None], rast, self.tri)
        # scalar loss <out, dy> (end-to-end runs read it back): one pass over the image; its gradient w.r.t. out is dy by
        # construction and is fed straight into backward (same for both arms)
        loss = torch.dot(out.detach().reshape(-1), self.dy.reshape(-1)) if with_loss else None
        out.backward(self.dy)
        return loss, self.reducer.finish()


def load_module(impl):
    if impl == "ours":
        import nvdiffrast_b200.torch as dr
        return dr, "ours"
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from util import load_reference
    ref = load_reference()
    return ref, ("reference-cuda" if ref is not None else None)


# ----------------------------------------------------------------------------------------------
# CPU oracle timing (cpu_baseline / reference arm fallback)


def cpu_oracle_rate(workload, res, samples=1):
    """Mpixel/s of the CPU oracle (port of the reference math, all host threads) on `samples` images of the workload."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle
    from util import oracle_chain
    oracle.build()
    if workload in ("c3", "c5"):
        c = wl.config_c3(n=samples, res=res)
    elif workload == "c2":
        c = wl.config_c2(n=samples, res=res)
    else:
        c = wl.config_c4(n=samples, res=res)
    dy = np.random.default_rng(2).normal(size=(samples, res, res, 3)).astype(np.float32)
    t0 = time.perf_counter()
    oracle_chain(c, samples, workload in ("c3", "c5"), dy, max_mip_level=9)
    dt = time.perf_counter() - t0
    return samples * res * res / dt / 1e6, dt, oracle.num_threads()


# ----------------------------------------------------------------------------------------------


_REAL_STDOUT = None


def claim_stdout():
    """stdout must carry exactly ONE JSON line.  Libraries (NCCL's version banner, torchrun notices) write to fd 1 behind
    Python's back, so fd 1 is pointed at stderr for the whole run and the result line goes to a private copy of the
    original stdout."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(obj):
    _REAL_STDOUT.write(json.dumps(obj) + "\n")
    _REAL_STDOUT.flush()


DEFAULTS = {"c3": (2048, 16), "c2": (1024, 64), "c4": (2048, 8), "c5": (2048, 32)}
WORKLOAD_TEXT = {
    "c3": "c3: Spot-like sphere 5856 tris, %d^2, %d images/GPU, rasterize+interpolate(uv,da)+texture(trilinear 1024^2x3)+antialias fwd+bwd",
    "c5": "c5: Spot-like sphere 5856 tris, %d^2, %d images/GPU (256 over 8 GPUs), rasterize+interpolate(uv,da)+texture(trilinear 1024^2x3)+antialias fwd+bwd",
    "c2": "c2: cube 12 tris, %d^2, %d images/GPU, rasterize+interpolate fwd+bwd",
    "c4": "c4: 1M-triangle sphere, %d^2, %d images/GPU, rasterize+interpolate fwd+bwd"}
MIN_TIMED_SECONDS = 1.0


def sync_all(world):
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()


def max_over_ranks(ms, device, world):
    t = torch.tensor([ms], device=device)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    return float(t.item())


def ops_table(trace, K_total, nbytes, peak):
    """Per-op CUDA-event times over the timed region -> {op: calls/step, ms, bytes, GB/s, frac}."""
    ops = {}
    for name, a, b in trace:
        d = ops.setdefault(name, [0.0, 0])
        d[0] += a.elapsed_time(b); d[1] += 1
    table = {}
    for name, (tot, cnt) in ops.items():
        avg = tot / cnt
        e = {"calls_per_step": cnt / K_total, "ms": avg}
        if name in nbytes:
            gbs = nbytes[name] / (avg * 1e-3) / 1e9
            e.update({"algorithmic_bytes": nbytes[name], "GBps": gbs, "frac": gbs / peak})
        table[name.replace("b200dr_", "")] = e
    return table, ops


def timed_run(pipe, K, W, world, rank, local, device, impl, min_seconds=MIN_TIMED_SECONDS):
    """W warm-up steps, then rounds of EXACTLY K steps, each round bracketed by CUDA events on the compute stream (barrier +
    synchronize on both sides of the whole region), repeated back to back until the region has lasted >= min_seconds so that
    the 25 ms clock sampler sees tens of samples.  Returns mean ms/step (max over ranks), rounds, clocks, op trace, launches."""
    for _ in range(W):
        pipe.step()
    trace, launches0 = [], 0
    if impl == "ours":
        from nvdiffrast_b200 import _lib
        _lib.set_trace(trace)
        launches0 = _lib.load().b200dr_launch_count()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    sync_all(world)
    rounds = []
    t_start = time.perf_counter()
    while True:
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(K):
            pipe.step()
        e1.record()
        e1.synchronize()
        rounds.append(e0.elapsed_time(e1))
        go = torch.tensor([1.0 if (time.perf_counter() - t_start) < min_seconds and len(rounds) < 200 else 0.0], device=device)
        if world > 1:                                           # every rank runs the same number of rounds: rank 0 decides
            torch.distributed.broadcast(go, 0)
        if float(go.item()) == 0.0:
            break
    sync_all(world)
    clocks = sampler.stop() if rank == 0 else None
    launches = 0
    if impl == "ours":
        _lib.set_trace(None)
        launches = int(_lib.load().b200dr_launch_count() - launches0)
    ms_step = max_over_ranks(sum(rounds) / (K * len(rounds)), device, world)
    return {"ms_per_step": ms_step, "rounds": len(rounds), "round_ms": [r / K for r in rounds[:16]], "clocks": clocks,
            "trace": trace, "launches": launches}


def e2e_run(pipe, K, world, device, min_seconds=MIN_TIMED_SECONDS):
    """Double-buffered end-to-end loop: step i+1's inputs travel host->device on a copy stream while step i computes; the
    loss and the vertex gradient of step i are read back asynchronously and consumed (event wait) while step i+1 is in
    flight.  All uploads and read-backs happen inside the timed region; rounds of K steps until >= 1 s."""
    copy_stream = torch.cuda.Stream(device=device)
    main = torch.cuda.current_stream()
    ready = [torch.cuda.Event(), torch.cuda.Event()]       # slot uploaded
    free = [torch.cuda.Event(), torch.cuda.Event()]        # slot no longer read by the compute stream
    done = [torch.cuda.Event(), torch.cuda.Event()]        # results of the step using this slot are on the host

    def run(steps):
        for sl in (0, 1):
            free[sl].record(main)
        copy_stream.wait_event(free[0])
        pipe.upload(0, copy_stream); ready[0].record(copy_stream)
        results = []
        for i in range(steps):
            sl = i & 1
            if i + 1 < steps:
                copy_stream.wait_event(free[sl ^ 1])
                pipe.upload(sl ^ 1, copy_stream); ready[sl ^ 1].record(copy_stream)
            main.wait_event(ready[sl])
            pipe.use(sl)
            loss, g = pipe.step(with_loss=True)
            free[sl].record(main)
            pipe.loss_host[sl].copy_(loss.detach().reshape(1), non_blocking=True)
            pipe.grad_host[sl].copy_(g["verts"], non_blocking=True)
            done[sl].record(main)
            if i > 0:
                done[sl ^ 1].synchronize()                 # the caller consumes the previous step's result
                results.append(float(pipe.loss_host[sl ^ 1][0]))
        done[(steps - 1) & 1].synchronize()
        results.append(float(pipe.loss_host[(steps - 1) & 1][0]))
        return results

    run(2)
    sync_all(world)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    total_ms, rounds, t_start = 0.0, 0, time.perf_counter()
    while True:
        e0.record()
        losses = run(K)
        e1.record(); e1.synchronize()
        assert len(losses) == K and all(np.isfinite(l) for l in losses)
        total_ms += e0.elapsed_time(e1); rounds += 1
        go = torch.tensor([1.0 if (time.perf_counter() - t_start) < min_seconds and rounds < 200 else 0.0], device=device)
        if world > 1:
            torch.distributed.broadcast(go, 0)
        if float(go.item()) == 0.0:
            break
    sync_all(world)
    return max_over_ranks(total_ms / (K * rounds), device, world), rounds


def side_config(mod, name, K, W, device, peak):
    """One more BASELINE configuration on this GPU (N = 1 only), same timing rules, own clock record."""
    res, n_img = DEFAULTS[name]
    pipe = Pipeline(mod, name, n_img, res, device, seed=0)
    r = timed_run(pipe, K, W, 1, 0, device.index or 0, device, "ours")
    P = n_img * res * res
    A, D, C = (2, 2, 3) if pipe.full else (3, 0, 3)
    nbytes = op_bytes("c3" if pipe.full else name, P, A, D, C)
    table, ops = ops_table(r["trace"], K * r["rounds"], nbytes, peak)
    chain = sum(nbytes[n] * ops[n][1] / (K * r["rounds"]) for n in ops if n in nbytes)
    api = op_by_op_bytes(nbytes, ops, K * r["rounds"])
    out = {"workload": WORKLOAD_TEXT[name] % (res, n_img), "ms_per_step": r["ms_per_step"], "value": P / (r["ms_per_step"] * 1e-3) / 1e6,
           "unit": "Mpixel/s", "steps": K, "timed_rounds": r["rounds"], "clocks": r["clocks"],
           "ops": {k: {"ms": v["ms"], "frac": v.get("frac")} for k, v in table.items()},
           "chain_roofline": {"algorithmic_bytes_per_step": chain, "frac": chain / (r["ms_per_step"] * 1e-3) / 1e9 / peak,
                              "op_by_op_bytes_per_step": api, "effective_frac_vs_op_by_op_bytes": api / (r["ms_per_step"] * 1e-3) / 1e9 / peak}}
    if any(n in ops for n in replaced_bytes(nbytes)):
        # the same chain with the fused backward nodes switched off (B200DR_FUSION=0), same inputs, same box
        mod.set_fusion(False)
        try:
            ru = timed_run(pipe, K, 1, 1, 0, device.index or 0, device, "ours", 0.5)
        finally:
            mod.set_fusion(True)
        tu, _ = ops_table(ru["trace"], K * ru["rounds"], nbytes, peak)
        out["ops_unfused"] = {"ms_per_step": ru["ms_per_step"], "ops": {k: {"ms": v["ms"], "frac": v.get("frac")} for k, v in tu.items()}}
    del pipe
    torch.cuda.empty_cache()
    return out


def target_kernels(mod, device, peak, n_img=16, res=2048):
    """The north-star's >= 70 % kernels in isolation: interpolate fwd/bwd (A = 3, no pixel differentials) and antialias
    fwd/bwd (C = 3) on the c3 geometry at 2048^2, timed per C-ABI op with CUDA events, >= 0.5 s per pair."""
    from nvdiffrast_b200 import _lib
    c = wl.config_c3(n=1, res=res, tex_res=8)
    v = wl.sphere_mesh()[0]
    verts = torch.from_numpy(np.concatenate([v, np.ones((v.shape[0], 1), np.float32)], 1)).to(device)
    mvp = torch.from_numpy(wl.camera_batch(n_img, 0, far=200.0)).to(device)
    pos = torch.matmul(verts[None], mvp.transpose(1, 2)).contiguous()
    tri = torch.from_numpy(c["tri"]).to(device)
    ctx = mod.RasterizeCudaContext(device)
    with torch.no_grad():
        rast, _ = mod.rasterize(ctx, pos, tri, (res, res))
    g = torch.Generator(device=device); g.manual_seed(7)
    attr = torch.rand((1, v.shape[0], 3), device=device, generator=g)
    dy = torch.randn((n_img, res, res, 3), device=device, generator=g)
    topo = mod.antialias_construct_topology_hash(tri)
    P = n_img * res * res
    nbytes = {"b200dr_interpolate_fwd": 28 * P, "b200dr_interpolate_grad": 44 * P, "b200dr_antialias_fwd": 40 * P, "b200dr_antialias_grad": 24 * P}

    def interp():
        a = attr.detach().requires_grad_(True); r = rast.detach().requires_grad_(True)
        o, _ = mod.interpolate(a, r, tri)
        o.backward(dy)

    def aa():
        col = dy.detach().requires_grad_(True); pp = pos.detach().requires_grad_(True)
        o = mod.antialias(col, rast, pp, tri, topology_hash=topo)
        o.backward(dy)

    out = {}
    sampler = ClockSampler(device.index or 0); sampler.start()
    for fn in (interp, aa):
        for _ in range(3):
            fn()
        trace = []
        _lib.set_trace(trace)
        torch.cuda.synchronize()
        t0 = time.perf_counter(); it = 0
        while time.perf_counter() - t0 < 0.5 or it < 10:
            fn(); it += 1
            if it % 10 == 0:
                torch.cuda.synchronize()
        torch.cuda.synchronize()
        _lib.set_trace(None)
        table, _ = ops_table(trace, it, nbytes, peak)
        for k, e in table.items():
            if "frac" in e:
                out[k] = {"ms": e["ms"], "bytes_per_pixel": nbytes["b200dr_" + k] // P, "GBps": e["GBps"], "frac": e["frac"], "calls": it}
    out["clocks"] = sampler.stop()
    out["config"] = "c3 geometry, %d^2, %d images, A = C = 3, coverage %.3f" % (res, n_img, float((rast[..., 3] > 0).float().mean()))
    return out


def envphong_shaped(mod, device, peak, n_img=4, res=1024, face=512, C=3):
    """Cube-map texturing at the shape of the reference's envphong sample (6 x 512^2 x 3 environment map sampled trilinearly
    by reflection-like directions over a 1024^2 image): a unit sphere's normals inside the silhouette, one
    constant direction outside; texture forward + backward timed per C-ABI op."""
    from nvdiffrast_b200 import _lib
    g = torch.Generator(device=device); g.manual_seed(11)
    tex = torch.rand((1, 6, face, face, C), device=device, generator=g)
    ys, xs = torch.meshgrid(torch.linspace(-1.2, 1.2, res, device=device), torch.linspace(-1.2, 1.2, res, device=device), indexing="ij")
    r2 = xs * xs + ys * ys
    inside = r2 < 1.0
    z = torch.sqrt(torch.clamp(1.0 - r2, min=1e-4))
    d = torch.stack([torch.where(inside, xs, torch.zeros_like(xs)), torch.where(inside, ys, torch.zeros_like(ys)), torch.where(inside, z, torch.ones_like(z))], -1)
    px = 2.4 / res
    dX = torch.stack([torch.full_like(xs, px), torch.zeros_like(xs), -xs / z * px], -1) * inside[..., None]
    dY = torch.stack([torch.zeros_like(xs), torch.full_like(xs, px), -ys / z * px], -1) * inside[..., None]
    uv = d[None].repeat(n_img, 1, 1, 1).contiguous()
    uv_da = torch.stack([dX[..., 0], dY[..., 0], dX[..., 1], dY[..., 1], dX[..., 2], dY[..., 2]], -1)[None].repeat(n_img, 1, 1, 1).contiguous()
    dy = torch.randn((n_img, res, res, C), device=device, generator=g)
    mip = mod.texture_construct_mip(tex, cube_mode=True)
    P = n_img * res * res
    nbytes = {"b200dr_texture_fwd": (12 + 4 * C + 24) * P, "b200dr_texture_grad": (24 + 4 * C + 48) * P}

    def step():
        t = tex.detach().requires_grad_(True); u = uv.detach().requires_grad_(True); da = uv_da.detach().requires_grad_(True)
        o = mod.texture(t, u, da, mip=mip, filter_mode="linear-mipmap-linear", boundary_mode="cube")
        o.backward(dy)

    for _ in range(3):
        step()
    trace = []
    _lib.set_trace(trace)
    torch.cuda.synchronize()
    t0 = time.perf_counter(); it = 0
    while time.perf_counter() - t0 < 0.4 or it < 10:
        step(); it += 1
        if it % 10 == 0:
            torch.cuda.synchronize()
    torch.cuda.synchronize()
    _lib.set_trace(None)
    table, _ = ops_table(trace, it, nbytes, peak)
    return {"workload": "envphong-shaped: cube map 6x%d^2x%d, %d^2, %d images, trilinear, fwd+bwd of texture()" % (face, C, res, n_img),
            "coverage": float(inside.float().mean()),
            "ops": {k: {"ms": v["ms"], "bytes_per_pixel": nbytes["b200dr_" + k] // P, "frac": v.get("frac")} for k, v in table.items() if "frac" in v}}


def graph_leg(res=256, n_img=1, tex_res=256):
    """Launch-bound regime (a fitting loop at low resolution): the same full chain, eager vs replayed from ONE CUDA graph
    (nvdiffrast_b200/graphs.py).  Timed with CUDA events around 100 eager steps / 300 replays after warm-up; inputs stay on
    the device (the regime of interest is the per-step launch cost, not the copies)."""
    import nvdiffrast_b200.torch as dr
    from nvdiffrast_b200 import graphs
    device = torch.device("cuda", 0)
    c = wl.config_c3(n=n_img, res=res, tex_res=tex_res)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device)
    ctx = dr.RasterizeCudaContext(device)
    tri, uv_idx = t(c["tri"]), t(c["uv_idx"])
    topo = dr.antialias_construct_topology_hash(tri)
    pos = t(c["pos"]).requires_grad_(True); uv = t(c["uv"]).requires_grad_(True); tex = t(c["tex"]).requires_grad_(True)
    dy = torch.randn((n_img, res, res, 3), device=device, generator=torch.Generator(device=device).manual_seed(2))

    def step():
        rast, db = dr.rasterize(ctx, pos, tri, c["res"])
        texc, texd = dr.interpolate(uv[None], rast, uv_idx, rast_db=db, diff_attrs="all")
        color = dr.texture(tex, texc, texd, filter_mode="linear-mipmap-linear", max_mip_level=7)
        out = dr.antialias(color, rast, pos, tri, topology_hash=topo)
        return torch.autograd.grad(out, [pos, uv, tex], dy)

    def timed(fn, iters):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record(); e1.synchronize()
        return e0.elapsed_time(e1) / iters

    eager_ms = timed(step, 100)
    captured = graphs.capture(step)
    graph_ms = timed(captured.replay, 300)
    g_eager = step(); g_graph = captured.replay()
    torch.cuda.synchronize()
    err = max(float((a - b).abs().max()) / max(float(a.abs().max()), 1e-30) for a, b in zip(g_eager, g_graph))
    P = n_img * res * res
    return {"workload": "full chain fwd+bwd, 5856-triangle mesh, %d^2, %d image(s), texture %d^2: launch-bound" % (res, n_img, tex_res),
            "eager_ms_per_step": eager_ms, "graph_ms_per_step": graph_ms, "speedup": eager_ms / graph_ms,
            "graph_Mpixel_per_s": P / (graph_ms * 1e-3) / 1e6, "max_rel_grad_diff_graph_vs_eager": err}


def graph_leg_subprocess(timeout=180):
    """Own process: a failed capture can leave the CUDA context unusable, and must not take the bench line with it."""
    import subprocess
    try:
        pr = subprocess.run([sys.executable, os.path.abspath(__file__), "--graph-leg"], stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                            text=True, timeout=timeout, cwd=ROOT)
        lines = [l for l in pr.stdout.splitlines() if l.startswith("{")]
        if pr.returncode != 0 or not lines:
            return {"error": (pr.stderr or "no output").strip().splitlines()[-1][:300]}
        return json.loads(lines[-1])
    except Exception as e:      # timeout, JSON error: report, never fail the bench line
        return {"error": "%s: %s" % (type(e).__name__, str(e)[:300])}


def main():
    claim_stdout()
    if "--graph-leg" in sys.argv:
        emit(graph_leg())
        return
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, choices=["c3", "c2", "c4", "c5"])
    ap.add_argument("--res", type=int, default=0)
    ap.add_argument("--batch", type=int, default=0, help="images per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip other_configs / target_kernels")
    ap.add_argument("--min-seconds", type=float, default=MIN_TIMED_SECONDS, help="rounds of K steps are repeated until the timed region lasted this long")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    workload = args.workload or ("c3" if world == 1 else "c5")
    res = args.res or DEFAULTS[workload][0]
    n_img = args.batch or DEFAULTS[workload][1]
    K, W = args.steps, max(args.warmup, 3)

    mod, kind = load_module(args.impl)
    if args.impl == "reference" and mod is None:
        # reference CUDA build not loadable here: time the CPU oracle on rank 0 instead (bounded sample)
        if rank == 0:
            rate, dt, thr = cpu_oracle_rate(workload, res, 1)
            emit(({"impl": "reference", "metric": "Mpixels/sec fwd+bwd (rasterize+interpolate+texture+antialias) at %d^2" % res,
                              "value": rate, "unit": "Mpixel/s", "n_gpus": args.gpus, "steps": 1, "warmup": 0, "ms_per_step": dt * 1e3,
                              "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
                              "config": {"workload": workload, "res": res, "note": "CPU oracle port; reference CUDA build unavailable"},
                              "cpu_baseline": {"value": rate, "unit": "Mpixel/s", "cores": thr, "kind": "port", "sample": "1 image %dx%d" % (res, res)},
                              "e2e": {"value": rate, "unit": "Mpixel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        # stdout carries exactly one JSON line: keep NCCL's banner / debug text on stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        os.environ.setdefault("NCCL_DEBUG", "WARN")
        torch.distributed.init_process_group("nccl", device_id=device)

    pipe = Pipeline(mod, workload, n_img, res, device, seed=rank)
    P = n_img * res * res

    # ---- device-resident timing
    r = timed_run(pipe, K, W, world, rank, local, device, args.impl, args.min_seconds)
    ms_step = r["ms_per_step"]
    value = world * P / (ms_step * 1e-3) / 1e6

    # ---- end to end: inputs from pinned host memory every step, loss + vertex gradient read back every step
    e2e = None
    if not args.no_e2e:
        ms2, rounds2 = e2e_run(pipe, K, world, device, args.min_seconds)
        e2e = {"value": world * P / (ms2 * 1e-3) / 1e6, "unit": "Mpixel/s", "ms_per_step": ms2, "timed_rounds": rounds2,
               "h2d_bytes_per_step": pipe.h2d_bytes(), "d2h_bytes_per_step": pipe.d2h_bytes()}

    if rank != 0:
        if world > 1:
            torch.distributed.destroy_process_group()
        return

    full = pipe.full
    out = {"metric": ("Mpixels/sec fwd+bwd (rasterize+interpolate+texture+antialias) at %d^2" if full else "Mpixels/sec fwd+bwd (rasterize+interpolate) at %d^2") % res,
           "value": value, "unit": "Mpixel/s", "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_step,
           "timed_rounds": r["rounds"], "round_ms_per_step": r["round_ms"],
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
           "config": {"workload": WORKLOAD_TEXT[workload] % (res, n_img),
                      "global_batch": world * n_img, "pixels_per_step": world * P,
                      "l2": "per-step working set (>= %.1f GB of per-pixel tensors) exceeds the 126 MB L2; no explicit flush" % (P * 32 / 1e9),
                      "parallelism": "dp%d (batch shard; all shared-parameter gradients all-reduced every step: %d bytes per rank, "
                                     "the texture gradient overlapped with the rest of backward)" % (world, pipe.reducer.bytes_per_step)},
           "comm_bytes_per_step": pipe.reducer.bytes_per_step if world > 1 else 0,
           "clocks": r["clocks"]}
    if args.impl == "reference":
        out["impl"] = "reference"
        out["reference_note"] = "unmodified NVlabs/nvdiffrast v0.4.0 CUDA extension built for sm_100a (oracle/_ref), same harness; " \
                                "the reference has no CPU path (SURVEY.md F1)"
    if e2e is not None:
        out["e2e"] = e2e
    if args.impl == "ours":
        out["gpu_launches"] = r["launches"]
        peak, peak_src = hbm_peak()
        A, D, C = (2, 2, 3) if full else (3, 0, 3)
        nbytes = op_bytes("c3" if full else workload, P, A, D, C)
        Kt = K * r["rounds"]
        table, ops = ops_table(r["trace"], Kt, nbytes, peak)
        out["ops"] = table
        dom = max((n for n in ops if n in nbytes), key=lambda n: ops[n][0])
        avg = ops[dom][0] / ops[dom][1]
        ach = nbytes[dom] / (avg * 1e-3) / 1e9
        tr = measured_traffic().get(dom.replace("b200dr_", ""))
        out["roofline"] = {"kernel": dom.replace("b200dr_", ""), "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s",
                           "frac": ach / peak, "traffic": (tr["dram_bytes_per_pixel"] * P if tr else None),
                           "traffic_source": (tr.get("source") if tr else None), "peak_source": peak_src,
                           "algorithmic_bytes_per_launch": nbytes[dom], "avg_launch_ms": avg}
        if dom == "b200dr_texture_chain_grad":
            # what the HBM fraction of this kernel means: the fusion removed 152 of the 196 B/px its three ops moved, the arithmetic stayed
            out["roofline"]["note"] = ("fused texture+interpolate+rasterize backward: bound by instruction issue, not HBM (ncu, "
                                       "profiles/r02_ncu_chain_grad_strips.md: 638 warp-instructions per 32 pixels, issue slots 68 % busy at the "
                                       "20 warps/SM its 96 registers allow, DRAM 21 %); the three ops it replaces take 2.95 ms for the same work")
        chain_bytes = sum(nbytes[n] * ops[n][1] / Kt for n in ops if n in nbytes)
        out["chain_roofline"] = {"algorithmic_bytes_per_step": chain_bytes, "achieved_GBps": chain_bytes / (ms_step * 1e-3) / 1e9,
                                 "frac": chain_bytes / (ms_step * 1e-3) / 1e9 / peak}
        fused = any(n in ops for n in replaced_bytes(nbytes))
        out["fusion"] = {"enabled": bool(mod.get_fusion()), "chain_backward_fused": fused,
                         "note": ("texture+interpolate+rasterize backward run as ONE kernel that keeps g_uv/g_uv_da/g_rast/g_rast_db in registers: "
                                  "it moves 44 of the 196 B/px the three reference ops move and is bound by instruction issue, not HBM; "
                                  if full else
                                  "interpolate+rasterize backward run as ONE kernel that keeps g_rast in registers: 28 of 76 B/px; ") +
                                 "`ops_unfused` times the op-by-op path (B200DR_FUSION=0) on the same inputs" if fused else None}
        if fused:
            api_bytes = op_by_op_bytes(nbytes, ops, Kt)
            out["chain_roofline"]["op_by_op_bytes_per_step"] = api_bytes
            out["chain_roofline"]["effective_frac_vs_op_by_op_bytes"] = api_bytes / (ms_step * 1e-3) / 1e9 / peak
        if fused and world == 1 and not args.no_extras:
            mod.set_fusion(False)
            try:
                ru = timed_run(pipe, K, 1, world, rank, local, device, args.impl, min(args.min_seconds, 0.6))
            finally:
                mod.set_fusion(True)
            tu, opsu = ops_table(ru["trace"], K * ru["rounds"], nbytes, peak)
            cb = sum(nbytes[n] * opsu[n][1] / (K * ru["rounds"]) for n in opsu if n in nbytes)
            out["ops_unfused"] = {"ms_per_step": ru["ms_per_step"], "value": P / (ru["ms_per_step"] * 1e-3) / 1e6, "clocks": ru["clocks"],
                                  "ops": {k: {"ms": v["ms"], "frac": v.get("frac")} for k, v in tu.items()},
                                  "chain_roofline": {"algorithmic_bytes_per_step": cb, "frac": cb / (ru["ms_per_step"] * 1e-3) / 1e9 / peak}}
        if world == 1 and not args.no_extras and workload == "c3":
            del pipe
            torch.cuda.empty_cache()
            out["target_kernels"] = target_kernels(mod, device, peak)
            out["other_configs"] = {name: side_config(mod, name, K, W, device, peak) for name in ("c2", "c4", "c5")}
            out["other_configs"]["envphong_shaped"] = envphong_shaped(mod, device, peak)
            out["other_configs"]["cuda_graph_launch_bound"] = graph_leg_subprocess()
    if not args.no_cpu_baseline and world == 1:
        ns = {"c3": 12, "c5": 12, "c2": 16, "c4": 1}[workload]       # bounded sample: ~10 s of host time
        rate, dt, thr = cpu_oracle_rate(workload, res, ns)
        out["cpu_baseline"] = {"value": rate, "unit": "Mpixel/s", "cores": thr, "kind": "port",
                               "sample": "%d image(s) %dx%d of the same workload, fwd+bwd, %.1f s" % (ns, res, res, dt)}
    emit(out)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
