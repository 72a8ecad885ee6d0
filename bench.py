"""Headline benchmark: MoDA training rays/s (full fwd+bwd step of the articulated volume renderer,
128 samples/ray, 25 bones) on N B200s, next to the CPU reference path.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # the reference algorithm on the host CPU

A "step" = render_rays(training mode, perturb=1, noise_std=0) on 8192 synthetic rays per GPU -> scalar loss ->
backward -> one NCCL all-reduce of the flat gradient buffer (N>1) -> fused AdamW step.  Prints ONE JSON line
(rank 0).  Workload: BASELINE.json configs[1] ("full fwd+bwd training step, 8192 rays x 128 samples, 25 bones,
1 B200"); for N>1 every rank renders its own 8192-ray shard (weak scaling).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

RAYS_PER_GPU = 8192
SAMPLES = 128
BONES = 25
# algorithmic work, SURVEY.md 8(d): trunk 601600 + 2 x 47840 MAC per sample forward, x3 for fwd+dgrad+wgrad;
# the roofline numerator uses the conservative figure with per-ray-constant inputs hoisted (501.4 MFLOP/ray)
FLOP_PER_RAY_FULL = 535.5e6
FLOP_PER_RAY_HOISTED = 501.4e6


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def _loss_of(res):
    return ((res["img_coarse"] - 0.3) ** 2).mean() + ((res["sil_coarse"] - 0.5) ** 2).mean() \
        + res["frame_cyc_dis"].mean()


def run_ours(args):
    import torch.distributed as dist
    from moda_b200 import synth, models as MM, _lib
    from moda_b200.parallel import FlatParams
    from moda_b200.rendering import render_rays

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert _lib.lib().moda_device_check() == 0, _lib.lib().moda_last_error()

    R = args.rays
    prob = synth.make_problem(R * world, seed=0)  # same nets on every rank; rays sharded contiguously
    for k in list(prob["rays"]):
        prob["rays"][k] = prob["rays"][k][rank * R:(rank + 1) * R].contiguous()
    models, emb, rays = MM.build_models(prob, dev)
    models["coarse"].train()
    models["nerf_skin"].train()
    opts = synth.default_opts()
    flat = FlatParams(MM.parameters_of(models))
    optim = torch.optim.AdamW([flat.flat], lr=1e-4, fused=True)
    ray_keys = ("rays_o", "rays_d", "near", "far", "time_embedded", "bone_rts", "env_code")
    host = {k: prob["rays"][k].pin_memory() for k in ray_keys}
    h2d_bytes = sum(v.numel() * 4 for v in host.values())
    loss_host = torch.zeros(1).pin_memory()
    share = 1.0 / world

    def step(rays_dev):
        flat.zero_grad()
        res = render_rays(models, emb, rays_dev, N_samples=SAMPLES, perturb=1.0, noise_std=0.0, chunk=32768,
                          img_size=512, opts=opts)
        loss = _loss_of(res) * share
        loss.backward()
        flat.allreduce()
        optim.step()
        return loss

    def step_e2e():
        rd = {"xys": rays["xys"]}
        for k in ray_keys:
            t = host[k].to(dev, non_blocking=True)
            if k in ("bone_rts", "time_embedded", "env_code", "rays_o", "rays_d"):
                t.requires_grad_(True)
            rd[k] = t
        loss = step(rd)
        loss_host.copy_(loss.detach().reshape(1), non_blocking=True)
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, sampler=None):
        barrier()
        if sampler:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        clocks = sampler.stop() if sampler else None
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) / steps, clocks

    for _ in range(max(args.warmup, 3)):
        step(rays)
    launches0 = _lib.LAUNCHES
    ms_step, clocks = timed(lambda: step(rays), args.steps, ClockSampler(local) if rank == 0 else None)
    launches = (_lib.LAUNCHES - launches0) // args.steps
    for _ in range(2):
        step_e2e()
    ms_e2e, _ = timed(step_e2e, args.steps)

    # per-kernel device time of one step (CUDA events around every C-ABI call on the launching stream)
    roof = None
    if rank == 0:
        _lib.PROFILE = {}
        for _ in range(2):
            step(rays)
        summ = _lib.profile_summary()
        _lib.PROFILE = None
        peaks, src = _peaks()
        lin_ms = sum(ms for k, (n, ms) in summ.items()
                     if k.startswith(("moda_linear", "moda_tc", "moda_chain"))) / 2
        tot_ms = sum(ms for _, ms in summ.values()) / 2
        achieved = FLOP_PER_RAY_HOISTED * R / (lin_ms * 1e-3) / 1e12
        peak = peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"])
        roof = {"bound": "tensor", "kernel": "linear layers of nerf_coarse + nerf_skin (fwd, dgrad, wgrad)",
                "achieved": round(achieved, 3), "peak": peak, "unit": "TFLOP/s", "frac": round(achieved / peak, 5),
                "peak_source": src + " bf16 sustained", "traffic": None,
                "kernel_ms_per_step": round(lin_ms, 3), "all_kernels_ms_per_step": round(tot_ms, 3),
                "share_of_step": round(lin_ms / tot_ms, 4),
                "per_entry_ms": {k: round(ms / 2, 3) for k, (n, ms) in sorted(summ.items(), key=lambda kv: -kv[1][1])}}
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_baseline(sample_rays=args.cpu_rays, repeats=2)
    if rank == 0:
        rays_s = R * world / (ms_step * 1e-3)
        out = {"metric": "train rays/s (128 samp/ray)", "value": round(rays_s, 1), "unit": "rays/s",
               "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": round(ms_step, 3),
               "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
               "data": "synthetic",
               "config": {"workload": "full fwd+bwd training step of render_rays: %d rays/GPU x %d samples, %d bones, "
                                      "8x256 nerf_coarse + 5x64 nerf_skin, AdamW step, grad all-reduce" % (R, SAMPLES, BONES),
                          "rays_per_gpu": R, "samples_per_ray": SAMPLES, "bones": BONES, "parallelism": "dp%d" % world,
                          "l2_policy": "inputs+activations per step (>10 GB) exceed the 126 MB L2"},
               "e2e": {"value": round(R * world / (ms_e2e * 1e-3), 1), "unit": "rays/s", "h2d_bytes_per_step": h2d_bytes,
                       "d2h_bytes_per_step": 4, "ms_per_step": round(ms_e2e, 3)},
               "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "cpu_baseline": cpu}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def cpu_baseline(sample_rays=512, repeats=2):
    """The oracle port of the reference algorithm (oracle/restated.py, same chunking as the reference) timed on
    the host cores: training-mode fwd+bwd on a bounded sample of the same workload."""
    from moda_b200 import synth
    from oracle import restated as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    prob = synth.make_problem(sample_rays, seed=0)
    best = None
    for i in range(repeats + 1):
        p = O.to_dtype(prob, torch.float32)
        O.require_grads(p)
        jit = torch.rand(sample_rays, SAMPLES)
        t0 = time.perf_counter()
        res = O.render_rays(p, n_samples=SAMPLES, perturb=1.0, perturb_rand=jit)
        O.parity_loss(res).backward()
        dt = time.perf_counter() - t0
        if i > 0:
            best = dt if best is None else min(best, dt)
    return {"value": round(sample_rays / best, 2), "unit": "rays/s", "cores": cores, "kind": "port",
            "sample": "%d rays x %d samples fwd+bwd, best of %d after 1 warm-up (%.2f s each), torch %s CPU fp32"
                      % (sample_rays, SAMPLES, repeats, best, torch.__version__)}


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle port; the reference is Python and cannot travel
    to the GPU box) on the host cores, same metric / config."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    R = args.cpu_rays
    from moda_b200 import synth
    from oracle import restated as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    prob = synth.make_problem(R, seed=0)

    def step():
        p = O.to_dtype(prob, torch.float32)
        O.require_grads(p)
        res = O.render_rays(p, n_samples=SAMPLES, perturb=1.0, perturb_rand=torch.rand(R, SAMPLES))
        O.parity_loss(res).backward()

    for _ in range(min(args.warmup, 1)):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    v = round(R / dt, 2)
    sample = "%d rays x %d samples per step (bounded sample of the %d-ray workload)" % (R, SAMPLES, RAYS_PER_GPU)
    print(json.dumps({"impl": "reference", "metric": "train rays/s (128 samp/ray)", "value": v, "unit": "rays/s",
                      "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": args.steps, "warmup": min(args.warmup, 1),
                      "ms_per_step": round(dt * 1e3, 2), "higher_is_better": True, "scaling": "weak",
                      "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                      "config": {"workload": "full fwd+bwd training step of render_rays: %d rays x %d samples, %d bones "
                                             "(CPU sample of the 8192-ray workload)" % (R, SAMPLES, BONES)},
                      "cpu_baseline": {"value": v, "unit": "rays/s", "cores": cores, "kind": "port", "sample": sample},
                      "e2e": {"value": v, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rays", type=int, default=RAYS_PER_GPU)
    ap.add_argument("--cpu-rays", type=int, default=512)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        args.steps = min(args.steps, 3)
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
