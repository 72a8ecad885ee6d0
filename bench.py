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
import datetime
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
# stdout carries exactly one JSON line: NCCL's own banner ("NCCL version ...", printed when NCCL_DEBUG is set in the
# environment) goes to stderr
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

import torch  # noqa: E402

RAYS_PER_GPU = 8192
SAMPLES = 128
BONES = 25
# algorithmic work, SURVEY.md 8(d): trunk 601600 + 2 x 47840 MAC per sample forward, x3 for fwd+dgrad+wgrad;
# the roofline numerator uses the conservative figure with per-ray-constant inputs hoisted (501.4 MFLOP/ray)
FLOP_PER_RAY_FULL = 535.5e6
FLOP_PER_RAY_HOISTED = 501.4e6
TRUNK_MAC_PER_SAMPLE = 601600 - 91 * 128   # nerf_coarse, one pass, per-ray-constant dir/env columns hoisted


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


_REAL_STDOUT = None


def _quiet_stdout():
    """stdout carries exactly ONE line, the JSON result: until _emit() everything written to file descriptor 1 (NCCL
    prints its "NCCL version ..." banner there at the first collective, whatever NCCL_DEBUG_FILE says) goes to stderr."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def _emit(obj):
    global _REAL_STDOUT
    sys.stdout.flush()
    if _REAL_STDOUT is not None:
        os.dup2(_REAL_STDOUT, 1)
        os.close(_REAL_STDOUT)
        _REAL_STDOUT = None
    print(json.dumps(obj), flush=True)


def _traffic(entries):
    """DRAM bytes (read + write) per launch of the dominant kernel, from the committed `ncu --set full` capture
    (profiles/traffic.json, written by tools/ncu_summary.py --traffic); null when there is none."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(p):
        return None, None
    d = json.load(open(p))
    vals = [d["kernels"][k]["dram_bytes_per_launch"] for k in entries if k in d.get("kernels", {})]
    if not vals:
        return None, None
    return sum(vals) / len(vals), d.get("source")


def _step_dram(per_entry):
    """DRAM bytes of one training step as far as the committed ncu captures cover it: per-launch bytes of profiles/traffic.json
    x the launches per step counted here (the multi-job weight-gradient launches by their shape).  Returns (bytes, covered
    kernels, C-ABI entries without a capture)."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(p):
        return None, [], []
    k = json.load(open(p)).get("kernels", {})
    n = lambda e: per_entry.get(e, (0, 0))[0]
    # launches per step of each captured kernel: the trunk pass has one 256x256 multi-job weight-gradient launch, one
    # 256x64 (PE inputs) and one single 128x256 (direction layer); each nerf_skin evaluation one 64x64 multi-job launch
    plan = [("moda_chain_trunk_fwd", n("moda_chain_trunk_fwd")), ("moda_chain_trunk_bwd", n("moda_chain_trunk_bwd")),
            ("moda_chain_skin_fwd", n("moda_chain_skin_fwd")), ("moda_chain_skin_bwd", n("moda_chain_skin_bwd")),
            ("moda_tc_wgrad_multi<256,256>", n("moda_chain_trunk_bwd")), ("moda_tc_wgrad_multi<256,64>", n("moda_chain_trunk_bwd")),
            ("moda_tc_wgrad<128,256>", n("moda_tc_wgrad")), ("moda_tc_wgrad_multi<64,64>", n("moda_chain_skin_bwd")),
            ("moda_skin_warp_fwd", n("moda_skin_warp_fwd")), ("moda_skin_warp_bwd", n("moda_skin_warp_bwd"))]
    tot, have, miss = 0.0, [], []
    for key, cnt in plan:
        if cnt and key in k:
            tot += k[key]["dram_bytes_per_launch"] * cnt
            have.append(key)
        elif cnt:
            miss.append(key)
    return tot, have, miss


class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU during the timed region (B200_PROFILING.md clocks line).
    NVML in a thread (a query takes well under a millisecond, so a 60 ms timed region still yields a dozen samples);
    falls back to polling `nvidia-smi -lms` when the NVML binding is unavailable."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    BITS = {"sw_power_cap": 0x4, "hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}

    def __init__(self, index):
        self.index, self.rows, self.proc, self.nvml, self.stop_flag = index, [], None, None, False
        self.sm, self.mx, self.reasons, self.th = [], None, set(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            # LOCAL_RANK indexes the visible devices; map through CUDA_VISIBLE_DEVICES when it is a plain index list
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = index
            if vis:
                ids = [v.strip() for v in vis.split(",")]
                if index < len(ids) and ids[index].isdigit():
                    phys = int(ids[index])
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _poll_nvml(self):
        n = self.nvml
        while not self.stop_flag:
            try:
                self.sm.append(float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)))
                try:
                    r = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                except Exception:
                    r = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                for name, bit in self.BITS.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.004)

    def start(self):
        if self.nvml is not None:
            self.th = threading.Thread(target=self._poll_nvml, daemon=True)
            self.th.start()
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.nvml is not None:
            self.stop_flag = True
            self.th.join(timeout=1)
            sm = sorted(self.sm)
            return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.mx, "reasons": sorted(self.reasons),
                    "samples": len(sm), "source": "nvml"}
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
                "samples": len(sm), "source": "nvidia-smi"}


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
        # a mismatched collective should fail within minutes, not hold the box for the default 10
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))
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
    # AdamW on the flat buffer (FlatParams.adamw_step -> moda_adamw_flat: torch.optim.AdamW's arithmetic and defaults,
    # checked against it in tests/test_gpu_round2b.py; graph-capturable)
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
        flat.adamw_step(lr=1e-4)
        return loss

    # end-to-end input pipeline: two device buffer sets; while step i computes, the host->device copy of step i+1's
    # rays runs on a copy stream from pinned memory (every step's inputs are copied inside the timed region, one copy
    # set per step), and the step's loss is read back to pinned host memory
    copy_stream = torch.cuda.Stream(device=dev)
    bufs = [{k: torch.empty_like(rays[k]) for k in ray_keys} for _ in range(2)]
    ready_ev = [torch.cuda.Event() for _ in range(2)]
    free_ev = [torch.cuda.Event() for _ in range(2)]
    pipe = {"i": 0, "primed": False}

    def issue_copy(slot):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(free_ev[slot])      # the step that last read this buffer set has finished
            for k in ray_keys:
                bufs[slot][k].copy_(host[k], non_blocking=True)
            ready_ev[slot].record(copy_stream)

    def step_e2e():
        slot = pipe["i"] & 1
        if not pipe["primed"]:
            issue_copy(slot)
            pipe["primed"] = True
        issue_copy(slot ^ 1)
        cur = torch.cuda.current_stream()
        cur.wait_event(ready_ev[slot])
        rd = {"xys": rays["xys"]}
        for k in ray_keys:
            t = bufs[slot][k].detach()
            if k in ("bone_rts", "time_embedded", "env_code", "rays_o", "rays_d"):
                t.requires_grad_(True)
            rd[k] = t
        loss = step(rd)
        free_ev[slot].record(cur)
        loss_host.copy_(loss.detach().reshape(1), non_blocking=True)
        pipe["i"] += 1
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, sampler=None):  # noqa: E306
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

    def capture(rays_static):
        """The whole step (zero-grad, forward, backward, gradient all-reduce, AdamW) as ONE CUDA graph over static
        input buffers: ~180 C-ABI launches + ~100 small framework kernels replayed with a single launch call.  Every
        rank captures (the all-reduce is part of the graph).  Returns (graph, loss tensor, launches per step) or the
        reason why capture was not possible (the eager path is then timed instead)."""
        try:
            barrier()
            g = torch.cuda.CUDAGraph()
            n0 = _lib.LAUNCHES
            with torch.cuda.graph(g, capture_error_mode="thread_local"):
                loss = step(rays_static)
            n = _lib.LAUNCHES - n0
            for _ in range(2):
                g.replay()
            barrier()
            if not bool(torch.isfinite(loss.detach()).all()):
                return "graph replay produced a non-finite loss"
            return g, loss, n
        except Exception as e:  # noqa: BLE001 -- any capture failure selects the eager path
            try:
                torch.cuda.synchronize()
            except Exception:
                pass
            return "%s: %s" % (type(e).__name__, str(e).splitlines()[0][:200])

    def agree(ok):
        """Capture is a collective decision: every rank replays, or none does."""
        t = torch.tensor([1 if ok else 0], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return bool(int(t))

    use_graph = args.graph != "off"
    try:   # the autograd nodes of the flat parameter views were created on the default stream: expected, not a problem
        torch.autograd.graph.set_warn_on_accumulate_grad_stream_mismatch(False)
    except Exception:
        pass
    # BASELINE configs[2] (strong scaling): the SAME 8192-ray global batch sharded N ways, measured next to the weak-
    # scaling headline (8192 rays per GPU) in the same run
    strong = None
    if world > 1 and R % world == 0:
        Rs = R // world
        sub = {k: v[rank * Rs:(rank + 1) * Rs].detach().clone() for k, v in rays.items()}
        for k in ("bone_rts", "time_embedded", "env_code", "rays_o", "rays_d"):
            sub[k].requires_grad_(True)
        for _ in range(max(args.warmup, 3)):
            step(sub)
        ms_strong, _ = timed(lambda: step(sub), args.steps)
        strong = {"global_rays": R, "rays_per_gpu": Rs, "ms_per_step_eager": round(ms_strong, 3), "cuda_graph": False}
        if use_graph:
            cap = capture(sub)
            if agree(not isinstance(cap, str)):
                ms_g, _ = timed(cap[0].replay, args.steps)
                strong["cuda_graph"] = True
                ms_strong = min(ms_strong, ms_g)
                strong["ms_per_step_graph"] = round(ms_g, 3)
            else:
                strong["cuda_graph_unavailable"] = cap if isinstance(cap, str) else "another rank could not capture"
            del cap
        strong.update({"ms_per_step": round(ms_strong, 3), "value": round(R / (ms_strong * 1e-3), 1), "unit": "rays/s",
                       "note": "same step on the 8192-ray global batch sharded over the ranks (BASELINE configs[2]); "
                               "efficiency vs N=1 = value / (N=1 value of the headline metric)"})
    for _ in range(max(args.warmup, 3)):
        step(rays)
    launches0 = _lib.LAUNCHES
    ms_step, clocks = timed(lambda: step(rays), args.steps, ClockSampler(local) if rank == 0 else None)
    launches = (_lib.LAUNCHES - launches0) // args.steps
    for _ in range(2):
        step_e2e()
    ms_e2e, _ = timed(step_e2e, args.steps)
    graph_info = {"used": False}
    if use_graph:
        # static input set for the graph; the e2e variant copies each step's host rays into a staging set on the copy
        # stream (as above) and from there into the static set with device-to-device copies on the compute stream
        static = {k: (rays[k].detach().clone().requires_grad_(rays[k].requires_grad) if torch.is_tensor(rays[k]) else rays[k])
                  for k in rays}
        cap = capture(static)
        if agree(not isinstance(cap, str)):
            g, g_loss, g_launches = cap
            ms_g, clocks_g = timed(g.replay, args.steps, ClockSampler(local) if rank == 0 else None)

            def step_e2e_graph():
                slot = pipe["i"] & 1
                if not pipe["primed"]:
                    issue_copy(slot)
                    pipe["primed"] = True
                issue_copy(slot ^ 1)
                cur = torch.cuda.current_stream()
                cur.wait_event(ready_ev[slot])
                with torch.no_grad():
                    for k in ray_keys:
                        static[k].copy_(bufs[slot][k], non_blocking=True)
                free_ev[slot].record(cur)
                g.replay()
                loss_host.copy_(g_loss.detach().reshape(1), non_blocking=True)
                pipe["i"] += 1

            pipe["primed"] = False
            for _ in range(2):
                step_e2e_graph()
            ms_e2e_g, _ = timed(step_e2e_graph, args.steps)
            graph_info = {"used": True, "ms_per_step_eager": round(ms_step, 3), "ms_per_step_graph": round(ms_g, 3),
                          "e2e_ms_per_step_eager": round(ms_e2e, 3), "e2e_ms_per_step_graph": round(ms_e2e_g, 3),
                          "kernels": "the step captured once as a CUDA graph (zero-grad, forward, backward, NCCL "
                                     "all-reduce, AdamW) and replayed; same kernels, same work"}
            if ms_g < ms_step:
                ms_step, clocks, launches = ms_g, clocks_g, g_launches
            ms_e2e = min(ms_e2e, ms_e2e_g)
        else:
            graph_info = {"used": False, "unavailable": cap if isinstance(cap, str) else "another rank could not capture"}

    # per-kernel device time of one step (CUDA events around every C-ABI call on the launching stream)
    roof = None
    # every rank runs these two steps (step() contains the gradient all-reduce: a collective issued by rank 0 alone
    # would never complete); only rank 0 records events and summarises
    # (single-stream launches for these two steps: with the weight-gradient kernels alternating over two streams the
    # per-call event intervals overlap and would be counted twice)
    from moda_b200 import config as _cfg
    side_was = _cfg.side_stream
    _cfg.side_stream = False
    if rank == 0:
        _lib.PROFILE = {}
    for _ in range(2):
        step(rays)
    _cfg.side_stream = side_was
    if rank == 0:
        summ = _lib.profile_summary()
        _lib.PROFILE = None
        peaks, src = _peaks()
        peak = peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"])
        per_entry = {k: (n / 2.0, ms / 2.0) for k, (n, ms) in summ.items()}   # launches and ms per step
        tot_ms = sum(ms for _, ms in per_entry.values())
        # dominant kernel: chain_kernel<128, 8, 1, 1, *> -- one persistent tcgen05 launch per pass of nerf_coarse
        # (forward chain and adjoint chain are the same kernel template).  Algorithmic work of one launch: the hoisted
        # layer MACs of the pass, 2 x 589952 FLOP per sample.
        dom = ("moda_chain_trunk_fwd", "moda_chain_trunk_bwd")
        dom_n = sum(per_entry[k][0] for k in dom if k in per_entry)
        dom_ms = sum(per_entry[k][1] for k in dom if k in per_entry)
        P = R * SAMPLES
        flop_launch = 2.0 * TRUNK_MAC_PER_SAMPLE * P
        achieved = flop_launch / (dom_ms / max(dom_n, 1) * 1e-3) / 1e12 if dom_ms > 0 else 0.0
        traffic, traffic_src = _traffic(dom)
        lin_ms = sum(ms for k, (n, ms) in per_entry.items() if k.startswith(("moda_linear", "moda_tc", "moda_chain")))
        agg = FLOP_PER_RAY_HOISTED * R / (lin_ms * 1e-3) / 1e12
        roof = {"bound": "tensor", "kernel": "chain::chain_kernel<128,8,1,1,*> (moda_chain_trunk_fwd / moda_chain_trunk_bwd)",
                "achieved": round(achieved, 3), "peak": peak, "unit": "TFLOP/s", "frac": round(achieved / peak, 5),
                "peak_source": src + " bf16 sustained (kernel timed inside the step)",
                "flop_per_launch": flop_launch, "launches_per_step": dom_n, "ms_per_launch": round(dom_ms / max(dom_n, 1), 4),
                "share_of_step": round(dom_ms / tot_ms, 4), "traffic": traffic, "traffic_source": traffic_src,
                "timing": "CUDA events around each C-ABI call on the launching stream, 2 steps after the timed region",
                "step_aggregate": {"kernels": "all linear-layer kernels of nerf_coarse + nerf_skin (fwd, dgrad, wgrad)",
                                   "flop_per_step": FLOP_PER_RAY_HOISTED * R, "ms_per_step": round(lin_ms, 3),
                                   "achieved": round(agg, 3), "frac": round(agg / peak, 5),
                                   "share_of_step": round(lin_ms / tot_ms, 4)},
                "all_kernels_ms_per_step": round(tot_ms, 3),
                # the chains run xyz_encoding_final folded into dir_encoding: `achieved` keeps the
                # ALGORITHMIC count of SURVEY 8(d) as its numerator, the tensor cores execute 65 536 MAC per sample less
                "executed_flop_per_launch": 2.0 * (TRUNK_MAC_PER_SAMPLE - (65536 if _cfg.fold_final else 0)) * P,
                "per_entry_ms": {k: round(ms, 3) for k, (n, ms) in sorted(per_entry.items(), key=lambda kv: -kv[1][1])}}
        sd, sd_have, sd_miss = _step_dram({k: v for k, v in per_entry.items()})
        if sd is not None:
            hbm = peaks.get("hbm_gbs")
            roof["step_dram_bytes"] = sd
            roof["step_dram_note"] = ("ncu per-launch DRAM bytes x launches per step; captured: %s; not captured: %s%s"
                                      % (", ".join(sd_have), ", ".join(sd_miss) or "-",
                                         ("; = %.2f ms at the measured %.0f GB/s" % (sd / hbm / 1e6, hbm)) if hbm else ""))
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_baseline(sample_rays=args.cpu_rays, repeats=2)
    # secondary figures of BASELINE.json's metric, folded into the same line so that the driver's BENCH / SCALE files
    # carry them: the DQ-skinning microbench (configs[3], one GPU), the density grid (configs[4], sharded over the ranks) and
    # the default-flag MoDA step (one GPU)
    extra = {}
    if not args.no_extra:
        torch.cuda.empty_cache()
        g = measure_grid(args, dev, world, rank, with_cpu=False, steps=3)
        if rank == 0:
            extra["grid"] = g
        if world == 1:
            extra["dqs"] = measure_dqs(args, dev, with_cpu=False, steps=3)
            try:   # the default-flag MoDA step (`--workload full` is the stand-alone form)
                torch.cuda.empty_cache()
                extra["full"] = measure_full(args, dev, steps=5)
            except Exception as e:   # a secondary figure must not take the headline line down with it
                extra["full"] = {"error": "%s: %s" % (type(e).__name__, e)}
    if rank == 0:
        rays_s = R * world / (ms_step * 1e-3)
        out = {"metric": "train rays/s (128 samp/ray)", "value": round(rays_s, 1), "unit": "rays/s",
               "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": round(ms_step, 3),
               "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
               "dtype": "f16 operands / f32 accumulate (tcgen05 MLP chains; nerf_skin forward split hi+lo fp16); f32 elsewhere",
               "data": "synthetic",
               "config": {"workload": "full fwd+bwd training step of render_rays: %d rays/GPU x %d samples, %d bones, "
                                      "8x256 nerf_coarse + 5x64 nerf_skin, AdamW step, grad all-reduce" % (R, SAMPLES, BONES),
                          "rays_per_gpu": R, "samples_per_ray": SAMPLES, "bones": BONES, "parallelism": "dp%d" % world,
                          "l2_policy": "inputs+activations per step (>10 GB) exceed the 126 MB L2",
                          "launch_modes": {"trunk_chains_as_cta_pairs": bool(_cfg.trunk_pair and _lib.lib().moda_chain_pair_available()),
                                           "trunk_tiles_in_flight_per_cta": _cfg.trunk_slots if _cfg.trunk_pair else 1,
                                           "wgrad_side_stream": bool(_cfg.side_stream)}},
               "e2e": {"value": round(R * world / (ms_e2e * 1e-3), 1), "unit": "rays/s", "h2d_bytes_per_step": h2d_bytes,
                       "d2h_bytes_per_step": 4, "ms_per_step": round(ms_e2e, 3),
                       "input_pipeline": "pinned host rays -> double-buffered H2D prefetch on a copy stream, one copy set per step"},
               "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "cpu_baseline": cpu,
               "strong_scaling": strong, "cuda_graph": graph_info, "extra": extra}
        _emit(out)
    _finish(world)


def _finish(world):
    """Leaves the process without tearing NCCL down: destroying a process group while captured graphs that contain its
    collectives are still alive can block forever (seen at N = 2), and nothing is left to do after the JSON line."""
    sys.stdout.flush()
    sys.stderr.flush()
    if world > 1:
        import torch.distributed as dist
        try:
            dist.barrier()
            torch.cuda.synchronize()
        except Exception:
            pass
        os._exit(0)


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
    _emit(({"impl": "reference", "metric": "train rays/s (128 samp/ray)", "value": v, "unit": "rays/s",
                      "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": args.steps, "warmup": min(args.warmup, 1),
                      "ms_per_step": round(dt * 1e3, 2), "higher_is_better": True, "scaling": "weak",
                      "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                      "config": {"workload": "full fwd+bwd training step of render_rays: %d rays x %d samples, %d bones "
                                             "(CPU sample of the 8192-ray workload)" % (R, SAMPLES, BONES)},
                      "cpu_baseline": {"value": v, "unit": "rays/s", "cores": cores, "kind": "port", "sample": sample},
                      "e2e": {"value": v, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


# ---------------------------------------------------------------------------------------------------------------
# Secondary workloads (BASELINE.json configs[3] and configs[4]); the default line stays the training step.

def _event_ms(fn, steps, warmup):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def _traffic_of(key):
    """DRAM bytes per launch of one kernel from the committed ncu capture (profiles/traffic.json); None if absent."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(p):
        return None
    k = json.load(open(p)).get("kernels", {}).get(key)
    return k["dram_bytes_per_launch"] if k else None


def measure_dqs(args, dev, with_cpu=True, steps=None):
    """BASELINE configs[3]: 16.78 M points (131072 rays x 128) x 25 Gaussian bones, per-ray dual quaternions;
    Gaussian skinning + backward warp, then skinning + forward warp (geom_utils.py:202-302, 372-517), forward and
    backward timed separately, without and with streamed delta-logits (pitch 32 fp32).  Returns the JSON object."""
    from moda_b200 import synth, geom_utils as G, _lib
    assert _lib.lib().moda_device_check() == 0, _lib.lib().moda_last_error()
    steps = steps or args.steps
    R, S, B = args.dqs_rays, SAMPLES, BONES
    sp = synth.make_skin_problem(R, S, seed=0)
    xyz_host = sp["xyz"].pin_memory()
    rts_host = sp["bone_rts"].pin_memory()
    xyz = sp["xyz"].to(dev).requires_grad_(True)
    bones = sp["bones_rst"].to(dev).requires_grad_(True)
    aux = sp["skin_aux"].to(dev).requires_grad_(True)
    rts = sp["bone_rts"].to(dev).requires_grad_(True)
    P = R * S
    peaks, src = _peaks()
    out = {}
    sampler = ClockSampler(dev.index or 0)
    sampler.start()
    for tag, with_delta in (("no_delta", False), ("delta_streamed", True)):
        d_bw = d_fw = None
        if with_delta:
            d_bw = (0.1 * torch.randn(R, S, 32, device=dev)).requires_grad_(True)
            d_fw = (0.1 * torch.randn(R, S, 32, device=dev)).requires_grad_(True)
            with torch.no_grad():
                d_bw[..., B:] = 0
                d_fw[..., B:] = 0
        state = {}

        def fwd():
            can = G.warp_points(xyz, bones, rts, aux, d_bw, backward=True)
            cyc = G.warp_points(can, bones, rts, aux, d_fw, backward=False)
            state["y"] = (can, cyc)

        g1, g2 = torch.randn(R, S, 3, device=dev), torch.randn(R, S, 3, device=dev)

        def bwd():
            can, cyc = state["y"]
            leaves = [xyz, bones, aux, rts] + ([d_bw, d_fw] if with_delta else [])
            torch.autograd.grad([can, cyc], leaves, [g1, g2], retain_graph=True)

        ms_f = _event_ms(fwd, steps, max(args.warmup, 3))
        ms_b = _event_ms(bwd, steps, max(args.warmup, 3))
        # algorithmic bytes per point: forward 12 in + 2 x 12 out + 1800/128 per-ray bone data
        # [+ 2 x 25 x 4 delta]; backward: xyz, xyz_can, 2 gradients in, 1 out = 60 + 14 [+ 2 x 2 x 25 x 4 read+write]
        bf = 50.1 + (200.0 if with_delta else 0.0)
        bb = 74.0 + (400.0 if with_delta else 0.0)
        out[tag] = {"fwd_ms": round(ms_f, 3), "bwd_ms": round(ms_b, 3),
                    "fwd_gpts_s": round(P / ms_f / 1e6, 2), "bwd_gpts_s": round(P / ms_b / 1e6, 2),
                    "fwd_hbm_frac": round(bf * P / (ms_f * 1e-3) / 1e9 / peaks["hbm_gbs"], 4),
                    "bwd_hbm_frac": round(bb * P / (ms_b * 1e-3) / 1e9 / peaks["hbm_gbs"], 4),
                    "alg_bytes_per_point": {"fwd": bf, "bwd": bb}}
        del state, d_bw, d_fw
    # end to end through the public API with HOST buffers: points + per-ray transforms copied in, both warped point
    # sets copied out, inside the timed region
    can_host = torch.empty(R, S, 3).pin_memory()
    cyc_host = torch.empty(R, S, 3).pin_memory()

    def e2e():
        with torch.no_grad():
            x = xyz_host.to(dev, non_blocking=True)
            r = rts_host.to(dev, non_blocking=True)
            can = G.warp_points(x, bones, r, aux, None, backward=True)
            cyc = G.warp_points(can, bones, r, aux, None, backward=False)
            can_host.copy_(can, non_blocking=True)
            cyc_host.copy_(cyc, non_blocking=True)

    ms_e = _event_ms(e2e, steps, 2)
    clocks = sampler.stop()
    cpu = None
    if with_cpu:
        from oracle import restated as O
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        Rc = 2048
        spc = synth.make_skin_problem(Rc, S, seed=0)
        best = None
        for i in range(3):
            t0 = time.perf_counter()
            with torch.no_grad():
                O.skin_warp_roundtrip(spc["bones_rst"], spc["bone_rts"].reshape(Rc, B, 8), spc["skin_aux"], spc["xyz"])
            dt = time.perf_counter() - t0
            if i:
                best = dt if best is None else min(best, dt)
        cpu = {"value": round(Rc * S / best / 1e9, 5), "unit": "Gpts/s", "cores": cores, "kind": "port",
               "sample": "%d points forward (both warps, no delta), best of 2 after 1 warm-up" % (Rc * S)}
    a = out["delta_streamed"]
    return {"metric": "DQ skinning Gpts/s (bw + fw warp, 25 Gaussian bones)", "value": out["no_delta"]["fwd_gpts_s"],
            "unit": "Gpts/s", "n_gpus": 1, "steps": steps, "warmup": max(args.warmup, 3),
            "ms_per_step": out["no_delta"]["fwd_ms"], "higher_is_better": True, "scaling": "replicas only",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "DQ skinning microbench: %d points x %d bones, fw+bw warp" % (P, B),
                       "l2_policy": "working set (>0.6 GB) exceeds the 126 MB L2"},
            "variants": out,
            "e2e": {"value": round(P / ms_e / 1e6, 3), "unit": "Gpts/s", "ms_per_step": round(ms_e, 3),
                    "h2d_bytes_per_step": int(xyz_host.numel() * 4 + rts_host.numel() * 4),
                    "d2h_bytes_per_step": int(2 * can_host.numel() * 4),
                    "note": "forward (both warps, no delta) from pinned host points to pinned host results"},
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "skin_warp_fwd_kernel x2 with streamed delta logits",
                         "achieved": round(a["fwd_hbm_frac"] * peaks["hbm_gbs"], 1), "peak": peaks["hbm_gbs"],
                         "unit": "GB/s", "frac": a["fwd_hbm_frac"], "peak_source": src,
                         "traffic": _traffic_of("skin_warp_fwd_delta")},
            "cpu_baseline": cpu}


FULL_LOSS_KEYS = ("img_loss_samp", "sil_loss_samp", "flo_loss_samp", "feat_err", "proj_err", "frnd_loss_samp", "frame_cyc_dis")


def measure_full(args, dev, steps=None):
    """The DEFAULT-flag MoDA training step: core path + nerf_feat feature rendering + Sinkhorn
    feature matching + key-point reprojection + third warp with flow rendering + nerf_vis loss + per-ray terms
    (nnutils/moda.py:344-348, 447-449; rendering.py:405-578), 8192 rays x 128 samples, fwd + bwd + AdamW, one GPU."""
    from moda_b200 import synth, models as MM, _lib
    from moda_b200.parallel import FlatParams
    from moda_b200.rendering import render_rays
    steps = steps or args.steps
    R = args.rays
    prob = synth.make_full_problem(R, seed=0)
    models, emb, rays = MM.build_full_models(prob, dev)
    for k in ("coarse", "nerf_skin", "nerf_feat", "nerf_vis"):
        models[k].train()
    opts = synth.full_opts()
    flat = FlatParams(MM.parameters_of(models))
    bound = prob["obj_bound"].numpy()

    def step():
        flat.zero_grad()
        res = render_rays(models, emb, rays, N_samples=SAMPLES, perturb=1.0, noise_std=0.0, chunk=32768, obj_bound=bound,
                          img_size=prob["img_size"], opts=opts)
        loss = res["vis_loss"]
        for k in FULL_LOSS_KEYS:
            loss = loss + res[k].mean()
        loss.backward()
        flat.allreduce()
        flat.adamw_step(lr=1e-4)
        return loss

    sampler = ClockSampler(dev.index or 0)
    n0 = _lib.LAUNCHES
    step()
    launches = _lib.LAUNCHES - n0
    sampler.start()
    ms = _event_ms(step, steps, max(args.warmup, 3))
    clocks = sampler.stop()
    _lib.PROFILE = {}
    step()
    summ = _lib.profile_summary()
    _lib.PROFILE = None
    loss = float(step().detach())
    return {"metric": "train rays/s, default-flag step (128 samp/ray)", "value": round(R / (ms * 1e-3), 1), "unit": "rays/s",
            "n_gpus": 1, "steps": steps, "warmup": max(args.warmup, 3), "ms_per_step": round(ms, 3), "higher_is_better": True,
            "data": "synthetic", "loss_finite": bool(loss == loss and abs(loss) < 1e30), "gpu_launches": int(launches),
            "config": {"workload": "default-flag MoDA step: %d rays x %d samples, %d bones, nerf_coarse + nerf_skin + nerf_feat "
                                   "(5x128) + nerf_vis (5x64), Sinkhorn feature matching on a 20^3 lattice, reprojection, "
                                   "third warp + flow rendering, AdamW" % (R, SAMPLES, BONES)},
            "clocks": clocks,
            "per_entry_ms": {k: round(v[1], 3) for k, v in sorted(summ.items(), key=lambda kv: -kv[1][1])[:12]}}


def run_full(args):
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    _emit(measure_full(args, dev))


def run_dqs(args):
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    _emit(measure_dqs(args, dev, with_cpu=not args.no_cpu))


def measure_grid(args, dev, world, rank, with_cpu=True, steps=None):
    """BASELINE configs[4]: canonical density on a G^3 lattice (train_utils.py:1377-1404), x-slabs per rank, gathered
    with one all-gather.  Every rank calls this; rank 0 gets the JSON object, the others None."""
    import torch.distributed as dist
    from moda_b200 import synth, models as MM, _lib
    from moda_b200.extract import density_grid
    assert _lib.lib().moda_device_check() == 0, _lib.lib().moda_last_error()
    steps = steps or args.steps
    Gs = args.grid
    prob = synth.make_problem(8, seed=0)
    models, emb, _ = MM.build_models(prob, dev, requires_grad=False)
    lo, hi = rank * Gs // world, (rank + 1) * Gs // world
    full = torch.empty(Gs, Gs, Gs, device=dev) if world > 1 else None
    vol_host = torch.empty(Gs, Gs, Gs).pin_memory() if rank == 0 else None

    def step():
        vol = density_grid(models["coarse"], Gs, (0.3, 0.3, 0.3), emb["xyz"], x_range=(lo, hi))
        if world > 1:
            dist.all_gather_into_tensor(full, vol)
            return full
        return vol

    def step_e2e():   # what extract_mesh does next: the volume goes to the host for marching cubes
        v = step()
        if rank == 0:
            vol_host.copy_(v, non_blocking=True)

    def timed(fn):
        for _ in range(max(args.warmup, 3)):
            fn()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    sampler = ClockSampler(dev.index or 0) if rank == 0 else None
    if sampler:
        sampler.start()
    ms = timed(step)
    ms_e = timed(step_e2e)
    clocks = sampler.stop() if sampler else None
    if rank != 0:
        return None
    peaks, src = _peaks()
    pts = Gs ** 3
    flop = pts * 491264 * 2.0
    peak = peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"]) * world
    cpu = None
    if world == 1 and with_cpu:
        from oracle import restated as O
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        Gc = 48
        t0 = time.perf_counter()
        with torch.no_grad():
            O.density_grid(prob["coarse"], Gc, (0.3, 0.3, 0.3))
        dt = time.perf_counter() - t0
        cpu = {"value": round(Gc ** 3 / dt / 1e6, 4), "unit": "Mpts/s", "cores": cores, "kind": "port",
               "sample": "%d^3 grid (%.2f s), one pass" % (Gc, dt)}
    return {"metric": "density-grid Mpts/s (sigma_only nerf_coarse, %d^3)" % Gs, "value": round(pts / ms / 1e3, 1),
            "unit": "Mpts/s", "n_gpus": world, "steps": steps, "warmup": max(args.warmup, 3),
            "ms_per_step": round(ms, 3), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f16 operands, f32 accumulate", "data": "synthetic",
            "config": {"workload": "mesh-extraction density query, %d^3 lattice, x-slabs over %d rank(s)" % (Gs, world)},
            "e2e": {"value": round(pts / ms_e / 1e3, 1), "unit": "Mpts/s", "ms_per_step": round(ms_e, 3),
                    "h2d_bytes_per_step": 0, "d2h_bytes_per_step": int(pts * 4),
                    "note": "lattice generated on the device (it is a function of G and the bound); the (G,G,G) fp32 "
                            "volume is copied to pinned host memory inside the timed region"},
            "clocks": clocks,
            "roofline": {"bound": "tensor", "kernel": "chain_kernel sigma-only program", "achieved": round(flop / (ms * 1e-3) / 1e12, 2),
                         "peak": peak, "unit": "TFLOP/s", "frac": round(flop / (ms * 1e-3) / 1e12 / peak, 4),
                         "peak_source": src + " bf16 sustained", "traffic": _traffic_of("chain_trunk_sigma")},
            "cpu_baseline": cpu}


def run_grid(args):
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    out = measure_grid(args, dev, world, rank, with_cpu=not args.no_cpu)
    if rank == 0:
        _emit(out)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rays", type=int, default=RAYS_PER_GPU)
    ap.add_argument("--cpu-rays", type=int, default=512)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--graph", default="auto", choices=["auto", "off"],
                    help="auto: also time the step replayed as one CUDA graph and report the faster of the two")
    ap.add_argument("--no-extra", action="store_true", help="skip the DQ-skinning / density-grid figures of the default line")
    ap.add_argument("--workload", default="train", choices=["train", "dqs", "grid", "full"],
                    help="train: the headline training step (default); dqs / grid: BASELINE configs[3] / configs[4]")
    ap.add_argument("--dqs-rays", type=int, default=131072)
    ap.add_argument("--grid", type=int, default=256)
    args = ap.parse_args()
    _quiet_stdout()
    if args.workload == "dqs":
        return run_dqs(args)
    if args.workload == "grid":
        return run_grid(args)
    if args.workload == "full":
        return run_full(args)
    if args.impl == "reference":
        args.steps = min(args.steps, 3)
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
