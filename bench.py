#!/usr/bin/env python
"""Headline benchmark: rendered rays/s of the fused SDF renderer (BASELINE.json `metric`).

    python bench.py --gpus 1 --steps 20 --warmup 5            # our arm, N=1
    torchrun --nproc-per-node N ... bench.py --gpus N ...     # our arm, N ranks (one per GPU, no data-path collective)
    python bench.py --impl reference --steps 3 --warmup 1     # the reference's CPU path (oracle port) on host cores

A step = one `NeuSRenderer.render` forward over one batch of synthetic rays of BASELINE config 2:
64x64-ray patch x 64 samples/ray, D=8, W=128 FiLM-SIREN SDF + colour MLP (sphere_init SDF weights from the
committed fixture), `bs` object instances per GPU (R = bs * 4096 rays per step).  Prints ONE JSON line.
"""
import argparse
import json
import os
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

PATCH, N_SAMPLES, N_IMPORTANCE, DEPTH, WIDTH = 64, 64, 0, 8, 128
FLOP_PER_POINT = 494848          # fwd 115200 + reverse 115072 + colour 17152 MAC, x2
BYTES_PER_RAY_IN, BYTES_PER_RAY_OUT, BYTES_PER_POINT = 32, 24, 60   # SURVEY.md 8(d), dict contract A
FP32_FMA_LANES_PER_SM = 128


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm_gbs=d["hbm_gbs"], bf16_tflops=d["bf16_tflops"], sm_max_mhz=d.get("sm_max_mhz", 1965.0),
                    source="measured")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, sm_max_mhz=1965.0, source="fallback")


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU through NVML during the timed region."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {"hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40,
                     "hw_power_brake": 0x80, "sw_power_cap": 0x4}
            while not self._stop_evt.is_set():
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
                time.sleep(0.05)
        except Exception as e:  # noqa: BLE001
            self.reasons.add(f"nvml_unavailable:{type(e).__name__}")

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def make_inputs(bs, seed):
    from helpers import load_params
    from oracle import neus_oracle as O  # input generator only (synthetic_rays); not on the timed path
    P = load_params("params_D8.npz")
    ro, rd, near, far = O.synthetic_rays(bs, PATCH, seed=seed)
    z = torch.randn(bs, 64, generator=torch.Generator().manual_seed(seed))
    return P, ro, rd, near, far, z


def run_reference_arm(args):
    """The reference's own CPU implementation of the path = the oracle port (the reference is Python/torch and
    cannot travel to the GPU box; the port is pinned bit-exactly to it by tests/test_oracle.py)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import neus_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    P, ro, rd, near, far, z = make_inputs(1, 1234)
    w = O.style_mlp(P, z)
    R = ro.shape[0]

    def step():
        with torch.no_grad():
            return O.render(P, ro, rd, near, far, w=w, n_samples=N_SAMPLES, n_importance=N_IMPORTANCE,
                            cos_anneal_ratio=1.0)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    val = R * args.steps / dt
    line = {
        "impl": "reference", "metric": "rendered_rays_per_sec", "value": val, "unit": "rays/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(1, args.gpus),
        "cpu_baseline": {"value": val, "unit": "rays/s", "cores": cores, "kind": "port",
                         "sample": f"{R} rays (1 instance, 64x64 patch) x 64 samples per step, torch-CPU fp32, "
                                   f"{torch.get_num_threads()} threads"},
        "e2e": {"value": val, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(bs, n_gpus):
    return {"workload": f"cfg2: 64x64-ray patch x 64 samples/ray (n_importance=0), FiLM-SIREN D=8 W=128 SDF+colour "
                        f"MLP + analytic normal + NeuS compositing, {bs} instance(s)/GPU = {bs * PATCH * PATCH} rays/step/GPU",
            "patch": PATCH, "n_samples": N_SAMPLES, "n_importance": N_IMPORTANCE, "D": DEPTH, "W": WIDTH,
            "instances_per_gpu": bs, "parallelism": f"dp{n_gpus} (independent instances per rank, no forward collective)",
            "l2": "flushed between timed steps by a 256 MiB write"}


def cpu_baseline_leg(P, n_runs=2):
    from oracle import neus_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    _, ro, rd, near, far, z = make_inputs(1, 1234)
    w = O.style_mlp(P, z)
    ts = []
    with torch.no_grad():
        O.render(P, ro[:512], rd[:512], near[:512], far[:512], w=w, n_samples=N_SAMPLES, cos_anneal_ratio=1.0)
        for _ in range(n_runs):
            t0 = time.perf_counter()
            O.render(P, ro, rd, near, far, w=w, n_samples=N_SAMPLES, n_importance=N_IMPORTANCE, cos_anneal_ratio=1.0)
            ts.append(time.perf_counter() - t0)
    t = sorted(ts)[len(ts) // 2]
    return {"value": ro.shape[0] / t, "unit": "rays/s", "cores": cores, "kind": "port",
            "sample": f"{ro.shape[0]} rays x 64 samples (1 instance of the workload), median of {n_runs} runs, "
                      f"torch-CPU fp32 oracle port, {torch.get_num_threads()} threads"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--kernel", default="auto", choices=["auto", "ffma", "tcgen05"])
    ap.add_argument("--bs", type=int, default=4, help="object instances per GPU per step (4 = the reference's "
                    "largest single training chunk, generator.py:14,289)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return run_reference_arm(args)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- object_intrinsics_b200 has no CPU path (use --impl reference "
                         "for the CPU baseline)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    from object_intrinsics_b200 import fields
    from object_intrinsics_b200.renderer import NeuSRenderer
    bs = args.bs
    P, ro, rd, near, far, z = make_inputs(bs, 1234 + rank)
    sdf, col, devn = fields.build_networks(D=DEPTH, device=dev)
    fields.load_flat_params(sdf, col, devn, P)
    renderer = NeuSRenderer(nerf=None, sdf_network=sdf, deviation_network=devn, color_network=col,
                            n_samples=N_SAMPLES, n_importance=N_IMPORTANCE, n_outside=0, up_sample_steps=1, perturb=0,
                            impl=args.kernel)
    R = ro.shape[0]
    N = R * N_SAMPLES
    host = [t.pin_memory() for t in (ro, rd, near, far, z)]
    d_ro, d_rd, d_near, d_far, d_z = [t.to(dev) for t in host]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def step_resident():
        with torch.no_grad():
            w = sdf.style(d_z)
            return renderer.render(d_ro, d_rd, d_near, d_far, cos_anneal_ratio=1.0, perturb_overwrite=0, z=d_z, w=w)

    # end-to-end leg: host (pinned) buffers in, rendered patch + mask out, through the public API.  Two result
    # buffers so that the host consumes step i-1 while the device works on step i (a streaming consumer).
    h_color = [torch.empty((R, 3), dtype=torch.float32).pin_memory() for _ in range(2)]
    h_mask = [torch.empty((R, 1), dtype=torch.float32).pin_memory() for _ in range(2)]
    e2e_done = [torch.cuda.Event(), torch.cuda.Event()]
    e2e_checksum = [0.0]

    def step_e2e(i):
        with torch.no_grad():
            a = [t.to(dev, non_blocking=True) for t in host]
            w = sdf.style(a[4])
            out = renderer.render(a[0], a[1], a[2], a[3], cos_anneal_ratio=1.0, perturb_overwrite=0, z=a[4], w=w)
            h_color[i & 1].copy_(out["color_fine"], non_blocking=True)
            h_mask[i & 1].copy_(out["weight_sum"], non_blocking=True)
            e2e_done[i & 1].record()

    def consume_e2e(i):
        e2e_done[i & 1].synchronize()
        e2e_checksum[0] += float(h_mask[i & 1][0, 0]) + float(h_color[i & 1][0, 0])   # the host reads the result

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_resident()
    step_e2e(0)
    consume_e2e(0)
    barrier()

    sampler = ClockSampler(local_rank)
    sampler.start()
    # ---- timed region: K steps, device-timed, inputs resident in HBM
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    stops = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    cores = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for pair in cores:
        for e in pair:
            e.record()   # torch creates the cudaEvent_t lazily; the C-ABI needs the handle
    barrier()
    # K steps are enqueued back to back (no host sync inside the timed region); each step is bracketed by its own
    # event pair, the L2 flush between steps sits outside the brackets
    for i in range(args.steps):
        flush.fill_(i & 0xFF)          # evict L2 between timed iterations (not timed)
        renderer.core_events = cores[i]
        starts[i].record()
        step_resident()
        stops[i].record()
    barrier()
    renderer.core_events = None
    core_ms = [a_.elapsed_time(b_) for a_, b_ in cores]
    total_ms = sum(s.elapsed_time(e) for s, e in zip(starts, stops))
    # ---- end-to-end: host buffers in, rendered patch + mask out, copies inside the timed region
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        step_e2e(i)
        if i > 0:
            consume_e2e(i - 1)
    consume_e2e(args.steps - 1)
    barrier()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop()

    # whole-job numbers: units summed over ranks / max elapsed time over ranks (no data-path collective)
    from object_intrinsics_b200.parallel import aggregate_throughput
    value, _, total_s = aggregate_throughput(R * args.steps, total_ms * 1e-3)
    e2e_value, _, _ = aggregate_throughput(R * args.steps, e2e_s)
    total_ms = total_s * 1e3
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- side measurement (rank 0, outside every timed region): the grad-mode render of the generator update,
    # forward + loss + hand-written backward (oi_render_backward), one instance of the workload
    r1 = R // bs
    bwd_ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
    for e in bwd_ev:
        e.record()
    renderer.bwd_events = bwd_ev
    gparams = list(sdf.parameters()) + list(col.parameters()) + list(devn.parameters())

    def grad_step():
        for p_ in gparams:
            p_.grad = None
        w = sdf.style(d_z[:1])
        out = renderer.render(d_ro[:r1], d_rd[:r1], d_near[:r1], d_far[:r1], cos_anneal_ratio=1.0,
                              perturb_overwrite=0, z=d_z[:1], w=w)
        img = out["color_fine"] + (1.0 - out["weight_sum"])
        ((img ** 2).mean() + 0.1 * out["gradient_error"]).backward()

    for _ in range(3):
        grad_step()
    torch.cuda.synchronize()
    gs = []
    for i in range(5):
        flush.fill_(i)
        a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a_.record()
        grad_step()
        b_.record()
        b_.synchronize()
        gs.append(a_.elapsed_time(b_))
    renderer.bwd_events = None
    gs.sort()
    grad_info = {"ms": gs[len(gs) // 2], "rays": r1, "bwd_kernels_ms": bwd_ev[0].elapsed_time(bwd_ev[1]),
                 "what": "grad-mode render of 1 instance: forward + loss + oi_render_backward (sweep kernel on tcgen05 "
                         "+ TMA-fed TF32 point-contraction), gradients on every nn.Parameter"}

    peaks = load_peaks()
    core_avg_ms = sum(core_ms) / len(core_ms)
    used_tc = args.kernel in ("auto", "tcgen05")   # auto resolves to the tcgen05 core for depth >= 2
    flops = N * FLOP_PER_POINT
    achieved_tflops = flops / (core_avg_ms * 1e-3) / 1e12
    if used_tc:
        peak, bound, peak_note = peaks["bf16_tflops"], "tensor", f"cuBLAS bf16 burst, {peaks['source']}"
    else:
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        peak = sms * FP32_FMA_LANES_PER_SM * 2 * peaks["sm_max_mhz"] * 1e6 / 1e12
        bound, peak_note = "fp32_fma", f"nominal {sms} SMs x 128 lanes x 2 x {peaks['sm_max_mhz']:.0f} MHz"
    kernel_name = "render_tc_kernel" if used_tc else "render_ffma_kernel"
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            tr = json.load(f)[kernel_name]
        traffic = tr["dram_bytes_per_launch"] * (R / tr["rays_per_launch"])   # ncu capture, scaled to this launch
    except Exception:  # noqa: BLE001
        pass
    alg_bytes = R * (BYTES_PER_RAY_IN + BYTES_PER_RAY_OUT) + N * BYTES_PER_POINT
    hbm_gbs = alg_bytes / (core_avg_ms * 1e-3) / 1e9
    line = {
        "metric": "rendered_rays_per_sec", "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32" if not used_tc else "f32 (fp16x2-split tcgen05 products, fp32 accumulate)",
        "data": "synthetic", "config": workload_config(bs, world),
        "e2e": {"value": e2e_value, "unit": "rays/s",
                "h2d_bytes_per_step": sum(t.numel() * 4 for t in host), "d2h_bytes_per_step": R * 16},
        "gpu_launches": args.steps * (renderer.last_launches + 1),
        "roofline": {"bound": bound, "kernel": kernel_name,
                     "achieved": achieved_tflops, "peak": peak, "unit": "TFLOP/s", "frac": achieved_tflops / peak,
                     "peak_source": peak_note, "traffic": traffic,
                     "core_kernel_ms": core_avg_ms, "algorithmic_flop_per_launch": flops,
                     "hbm": {"achieved": hbm_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                             "frac": hbm_gbs / peaks["hbm_gbs"], "algorithmic_bytes_per_launch": alg_bytes,
                             "note": "path is compute-bound (8100 FLOP/B); reported because BASELINE north_star asks"}},
        "clocks": clocks,
        "grad_step": grad_info,
    }
    if not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline_leg(P)
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
