#!/usr/bin/env python
"""Headline benchmark: rendered rays/s of the fused SDF renderer (BASELINE.json `metric`).

    python bench.py --gpus 1 --steps 20 --warmup 5            # our arm, N=1
    torchrun --nproc-per-node N ... bench.py --gpus N ...     # our arm, N ranks (one per GPU)
    python bench.py --impl reference --steps 3 --warmup 1     # the reference's CPU path (oracle port) on host cores

A step = one `NeuSRenderer.render` forward over one batch of synthetic rays of BASELINE config 2:
64x64-ray patch x 64 samples/ray, D=8, W=128 FiLM-SIREN SDF + colour MLP (sphere_init SDF weights from the
committed fixture), `bs` object instances per GPU (R = bs * 4096 rays per step).  Prints ONE JSON line.

Legs of the repo arm (all in the one line):
  value / ms_per_step   K forward steps, inputs resident in HBM, CUDA events, max over ranks        [headline bs]
  bs1                   the same at 1 instance per GPU (the 64x64-patch unit of the metric)
  e2e                   host (pinned) rays in, rendered patch + mask out, copies inside the timed region
  roofline              core kernel timed live by CUDA events recorded inside the C-ABI call
  grad_step             rank 0: grad-mode render + loss + oi_render_backward of 1 instance
  grad_step_ddp         ALL ranks: the same step through DistributedDataParallel (scripts/train.py:157-158): one
                        grad render + backward per rank, gradient all-reduce over NCCL/NVLink inside backward();
                        rank 0 checks that .grad equals the mean of the per-rank gradients
  train_step            rank 0: a training-shaped composition (gan_pose_trainer.py:77-101), see train_step_leg()
  cpu_baseline          rank 0, N=1: the oracle port on the host cores (bounded sample)
The repo arm imports nothing from oracle/ except inside cpu_baseline_leg() / run_reference_arm().
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

import bench_inputs as BI  # noqa: E402

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
    P = BI.load_flat_params("params_D8.npz")
    ro, rd, near, far = BI.synthetic_rays(bs, PATCH, seed=seed)
    return P, ro, rd, near, far, BI.latent(bs, seed)


def workload_config(bs, n_gpus):
    return {"workload": f"cfg2: 64x64-ray patch x 64 samples/ray (n_importance=0), FiLM-SIREN D=8 W=128 SDF+colour "
                        f"MLP + analytic normal + NeuS compositing, {bs} instance(s)/GPU = {bs * PATCH * PATCH} rays/step/GPU",
            "patch": PATCH, "n_samples": N_SAMPLES, "n_importance": N_IMPORTANCE, "D": DEPTH, "W": WIDTH,
            "instances_per_gpu": bs, "parallelism": f"dp{n_gpus} (independent instances per rank, no forward collective)",
            "l2": "flushed between timed steps by a 256 MiB write"}


# ------------------------------------------------------------------------------------------------------------------
# CPU legs (the only places that execute oracle/)
# ------------------------------------------------------------------------------------------------------------------
def run_reference_arm(args):
    """The reference's own CPU implementation of the path = the oracle port (the reference is Python/torch and
    cannot travel to the GPU box; the port is pinned bit-exactly to it by tests/test_oracle.py).  Same `config` as
    the repo arm; each step renders a BOUNDED SAMPLE of that workload: one of its `bs` instances (4096 rays x 64)."""
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
        "config": workload_config(args.bs, args.gpus),
        "cpu_baseline": {"value": val, "unit": "rays/s", "cores": cores, "kind": "port",
                         "sample": f"per step ONE instance of the workload's {args.bs} ({R} rays = one 64x64 patch, x 64 "
                                   f"samples), torch-CPU fp32 oracle port, {torch.get_num_threads()} threads"},
        "e2e": {"value": val, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def cpu_baseline_leg(n_runs=2):
    from oracle import neus_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    P, ro, rd, near, far, z = make_inputs(1, 1234)
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


# ------------------------------------------------------------------------------------------------------------------
# helpers of the repo arm
# ------------------------------------------------------------------------------------------------------------------
def lib_stamp():
    p = os.path.join(ROOT, "object_intrinsics_b200", "lib", "liboi_b200.stamp")
    return open(p).read().strip() if os.path.exists(p) else None


def recorded_traffic(kernel_name, rays):
    """dram bytes per launch of the dominant kernel from the committed ncu capture -- only if that capture was
    taken on THIS build (profiles/traffic.json carries the source stamp of the library it profiled)."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            tr = json.load(f)[kernel_name]
    except Exception:  # noqa: BLE001
        return None, "no capture recorded for this kernel"
    if tr.get("stamp") != lib_stamp():
        return None, f"capture {tr.get('capture')} was taken on another build (stamp mismatch): not reported"
    return tr["dram_bytes_per_launch"] * (rays / tr["rays_per_launch"]), tr.get("capture")


def timed_forward(renderer, sdf, tensors, steps, flush, barrier):
    """K resident forward steps; returns (sum of per-step ms, list of core-kernel ms)."""
    d_ro, d_rd, d_near, d_far, d_z = tensors

    def step():
        with torch.no_grad():
            w = sdf.style(d_z)
            return renderer.render(d_ro, d_rd, d_near, d_far, cos_anneal_ratio=1.0, perturb_overwrite=0, z=d_z, w=w)
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    stops = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    cores = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for pair in cores:
        for e in pair:
            e.record()   # torch creates the cudaEvent_t lazily; the C-ABI needs the handle
    barrier()
    # K steps are enqueued back to back (no host sync inside the timed region); each step is bracketed by its own
    # event pair, the L2 flush between steps sits outside the brackets.  A few extra flushes first: the device is busy
    # with them while the host enqueues step 0, so that no bracket contains host launch latency after the barrier.
    for _ in range(6):
        flush.fill_(0xFF)
    for i in range(steps):
        flush.fill_(i & 0xFF)          # evict L2 between timed iterations (not timed)
        renderer.core_events = cores[i]
        starts[i].record()
        step()
        stops[i].record()
    barrier()
    renderer.core_events = None
    return sum(s.elapsed_time(e) for s, e in zip(starts, stops)), [a.elapsed_time(b) for a, b in cores]


def ddp_leg(args, dev, rank, world, local_rank, P, dist, flush):
    """Row e2: per rank ONE grad-mode render of its own instance (4096 rays x 64) + backward through DDP."""
    from object_intrinsics_b200 import fields
    from object_intrinsics_b200.parallel import (RenderModule, aggregate_throughput, check_ddp_gradients,
                                                  training_loss)

    def build():
        sdf, col, devn = fields.build_networks(D=DEPTH, device=dev)
        fields.load_flat_params(sdf, col, devn, P)
        return RenderModule(sdf, col, devn, n_samples=N_SAMPLES, n_importance=N_IMPORTANCE, impl=args.kernel)
    local, wrapped = build(), build()
    ddp = torch.nn.parallel.DistributedDataParallel(wrapped, device_ids=[local_rank])
    _, ro, rd, near, far, z = make_inputs(1, 1234 + rank)
    inputs = tuple(t.to(dev) for t in (ro, rd, near, far, z))
    err, n_grad = check_ddp_gradients(ddp, local, inputs)
    del local
    steps = max(5, min(args.steps, 20))

    def step():
        for p in ddp.parameters():
            p.grad = None
        training_loss(ddp(*inputs)).backward()
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    dist.barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for i in range(steps):
        flush.fill_(i & 0xFF)
        ev[i][0].record()
        step()
        ev[i][1].record()
    torch.cuda.synchronize()
    dist.barrier()
    ms = sum(a.elapsed_time(b) for a, b in ev)
    R = ro.shape[0]
    value, _, total_s = aggregate_throughput(R * steps, ms * 1e-3)
    return {"ms_per_step": total_s * 1e3 / steps, "value": value, "unit": "rays/s (grad-mode render + backward + "
            "gradient all-reduce, whole job)", "rays_per_step_per_gpu": R, "steps": steps, "world_size": world,
            "allreduce_bytes_per_step": n_grad * 4, "grad_vs_rank_mean_rel_err": err,
            "nccl_env": {k: os.environ.get(k) for k in ("NCCL_P2P_LEVEL", "NCCL_IB_DISABLE")},
            "what": "RenderModule (sdf/color/deviation networks + renderer) wrapped in DistributedDataParallel("
                    "device_ids=[local_rank]) as scripts/train.py:157-158; every rank renders its own instance; "
                    "time = max over ranks, CUDA events"}


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
    ap.add_argument("--no-side-legs", action="store_true", help="skip grad_step / grad_step_ddp / train_step / bs1")
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
    import torch.distributed as dist
    from object_intrinsics_b200.parallel import aggregate_throughput, nccl_nvlink_env
    nccl_nvlink_env()                 # NVLink/NVSwitch only (NCCL_P2P_LEVEL=NVL, NCCL_IB_DISABLE=1)
    if world == 1:                    # plain `python bench.py`: a one-rank group so that the DDP leg runs at N=1 too
        os.environ.setdefault("MASTER_PORT", str(29400 + os.getpid() % 500))
        os.environ.setdefault("RANK", "0")
        os.environ.setdefault("WORLD_SIZE", "1")
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
    # the five input tensors live back to back in ONE pinned host buffer: one host->device copy per step
    sizes = [t.numel() for t in (ro, rd, near, far, z)]
    host_flat = torch.empty(sum(sizes), dtype=torch.float32).pin_memory()
    host, off = [], 0
    for t, n_el in zip((ro, rd, near, far, z), sizes):
        view = host_flat[off:off + n_el].view(t.shape)
        view.copy_(t)
        host.append(view)
        off += n_el
    resident = [t.to(dev) for t in host]
    d_ro, d_rd, d_near, d_far, d_z = resident
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    # end-to-end leg: host (pinned) buffers in, rendered patch + mask out, through the public API.  Two result
    # buffers so that the host consumes step i-1 while the device works on step i (a streaming consumer).
    h_color = [torch.empty((R, 3), dtype=torch.float32).pin_memory() for _ in range(2)]
    h_mask = [torch.empty((R, 1), dtype=torch.float32).pin_memory() for _ in range(2)]
    e2e_done = [torch.cuda.Event(), torch.cuda.Event()]
    e2e_checksum = [0.0]

    # Inputs travel on a copy stream one step ahead of the render that consumes them (a streaming producer): every
    # step still pays its own host->device copy and device->host read inside the timed region, they just overlap
    # the previous / next step's kernels instead of serialising with them.
    copy_stream = torch.cuda.Stream(device=dev)
    staged = {}

    def stage_inputs(i):
        with torch.cuda.stream(copy_stream):
            flat = host_flat.to(dev, non_blocking=True)
            a, off = [], 0
            for t, n_el in zip(host, sizes):
                a.append(flat[off:off + n_el].view(t.shape))
                off += n_el
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        staged[i] = (a, ev)

    def step_e2e(i, last=False):
        with torch.no_grad():
            if i not in staged:
                stage_inputs(i)
            a, ev = staged.pop(i)
            torch.cuda.current_stream(dev).wait_event(ev)
            if not last:
                stage_inputs(i + 1)
            a[0].record_stream(torch.cuda.current_stream(dev))   # all five are views of one allocation
            w = sdf.style(a[4])
            out = renderer.render(a[0], a[1], a[2], a[3], cos_anneal_ratio=1.0, perturb_overwrite=0, z=a[4], w=w)
            h_color[i & 1].copy_(out["color_fine"], non_blocking=True)
            h_mask[i & 1].copy_(out["weight_sum"], non_blocking=True)
            e2e_done[i & 1].record()

    def consume_e2e(i):
        e2e_done[i & 1].synchronize()
        e2e_checksum[0] += float(h_mask[i & 1][0, 0]) + float(h_color[i & 1][0, 0])   # the host reads the result

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        with torch.no_grad():
            renderer.render(d_ro, d_rd, d_near, d_far, cos_anneal_ratio=1.0, perturb_overwrite=0, z=d_z,
                            w=sdf.style(d_z))
    step_e2e(0, last=True)
    consume_e2e(0)
    barrier()

    sampler = ClockSampler(local_rank)
    sampler.start()
    # ---- timed region: K steps, device-timed, inputs resident in HBM
    total_ms, core_ms = timed_forward(renderer, sdf, resident, args.steps, flush, barrier)
    # ---- end-to-end: host buffers in, rendered patch + mask out, copies inside the timed region
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        step_e2e(i, last=(i == args.steps - 1))
        if i > 0:
            consume_e2e(i - 1)
    consume_e2e(args.steps - 1)
    barrier()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop()
    # per-rank view (every rank samples its own GPU): slowest rank's step time, lowest median SM clock, union of the
    # throttle reasons -- at N > 1 `value` is bounded by the slowest GPU of the box
    per_rank = None
    if world > 1:
        mine = torch.tensor([total_ms / args.steps, float(sum(core_ms) / len(core_ms)), float(clocks["sm_mhz"] or 0),
                             float(sum(1 << i for i, k in enumerate(("hw_slowdown", "sw_thermal_slowdown",
                                                                     "hw_thermal_slowdown", "hw_power_brake",
                                                                     "sw_power_cap")) if k in clocks["reasons"]))],
                            dtype=torch.float64, device=dev)
        allr = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = {"ms_per_step": [round(float(t[0]), 4) for t in allr],
                    "core_kernel_ms": [round(float(t[1]), 4) for t in allr],
                    "sm_mhz": [int(t[2]) for t in allr], "throttle_bits": [int(t[3]) for t in allr],
                    "throttle_bit_names": ["hw_slowdown", "sw_thermal_slowdown", "hw_thermal_slowdown",
                                           "hw_power_brake", "sw_power_cap"]}

    # whole-job numbers: units summed over ranks / max elapsed time over ranks (no data-path collective)
    value, _, total_s = aggregate_throughput(R * args.steps, total_ms * 1e-3)
    e2e_value, _, _ = aggregate_throughput(R * args.steps, e2e_s)
    total_ms = total_s * 1e3

    # ---- the same forward at ONE instance per GPU (the 64x64-patch unit the metric is named after)
    bs1 = None
    if not args.no_side_legs:
        r1 = R // bs
        one = [d_ro[:r1], d_rd[:r1], d_near[:r1], d_far[:r1], d_z[:1]]
        for _ in range(args.warmup):      # new tensor shapes: let the caching allocator settle before timing
            with torch.no_grad():
                renderer.render(*one[:4], cos_anneal_ratio=1.0, perturb_overwrite=0, z=one[4], w=sdf.style(one[4]))
        ms1, core1 = timed_forward(renderer, sdf, one, args.steps, flush, barrier)
        v1, _, t1 = aggregate_throughput(r1 * args.steps, ms1 * 1e-3)
        bs1 = {"value": v1, "unit": "rays/s", "ms_per_step": t1 * 1e3 / args.steps,
               "core_kernel_ms": sum(core1) / len(core1), "rays_per_step_per_gpu": r1}

    # ---- row e2: DDP grad step on every rank
    ddp_info = None
    if not args.no_side_legs:
        ddp_info = ddp_leg(args, dev, rank, world, local_rank, P, dist, flush)
    barrier()
    if rank != 0:
        dist.destroy_process_group()
        return

    line_extra = {}
    if not args.no_side_legs:
        line_extra["contract_b"] = contract_b_leg(renderer, sdf, resident, bs, flush, args.steps)
        line_extra["grad_step"] = grad_step_leg(renderer, sdf, col, devn, resident, R // bs, flush)
        try:
            line_extra["train_step"] = train_step_leg(dev, P, args.kernel, flush)
        except Exception as e:  # noqa: BLE001  (a side measurement must not take the headline down with it)
            line_extra["train_step"] = {"error": f"{type(e).__name__}: {e}"}

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
    traffic, traffic_note = recorded_traffic(kernel_name, R)
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
                     "peak_source": peak_note, "traffic": traffic, "traffic_source": traffic_note,
                     "split_ceiling_frac": (achieved_tflops / (peak / 3.0)) if used_tc else None,
                     "core_kernel_ms": core_avg_ms, "algorithmic_flop_per_launch": flops,
                     "hbm": {"achieved": hbm_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                             "frac": hbm_gbs / peaks["hbm_gbs"], "algorithmic_bytes_per_launch": alg_bytes,
                             "note": "path is compute-bound (8100 FLOP/B); reported because BASELINE north_star asks"}},
        "clocks": clocks,
        "per_rank": per_rank,
        "bs1": bs1,
        "grad_step_ddp": ddp_info,
        "lib_stamp": lib_stamp(),
    }
    line.update(line_extra)
    if not args.no_cpu_baseline and world == 1:
        line["cpu_baseline"] = cpu_baseline_leg()
    print(json.dumps(line), flush=True)
    dist.destroy_process_group()


def contract_b_leg(renderer, sdf, resident, bs, flush, steps):
    """Rank 0: the same rays through contract B -- render + Generator.render_maps in ONE library call
    (`OiRenderDesc.maps`): shading and the maps are composited in the tile tail of the render kernel, no per-point
    tensor is written; outputs per ray: 6 maps (14 floats) + weight_sum / weight_max / color_fine / s_val."""
    d_ro, d_rd, d_near, d_far, d_z = resident
    dev = d_ro.device
    lp = torch.tensor([0.33, 0.33, 0.33, 0.67, 0.67, 0.67, 0.01, 0.01, 0.01, 10.0], device=dev)   # lighting.py:9-12
    ldir = torch.tensor([[0.0, 0.0, -1.0]], device=dev).repeat(bs, 1)
    bg = torch.ones(bs, 3, device=dev)

    def step():
        with torch.no_grad():
            return renderer.render_with_maps(d_ro, d_rd, d_near, d_far, w=sdf.style(d_z), light_params=lp,
                                             light_dir=ldir, bg_color=bg, resolution=PATCH, cos_anneal_ratio=1.0,
                                             perturb_overwrite=0)
    def timed():
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for i in range(steps):
            flush.fill_(i & 0xFF)
            ev[i][0].record()
            step()
            ev[i][1].record()
        torch.cuda.synchronize()
        return sum(a.elapsed_time(b) for a, b in ev) / steps, renderer.last_launches
    R = d_ro.shape[0]
    ms, nl = timed()                       # default: compositing in the render kernel, maps kernel after it
    flags = renderer.flags
    renderer.flags = flags | 16            # maps composited in the tile tail as well: nothing per-point is written
    try:
        ms_k, nl_k = timed()
    finally:
        renderer.flags = flags
    return {"ms_per_step": ms, "value": R / (ms * 1e-3), "unit": "rays/s", "launches_per_step": nl,
            "output_bytes_per_ray": 14 * 4 + 6 * 4,
            "in_kernel_maps": {"ms_per_step": ms_k, "value": R / (ms_k * 1e-3), "launches_per_step": nl_k,
                               "per_point_bytes_written": 0},
            "what": "render + render_maps in one C-ABI call (OiRenderDesc.maps): no per-point tensor is handed to the "
                    "caller; default = per-point tensors through the workspace + maps kernel, in_kernel_maps = flags "
                    "bit 4, maps composited in the render kernel's tile tail"}


def grad_step_leg(renderer, sdf, col, devn, resident, r1, flush):
    """Rank 0, outside every timed region: the grad-mode render of the generator update, forward + loss +
    hand-written backward (oi_render_backward), one instance of the workload."""
    d_ro, d_rd, d_near, d_far, d_z = resident
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
    gs, ks = [], []
    for i in range(5):
        flush.fill_(i)
        a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a_.record()
        grad_step()
        b_.record()
        b_.synchronize()
        gs.append(a_.elapsed_time(b_))
        ks.append(bwd_ev[0].elapsed_time(bwd_ev[1]))
    renderer.bwd_events = None
    gs.sort()
    ks.sort()
    # free-running: the loop of a trainer (no host read-back between steps -- gan_pose_trainer.py has none), L2 flush
    # fills in stream between the steps and their measured duration subtracted
    n_free = 10
    a_, b_, c_ = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    torch.cuda.synchronize()
    a_.record()
    for i in range(n_free):
        flush.fill_(i)
    b_.record()
    for i in range(n_free):
        flush.fill_(i)
        grad_step()
    c_.record()
    c_.synchronize()
    free_ms = (b_.elapsed_time(c_) - a_.elapsed_time(b_)) / n_free
    return {"ms": gs[len(gs) // 2], "ms_free_running": free_ms, "rays": r1, "bwd_kernels_ms": ks[len(ks) // 2],
            "what": "grad-mode render of 1 instance: forward + loss + oi_render_backward (sweep kernel on tcgen05 "
                    "+ TMA-fed point-contraction on fp16 / TF32 operands), gradients on every nn.Parameter; `ms` = median of steps "
                    "synchronised one by one (the host's enqueue time is exposed), `ms_free_running` = 10 steps "
                    "back to back as a training loop issues them (L2 flushed between steps, flush time subtracted)"}


def train_step_leg(dev, P, kernel, flush):
    """Rank 0: a training-shaped composition of the path's kernels (no trainer code), mirroring
    src/trainers/gan_pose_trainer.py:77-101 -- implemented in tools_train_step.py so that it can be profiled alone."""
    import tools_train_step as T
    return {s: T.measure(dev, P, kernel, flush, s) for s in ("cfg5", "cfg3")}


if __name__ == "__main__":
    main()
