#!/usr/bin/env python
"""bench.py -- headline benchmark of the EncoderMap hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[3], the configuration the north-star target is quoted on; it fits one
GPU): the full-set sketch-map sigmoid cost, forward + backward, over 65 536 synthetic frames x 1 024
periodic dims -> 2-d latent.  One step = one evaluation of loss and dL/d(latent) over all pair tiles; with
N GPUs the tile list is cut into N contiguous ranges (inputs replicated) and the partial loss/gradient are
summed by ONE fused NCCL launch of libemk's own communicator -- total work is fixed, so this is STRONG scaling.
`per_rank` attributes the step time (kernel / collective incl. the wait for the slowest rank) rank by rank.

metric  : unique unordered pairs (incl. diagonal) per second, N(N+1)/2 / step time, whole job
value   : inputs resident in HBM, device-timed (CUDA events, max over ranks)
e2e     : same through the public API (`sigmoid_loss(...)(y_true, y_pred)` + backward) with HOST pinned
          buffers: H2D of both inputs and D2H of loss and gradient inside the timed region (y_true is handed
          over as the pinned host tensor; every rank streams the rows ITS tile range touches over its own host
          link, in chunks behind the pair tiles that need them)
roofline: the pair-tile kernel against the FP32 issue roofline of SURVEY.md section 8d
          (4 lane-instructions per (pair, dim) + 60 per pair;  peak = 148 SM x 128 lanes x sm_max clock from
          MEASURED_PEAKS.json, peak_measured = FFMA rate probed live on the device; the kernel is neither
          HBM- nor tensor-bound, so "bound" says "fp32-issue"; the HBM-bound back-mapping kernels carry their
          own rooflines under "extra")
cpu_baseline / --impl reference: the op-for-op float32 torch-CPU restatement of the reference (oracle/),
          all host threads, on a bounded sample (N=512 rows of the same data) -- TensorFlow is not installed
          in this image, so the reference itself cannot be run.
Extra keys report the secondary metrics (back-mapped frames/s for configs[4], per-batch cost for
configs[1]) with their own rooflines.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import statistics
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
# NCCL writes its version banner / debug log to stdout by default; stdout carries exactly one JSON line here
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

N_ROWS, N_DIMS, N_LATENT = 65536, 1024, 2
SIG = (4.5, 12, 6, 1, 2, 6)
PERIOD = 2 * math.pi
CPU_SAMPLE_ROWS = 512
BACKMAP_ATOMS, BACKMAP_FRAMES, BACKMAP_CHUNK = 1500, 1 << 20, 1 << 18


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return {"hbm_gbs": d["hbm_gbs"], "sm_max_mhz": d.get("sm_max_mhz", 1965.0), "source": "MEASURED_PEAKS.json"}
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0, "source": "fallback (B200_PROFILING.md)"}


def synth_high(n, d, device, seed):
    """cfg2/cfg4 data: 16 cluster centres U(-pi,pi)^d + N(0, 0.05^2), wrapped to (-pi, pi]."""
    import torch

    g = torch.Generator(device=device).manual_seed(seed)
    centres = (torch.rand(16, d, device=device, generator=g) * 2 - 1) * math.pi
    idx = torch.randint(0, 16, (n,), device=device, generator=g)
    x = centres[idx] + 0.05 * torch.randn(n, d, device=device, generator=g)
    x = torch.remainder(x + math.pi, 2 * math.pi) - math.pi
    z = 3 * torch.randn(n, N_LATENT, device=device, generator=g)
    return x.contiguous(), z.contiguous()


class ClockSampler:
    """Samples SM clock / throttle reasons of one GPU while the timed region runs (pynvml, 50 ms)."""

    def __init__(self, index):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake": 0x80, "sync_boost": 0x10, "app_clocks": 0x2}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.05)

    def __enter__(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thread:
            self._thread.join(timeout=2)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        busy = [s for s in self.samples if s > 0.5 * max(self.samples)] or self.samples
        return {"sm_mhz": statistics.median(busy), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def cpu_reference_value(steps, warmup, rows=CPU_SAMPLE_ROWS):
    """pairs/s of the float32 torch-CPU restatement (op for op, (N,N,D) broadcast + autograd backward) on `rows` rows."""
    import torch

    from oracle import em_oracle as O

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    x, z = synth_high(rows, N_DIMS, "cpu", 4321)
    f = O.sigmoid_loss(PERIOD, SIG)
    times = []
    for it in range(warmup + steps):
        zz = z.clone().requires_grad_(True)
        t0 = time.perf_counter()
        loss = f(x, zz)
        loss.backward()
        times.append(time.perf_counter() - t0)
    t = statistics.median(times[warmup:])
    pairs = rows * (rows + 1) / 2
    return pairs / t, t, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    value, t, cores = cpu_reference_value(max(1, args.steps), max(1, min(args.warmup, 2)))
    sample = f"N={CPU_SAMPLE_ROWS} rows x {N_DIMS} dims of the same synthetic data (the reference materialises (N,N,D): 1.07 GB per intermediate at N=512; N=65536 would need 17.6 TB)"
    line = {
        "impl": "reference", "metric": "sigmoid_cost_pairs_per_s_fwd_bwd", "value": value, "unit": "unique pairs/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "configs[3]: full-set sketch-map cost fwd+bwd, 65536 x 1024 periodic -> 2-d latent (CPU arm: bounded sample)",
                   "n_rows": N_ROWS, "n_dims": N_DIMS, "latent": N_LATENT, "periodicity": "2pi", "sig": list(SIG)},
        "cpu_baseline": {"value": value, "unit": "unique pairs/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "unique pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "TensorFlow is not installed in this image: the timed CPU path is oracle/em_oracle.py, a float32 torch-CPU restatement of the reference's ops incl. the (N,N,D) broadcast and autograd backward",
    }
    print(json.dumps(line))


def run_ours(args):
    import torch
    import torch.distributed as dist

    from encodermap_b200 import _lib, _ops
    from encodermap_b200.loss_functions import sigmoid_loss

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            sys.exit(f"--gpus {args.gpus} needs torchrun with --nproc-per-node {args.gpus}")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    group = None
    if world > 1:
        # NCCL prints its version banner with printf to stdout when the communicator comes up (NCCL_DEBUG=VERSION in the
        # launch environment); stdout must carry exactly one JSON line, so fd 1 points at stderr until NCCL is initialised
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            warm = torch.zeros(1, device=dev)
            dist.all_reduce(warm)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    L = _lib.lib()  # fails loudly if libemk.so is missing
    peaks = measured_peaks()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    from encodermap_b200 import parallel

    if world > 1:
        parallel.init_comm()    # libemk's own NCCL communicator: loss + gradient summed in ONE fused launch per step
    x, z = synth_high(N_ROWS, N_DIMS, dev, 4321)
    tr = _lib.pair_tile_range(N_ROWS, rank, world)
    pairs = N_ROWS * (N_ROWS + 1) / 2

    def step():
        loss, grad = _ops.sigmoid_cost_raw(x, z, PERIOD, SIG, tr, True)
        if world > 1:
            parallel.allreduce_cost(loss, grad)
        return loss, grad

    for _ in range(args.warmup):
        step()
    barrier()
    marks = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        e0.record()
        for _ in range(args.steps):
            k0, k1, c1 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            k0.record()
            loss, grad = _ops.sigmoid_cost_raw(x, z, PERIOD, SIG, tr, True)
            k1.record()
            if world > 1:
                parallel.allreduce_cost(loss, grad)
            c1.record()
            marks.append((k0, k1, c1))
        e1.record()
        barrier()
    total_ms = max_over_ranks(e0.elapsed_time(e1))
    ms_per_step = total_ms / args.steps
    value = pairs / (ms_per_step * 1e-3)
    kernel_ms = max_over_ranks(statistics.mean(a.elapsed_time(b) for a, b, _ in marks))
    loss_value = float(loss.item())
    # attribution of the step time: per rank and per step, the kernel and the collective (incl. the wait for the slowest rank)
    per_rank = None
    multi_gpu_check = None
    if world > 1:
        mine = torch.tensor([[a.elapsed_time(b), b.elapsed_time(c)] for a, b, c in marks], dtype=torch.float64, device=dev)
        allm = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allm, mine)
        allm = torch.stack(allm).cpu()               # (rank, step, {kernel, collective})
        kern, coll = allm[:, :, 0], allm[:, :, 1]
        per_rank = {"kernel_ms_mean": [round(v, 4) for v in kern.mean(1).tolist()],
                    "kernel_ms_max": [round(v, 4) for v in kern.max(1).values.tolist()],
                    "collective_ms_mean": [round(v, 4) for v in coll.mean(1).tolist()],
                    "collective_ms_min": [round(v, 4) for v in coll.min(1).values.tolist()],
                    "step_max_kernel_ms_mean": round(kern.max(0).values.mean().item(), 4),
                    "skew_ms_mean": round((kern.max(0).values - kern.min(0).values).mean().item(), 4),
                    "note": "collective_ms of a rank = its wait for the slowest rank of that step + the fused NCCL launch; "
                            "collective_ms_min over steps and ranks is the cost of the collective itself"}
        # the sharded result equals the one-GPU result (rank 0 evaluates every tile once, outside the timed region)
        if rank == 0:
            lfull, gfull = _ops.sigmoid_cost_raw(x, z, PERIOD, SIG, None, True)
            multi_gpu_check = {"loss_rel_diff": abs(loss_value - lfull.item()) / lfull.item(),
                               "grad_rel_diff": ((grad - gfull).norm() / gfull.norm()).item()}
            del lfull, gfull
        barrier()

    # ---- end to end: host pinned buffers in, loss + gradient out, through the public API --------------------
    xh = x.cpu().pin_memory()
    zh = z.cpu().pin_memory()
    gh = torch.empty_like(zh).pin_memory()
    f = sigmoid_loss(None, periodicity_overwrite=PERIOD, dist_dig_parameters_overwrite=SIG,
                     process_group=(dist.group.WORLD if world > 1 else None), check_finite=False)

    def e2e_step():
        # the public API takes the pinned host tensor itself: every rank streams the rows ITS tile range touches over its
        # own host link, in chunks behind the pair tiles that need them (no exchange of inputs between GPUs)
        zd = zh.to(dev, non_blocking=True).requires_grad_(True)
        l = f(xh, zd)
        l.backward()
        gh.copy_(zd.grad, non_blocking=True)
        return l.item()

    for _ in range(min(args.warmup, 2)):
        e2e_step()
    barrier()
    e0.record()
    for _ in range(args.steps):
        e2e_loss = e2e_step()
    e1.record()
    barrier()
    e2e_ms = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    # the end-to-end path (host input, streamed / interleaved over ranks) returns the same loss and gradient as the device path
    e2e_check = {"loss_rel_diff": abs(e2e_loss - loss_value) / abs(loss_value),
                 "grad_rel_diff": ((gh.to(dev) - grad).norm() / grad.norm()).item()}
    e2e_value = pairs / (e2e_ms * 1e-3)
    # bytes this rank copies in per step: the rows from the band of its first tile to the end + the latent
    if world > 2:   # 1/G of the rows per host link + NVLink all-gather (sigmoid_loss picks this above two ranks)
        h2d_mine = -(-N_ROWS // world) * N_DIMS * 4 + zh.numel() * 4
    else:
        first_row = min(N_ROWS, (_lib.pair_tile_decode(N_ROWS, tr[0])[0] * 128 // 8192) * 8192) if tr[1] > tr[0] else N_ROWS
        h2d_mine = (N_ROWS - first_row) * N_DIMS * 4 + zh.numel() * 4
    h2d_max = max_over_ranks(float(h2d_mine))

    extra = {}
    if rank == 0 and not args.no_extra:
        extra.update(secondary_metrics(dev, peaks))
    barrier()
    if world > 1 and not args.no_extra:
        # frame-sharded back-mapping (configs[4]): no communication, every rank does its slice of 1M frames
        fps = backmap_frames_per_s(dev, BACKMAP_FRAMES // world)
        t = torch.tensor([fps[1]], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            extra["backmap_sharded"] = {"frames_per_s": BACKMAP_FRAMES // world * world / (t.item() * 1e-3), "n_gpus": world,
                                        "frames": BACKMAP_FRAMES // world * world, "scaling": "strong", "ms": t.item()}
        # third metric at N GPUs: data-parallel training of configs[1] (global batch 4096 fixed, and 4096 per GPU)
        try:
            sys.path.insert(0, str(ROOT / "tools"))
            import train_harness

            dp = {"strong_global_batch_4096": train_harness.run_dp(dev, dist.group.WORLD, steps=20, batch_global=4096, weak=False),
                  "weak_4096_per_gpu": train_harness.run_dp(dev, dist.group.WORLD, steps=20, batch_global=4096, weak=True)}
        except Exception as e:  # noqa: BLE001
            dp = {"error": repr(e)[:300]}
        if rank == 0:
            extra["train_steps_data_parallel"] = dp

    if rank == 0:
        lane_instr = pairs * (4 * N_DIMS + 60)
        peak = 148 * 128 * peaks["sm_max_mhz"] * 1e6
        achieved = (lane_instr / world) / (kernel_ms * 1e-3)
        # measured FP32 denominator: register-only FFMA chains on this device, right now (emk_probe_fp32)
        import ctypes

        probe = ctypes.c_double(0.0)
        _lib.check(L.emk_probe_fp32(ctypes.byref(probe)))
        # DRAM bytes of one launch: NOT measured in this run -- read from the committed ncu --set full capture of this kernel
        traffic, traffic_source = None, None
        for name in ("r02_pair_tile_traffic.json", "r01_pair_tile_traffic.json"):
            tp = ROOT / "profiles" / name
            if tp.exists():
                traffic = json.loads(tp.read_text()).get("dram_bytes_per_launch")
                traffic_source = f"committed ncu capture profiles/{name} (dram__bytes_read.sum + dram__bytes_write.sum of one launch at n_gpus=1), not a live measurement"
                break
        cpu_v, cpu_t, cores = (None, None, None)
        cpu_baseline = None
        if world == 1 and not args.no_cpu:
            cpu_v, cpu_t, cores = cpu_reference_value(5, 1)
            v256, t256, _ = cpu_reference_value(5, 1, 256)
            v1024, t1024, _ = cpu_reference_value(2, 1, 1024)
            cpu_baseline = {"value": cpu_v, "unit": "unique pairs/s", "cores": cores, "kind": "port",
                            "sample": f"N={CPU_SAMPLE_ROWS} x {N_DIMS} of the same data, float32 torch-CPU restatement (oracle/), median of 5, {cpu_t * 1e3:.0f} ms/eval",
                            "other_samples": {"N=256": {"value": v256, "ms_per_eval": t256 * 1e3}, "N=1024": {"value": v1024, "ms_per_eval": t1024 * 1e3}}}
        line = {
            "metric": "sigmoid_cost_pairs_per_s_fwd_bwd", "value": value, "unit": "unique pairs/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "configs[3]: full-set sketch-map sigmoid cost fwd+bwd, 65536 frames x 1024 periodic dims -> 2-d latent, upper-triangular 128x64 pair tiles sharded over GPUs, all-reduce of loss + dL/dz",
                       "n_rows": N_ROWS, "n_dims": N_DIMS, "latent": N_LATENT, "periodicity": "2pi", "sig": list(SIG),
                       "ordered_pairs_per_s": value * 2 * N_ROWS / (N_ROWS + 1), "l2": "inputs (268 MB) larger than L2 (126 MB); no flush needed",
                       "parallelism": f"tile-shard x{world}", "loss": loss_value},
            "roofline": {"bound": "fp32-issue", "achieved": achieved / 1e12, "peak": peak / 1e12, "unit": "T lane-instr/s",
                         "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_source,
                         "traffic_note": "30x the algorithmic 0.27 GB: column blocks are re-read from DRAM once per 1024-row band (L2 keeps the band); 0.4 % of DRAM bandwidth, irrelevant for a kernel bound by FP32 issue",
                         "peak_measured": probe.value / 1e12, "frac_of_measured": achieved / probe.value if probe.value else None,
                         "peak_measured_how": "emk_probe_fp32: register-only FFMA chains, 8 CTAs x 256 threads per SM, best of 3, CUDA events",
                         "kernel": "pair_tile_kernel<periodic,cost>", "kernel_ms": kernel_ms,
                         "algorithmic": "4 FP32 lane-instr per (unique pair, dim) + 60 per unique pair (SURVEY.md 8d); per GPU = total / n_gpus",
                         "peak_source": f"148 SM x 128 lanes x {peaks['sm_max_mhz']} MHz ({peaks['source']}); no FP32 figure is measured there"},
            "clocks": clocks.summary(),
            "e2e": {"value": e2e_value, "unit": "unique pairs/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": int(h2d_max), "d2h_bytes_per_step": gh.numel() * 4 + 4, "check_vs_device_path": e2e_check,
                    "note": "sigmoid_loss(...)(y_true pinned host tensor, y_pred) + backward; h2d_bytes_per_step = the largest rank's share. n_gpus <= 2: every rank copies the rows its tile range touches in chunks on a side stream behind the pair tiles that need them; n_gpus > 2: every rank takes 1/n_gpus of the tiles of every 8192-row chunk, copies 1/n_gpus of the chunk over its own host link and the chunk is completed by an all-gather over NVLink behind the tiles of the previous chunk; loss + gradient read back"},
            "gpu_launches": args.steps * world,
            "per_rank": per_rank,
            "multi_gpu_check": multi_gpu_check,
            "cpu_baseline": cpu_baseline,
            "secondary": secondary_summary(extra, world),
            "extra": extra,
        }
        print(json.dumps(line))
    if world > 1:
        parallel.destroy_comm()
        dist.destroy_process_group()


def secondary_summary(extra, world):
    """The other two metrics of BASELINE.json lifted out of `extra`: back-mapped frames/s (with its HBM roofline) and train
    steps/s.  Whole-job numbers at this n_gpus; the driver can compute scaling from the per-N lines."""
    out = {}
    bm = extra.get("backmap_sharded") if world > 1 else extra.get("backmap_fwd")
    if bm:
        out["backmap_frames_per_s"] = {"value": bm["frames_per_s"], "unit": "frames/s", "n_gpus": world, "n_atoms": BACKMAP_ATOMS,
                                       "frames": bm["frames"], "scaling": "strong (1M frames split by frame range, no communication)"}
        if "roofline" in bm:
            out["backmap_frames_per_s"]["roofline"] = bm["roofline"]
    ts = extra.get("train_steps")
    if isinstance(ts, dict) and "error" not in ts:
        out["train_steps_per_s"] = {k: {m: v[m]["steps_per_s"] for m in v} for k, v in ts.items()}
    dp = extra.get("train_steps_data_parallel")
    if isinstance(dp, dict) and "error" not in dp:
        out["train_steps_per_s_data_parallel"] = {k: {m: (v[m].get("steps_per_s") if isinstance(v[m], dict) else v[m]) for m in v} for k, v in dp.items()}
    return out


def backmap_frames_per_s(dev, frames):
    """Forward back-mapping of `frames` frames of a 500-residue chain in chunks; returns (frames/s, ms)."""
    import torch

    from encodermap_b200 import _ops

    n = BACKMAP_ATOMS
    chunk = min(BACKMAP_CHUNK, frames)
    g = torch.Generator(device=dev).manual_seed(555)
    lengths = (0.13 + 0.02 * torch.rand(1, n - 1, device=dev, generator=g)).contiguous()
    ang = (1.9 + 0.3 * torch.rand(chunk, n - 2, device=dev, generator=g)).contiguous()
    dih = ((torch.rand(chunk, n - 3, device=dev, generator=g) * 2 - 1) * math.pi).contiguous()
    n_chunks = max(1, frames // chunk)
    for _ in range(2):
        _ops.BackMap.apply(lengths, ang, dih)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n_chunks):
        _ops.BackMap.apply(lengths, ang, dih)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    return n_chunks * chunk / (ms * 1e-3), ms


def secondary_metrics(dev, peaks):
    import torch

    from encodermap_b200 import _ops

    out = {}
    # configs[4]: back-mapping, 500 residues
    fps, ms = backmap_frames_per_s(dev, BACKMAP_FRAMES)
    n = BACKMAP_ATOMS
    bytes_per_frame = 4 * ((n - 2) + (n - 3)) + 12 * n
    out["backmap_fwd"] = {"frames_per_s": fps, "frames": BACKMAP_FRAMES, "n_atoms": n, "ms": ms,
                          "roofline": {"bound": "hbm", "achieved": fps * bytes_per_frame / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                       "frac": fps * bytes_per_frame / 1e9 / peaks["hbm_gbs"], "bytes_per_frame": bytes_per_frame,
                                       "note": "float64 NeRF chain, one lane per (frame, side): bound by issue slots (FP64 instructions count twice), not by HBM (DESIGN.md 4.2); 262144-frame chunks"}}
    # fwd + bwd on one chunk: dihedral gradients only (the ADC default, use_backbone_angles=False:
    # reference parameters.py:803) and with bond-angle gradients as well
    chunk = 1 << 15
    g = torch.Generator(device=dev).manual_seed(556)
    lengths = (0.13 + 0.02 * torch.rand(1, n - 1, device=dev, generator=g)).contiguous()
    w = torch.randn(chunk, n, 3, device=dev, generator=g)
    for key, with_angles in (("backmap_fwd_bwd", False), ("backmap_fwd_bwd_angles", True)):
        ang = (1.9 + 0.3 * torch.rand(chunk, n - 2, device=dev, generator=g)).requires_grad_(with_angles)
        dih = ((torch.rand(chunk, n - 3, device=dev, generator=g) * 2 - 1) * math.pi).requires_grad_(True)
        n_warm, n_timed = 3, 5
        for it in range(n_warm + n_timed):
            if it == n_warm:
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            xyz = _ops.BackMap.apply(lengths, ang, dih)
            xyz.backward(w)
            ang.grad = dih.grad = None
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n_timed
        # forward bytes + backward: xyz and grad_xyz in, grad_dihedrals out (+ lengths, angles in and grad_angles out)
        bpf = bytes_per_frame + 24 * n + 4 * (n - 3) + (4 * (n - 1) + 8 * (n - 2) if with_angles else 0)
        out[key] = {"frames_per_s": chunk / (ms * 1e-3), "frames": chunk, "ms": ms, "bytes_per_frame": bpf,
                    "frac_hbm": chunk / (ms * 1e-3) * bpf / 1e9 / peaks["hbm_gbs"],
                    "gradients": "dihedrals + angles" if with_angles else "dihedrals (ADC default)"}
    # configs[1]: per-batch cost, 4096 x 1024
    x, z = synth_high(4096, 1024, dev, 1234)
    for it in range(6):
        if it == 1:
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        _ops.sigmoid_cost_raw(x, z, PERIOD, SIG)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    p2 = 4096 * 4097 / 2
    # third metric: train steps/s, the callers of the hot path driven by tools/train_harness.py (torch adapter;
    # fit()/summary/callback overhead of the reference excluded); eager and whole-step CUDA-graph replay
    try:
        sys.path.insert(0, str(ROOT / "tools"))
        import train_harness

        out["train_steps"] = train_harness.run_all(dev, steps=20)
    except Exception as e:  # noqa: BLE001 - a secondary metric must not take the headline down
        out["train_steps"] = {"error": repr(e)}
    out["cfg2_batch_cost"] = {"pairs_per_s": p2 / (ms * 1e-3), "ms": ms, "n_rows": 4096, "n_dims": 1024,
                              "frac_fp32_issue": p2 * (4 * 1024 + 60) / (ms * 1e-3) / (148 * 128 * peaks["sm_max_mhz"] * 1e6)}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extra", action="store_true", help="skip the secondary metrics (profiling runs)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
