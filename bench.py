#!/usr/bin/env python
"""Headline benchmark: (query, product) pairs scored per second (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--model imagebert_zk|imagebert_lds|lxmert] [--impl reference]

A step is one pass of the scoring hot path over one batch of 256 synthetic pairs at the BASELINE configs[1]
shapes (12-layer ImageBert, 32 query tokens x 36 regions x 2048-d; configs[2] with --model lxmert).  For N > 1 the
driver launches this file under torchrun; every rank scores its own 256-pair batches (the pair list shards with no
data-path collective) and the per-rank fp32 scores are concatenated by ONE NCCL all-gather inside the timed region.

Printed JSON (rank 0, one line): value = whole-job pairs/s with inputs resident in HBM; e2e = the same metric through
MatchScorer.score with pinned HOST feeds (H2D of every step's inputs and D2H of its scores inside the timed region);
roofline = the tcgen05 GEMM kernel's achieved TFLOP/s over all its launches inside profiled steps, against the
measured cuBLAS bf16 peak; cpu_baseline = the fp32 oracle port of the same model timed on the host cores.
--impl reference times that CPU port alone (TF-1.12 / Python 2 cannot run).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from kddcup_2020_multimodalitiesrecall_2nd_place_b200 import synth  # noqa: E402
from kddcup_2020_multimodalitiesrecall_2nd_place_b200.config import (LDS, LXMERT, ZK, baseline_cfg2,  # noqa: E402
                                                                      flops_per_pair)

BATCH = 256
METRIC = "pairs_scored_per_sec"
UNIT = "pairs/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--model", default=ZK, choices=[ZK, LDS, LXMERT])
    ap.add_argument("--dtype", default="fp16", choices=["fp16", "bf16"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def workload_name(cfg, batch):
    if cfg.kind == LXMERT:
        return (f"LXMERT dual-stream ({cfg.n_layers} lang / {cfg.n_r_layers} vis / {cfg.n_x_layers} cross layers), "
                f"{cfg.lq} query tokens x {cfg.nbox} regions x {cfg.feat_dim}-d, batch={batch}")
    return (f"{cfg.n_layers}-layer ImageBert ({cfg.kind}), {cfg.lq} query tokens x {cfg.nbox} regions x "
            f"{cfg.feat_dim}-d, batch={batch}")


# ---------------------------------------------------------------------------------------------- clocks
class NvmlThreadSampler:
    """Fallback: samples SM clock / throttle reasons of one GPU through NVML from a thread of THIS process."""

    def __init__(self, index):
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            self.nv = pynvml
            pynvml.nvmlInit()
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:  # pragma: no cover
            self.nv = None
            self.err = repr(e)

    def _run(self):
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4),
            "hw_power_brake": getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80),
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.02)

    def start(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()

    def stop(self):
        self._stop.set()
        if self._thr is not None:
            self._thr.join()
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples), "source": "nvml thread"}


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region.  Primary source: an NVML poll (every 20 ms) from a thread
    of this process — an A/B at N = 1 measured no cost (69.1 vs 69.2 k pairs/s against an out-of-process sampler).
    Fall-back when pynvml is not usable: a separate `nvidia-smi -lms 50` process (the profiling recipe's clocks line),
    started before the barrier, of which only the samples stamped inside [start(), stop()] are used (it delivered a
    single sample inside a 360 ms region at N = 2, which is why it is not the primary)."""

    FIELDS = ("timestamp,clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")
    REASONS = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")

    def __init__(self, index):
        import shutil
        import subprocess
        import tempfile
        self.proc, self.fallback = None, None
        self.t0 = self.t1 = None
        exe = shutil.which("nvidia-smi")
        mode = os.environ.get("MMR_BENCH_SAMPLER", "nvml")
        if mode == "nvml":
            nv = NvmlThreadSampler(index)
            if nv.nv is not None or exe is None:
                self.fallback = nv
                return
        if exe is not None:
            self.log = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
            try:
                self.proc = subprocess.Popen([exe, "-i", str(index), f"--query-gpu={self.FIELDS}",
                                              "--format=csv,noheader,nounits", "-lms", "50"],
                                             stdout=self.log, stderr=subprocess.DEVNULL)
            except Exception:  # pragma: no cover
                self.proc = None
            if self.proc is not None:
                import atexit
                atexit.register(lambda p=self.proc: p.poll() is None and p.kill())
        if self.proc is None:
            self.fallback = NvmlThreadSampler(index)

    def start(self):
        self.t0 = time.time()
        if self.fallback is not None:
            self.fallback.start()

    def _parse(self):
        import datetime
        mhz, reasons, max_mhz = [], set(), None
        self.log.flush()
        with open(self.log.name) as f:
            for line in f:
                c = [x.strip() for x in line.split(",")]
                if len(c) != 7:
                    continue
                try:
                    ts = datetime.datetime.strptime(c[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                    clk = float(c[1])
                    max_mhz = float(c[2])
                except ValueError:
                    continue
                if self.t0 - 0.025 <= ts <= self.t1 + 0.025:
                    mhz.append(clk)
                    for name, v in zip(self.REASONS, c[3:]):
                        if v == "Active":
                            reasons.add(name)
        return mhz, reasons, max_mhz

    def stop(self):
        self.t1 = time.time()
        if self.fallback is not None:
            return self.fallback.stop()
        time.sleep(0.06)   # let the sample that covers the end of the region land in the file
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:  # pragma: no cover
            self.proc.kill()
        mhz, reasons, max_mhz = self._parse()
        try:
            os.unlink(self.log.name)
        except OSError:
            pass
        if not mhz:
            return {"sm_mhz": None, "sm_max_mhz": max_mhz, "reasons": [], "samples": 0, "source": "nvidia-smi (no samples)"}
        return {"sm_mhz": float(np.median(mhz)), "sm_max_mhz": int(max_mhz), "reasons": sorted(reasons),
                "samples": len(mhz), "source": "nvidia-smi -lms 50 subprocess"}


# ---------------------------------------------------------------------------------------------- CPU arm
def oracle_forward_fn(cfg):
    """fp32 CPU port of the same model (oracle/ is the checker; here it is only TIMED, never shipped)."""
    from oracle import imagebert, lxmert
    if cfg.kind == ZK:
        return lambda w, i: imagebert.zk_forward(w, i, cfg.n_layers)["probs"]
    if cfg.kind == LDS:
        return lambda w, i: imagebert.lds_forward(w, i, cfg.n_layers)["probs"]
    return lambda w, i: lxmert.forward(w, i, cfg.n_layers, cfg.n_r_layers, cfg.n_x_layers)["probs"]


def time_cpu_port(cfg, weights, sample_pairs, steps, warmup):
    from oracle import imagebert
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    fwd = oracle_forward_fn(cfg)
    w = imagebert.to_torch(weights)
    inp = imagebert.to_torch(synth.make_inputs(cfg, sample_pairs, seed=synth.SEED0 + 99))
    for _ in range(warmup):
        fwd(w, inp)
    t0 = time.perf_counter()
    for _ in range(steps):
        fwd(w, inp)
    dt = time.perf_counter() - t0
    return sample_pairs * steps / dt, dt / steps, cores


def run_reference(args, cfg, rank):
    if rank != 0:
        return
    sample = 16
    weights = synth.make_weights(cfg, seed=synth.SEED0)
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    # keep the whole run within a few minutes: probe one step, then cap the step count
    pps, s_per_step, cores = time_cpu_port(cfg, weights, sample, 1, 1)
    budget_steps = max(1, int(150.0 / max(s_per_step, 1e-3)))
    steps_run = min(steps, budget_steps)
    pps, s_per_step, cores = time_cpu_port(cfg, weights, sample, steps_run, min(warmup, 2))
    line = {
        "impl": "reference", "metric": METRIC, "value": pps, "unit": UNIT, "n_gpus": args.gpus, "steps": steps_run,
        "warmup": min(warmup, 2), "ms_per_step": s_per_step * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(cfg, args.batch), "sample_pairs_per_step": sample},
        "cpu_baseline": {"value": pps, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{sample} pairs per step of the same workload, fp32 PyTorch restatement of the "
                                   f"reference graph (TF-1.12/py2 not runnable), {cores} threads"},
        "e2e": {"value": pps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------- GPU arm
def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cfg = baseline_cfg2(args.model)
    if args.impl == "reference":
        run_reference(args, cfg, rank)
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference for "
                         "the CPU arm)")
    import torch.distributed as dist
    from kddcup_2020_multimodalitiesrecall_2nd_place_b200.scorer import MatchScorer
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, K, W = args.batch, args.steps, max(args.warmup, 3)

    weights = synth.make_weights(cfg, seed=synth.SEED0)
    sc = MatchScorer(cfg, weights, device=local, dtype=args.dtype, max_batch=B)
    # rotating resident input sets: 3 x 75.5 MB of fp32 features + ~220 MB of weights per step > the 126 MB L2
    n_sets = 3
    host_sets = [sc.to_feeds(synth.make_inputs(cfg, B, seed=synth.SEED0 + 1000 * rank + s, n_queries=max(1, B // 30)))
                 for s in range(n_sets)]
    dev_sets = [{k: v.to(dev) for k, v in hs.items()} for hs in host_sets]
    scores = torch.empty((K, B, 2), dtype=torch.float32, device=dev)
    gathered = torch.empty((world, K * B), dtype=torch.float32, device=dev) if world > 1 else None

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # every input set scores into its own buffer, so that a forward sees the same pointers every n_sets steps: the
    # scorer captures it into a CUDA graph the second time and replays it from then on.  Enough untimed warm-up steps
    # for every set to have been captured before the timed region starts.
    outs = [torch.empty((B, 2), dtype=torch.float32, device=dev) for _ in range(n_sets)]
    for i in range(2 * n_sets):                 # set-up, not warm-up: first pass eager, second pass = the captures
        sc.forward_device(dev_sets[i % n_sets], probs_out=outs[i % n_sets])
    for i in range(W):
        sc.forward_device(dev_sets[i % n_sets], probs_out=outs[i % n_sets])
    if world > 1:
        dist.all_gather_into_tensor(gathered.view(-1), scores[:, :, 1].contiguous().view(-1))
    # NVML initialisation takes ~15 ms: before the barrier, or rank 0 would enter the timed region that much after
    # the other ranks and they would wait for it in the all-gather
    sampler = ClockSampler(local) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    if sampler:
        sampler.start()
    e0.record()
    for k in range(K):
        sc.forward_device(dev_sets[k % n_sets], probs_out=outs[k % n_sets])
        scores[k].copy_(outs[k % n_sets], non_blocking=True)      # every step's scores are kept (and gathered below)
    e_fw = torch.cuda.Event(enable_timing=True)
    e_fw.record()
    if world > 1:
        dist.all_gather_into_tensor(gathered.view(-1), scores[:, :, 1].contiguous().view(-1))
    e1.record()
    barrier()
    clocks = sampler.stop() if sampler else None
    if os.environ.get("MMR_BENCH_DEBUG"):
        print(f"DEBUG rank {rank}: forwards {e0.elapsed_time(e_fw):.2f} ms, gather {e_fw.elapsed_time(e1):.2f} ms",
              file=sys.stderr, flush=True)
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    total_ms = float(ms.item())
    launches = sc.launches_per_forward() * K
    value = world * K * B / (total_ms * 1e-3)
    if not torch.isfinite(scores).all():
        raise SystemExit("bench.py: non-finite scores")

    # ---- end to end: pinned host feeds -> H2D -> kernels -> D2H scores, every step, through MatchScorer.score
    e2e = None
    if not args.no_e2e:
        Ke = min(K, 24)
        big = {k: torch.cat([host_sets[s % n_sets][k] for s in range(Ke)]).pin_memory() for k in host_sets[0]}
        out_host = torch.empty((Ke * B, 2), dtype=torch.float32).pin_memory()
        sc.score({k: v[: 6 * B] for k, v in big.items()})  # warm the copy stream / slots (and their two CUDA graphs)
        barrier()
        t0 = time.perf_counter()
        e0.record()
        sc.score(big, out=out_host)
        e1.record()
        barrier()
        wall = time.perf_counter() - t0
        ms2 = torch.tensor([max(e0.elapsed_time(e1), wall * 1e3)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
        h2d = sum(v[:B].numel() * v.element_size() for v in big.values())
        e2e = {"value": world * Ke * B / (float(ms2.item()) * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": B * 2 * 4, "steps": Ke}

    # ---- roofline: per-launch CUDA events (recorded by the library on the forward's stream) inside profiled steps.
    # Launches are classified by their algorithmic FLOPs; the dominant kernel is the 16-bit-output tcgen05 GEMM
    # (QKV + FFN-in projections), reported against the measured sustained cuBLAS peak.
    roofline = None
    if rank == 0:
        M = B * cfg.seq_len
        H, I = cfg.hidden, cfg.intermediate
        # FFN-in and FFN-out have the same FLOPs: they alternate, FFN-in first (launch order inside a layer)
        classes = {2.0 * M * 3 * H * H: "qkv", 2.0 * M * H * H: "out_proj_ln"} if cfg.kind != LXMERT else {}
        ffn_flops = 2.0 * M * H * I if cfg.kind != LXMERT else -1.0
        sc.set_profiling(True)
        agg, per = {}, {}
        n_prof = 5
        for k in range(n_prof):
            sc.forward_device(dev_sets[k % n_sets], probs_out=scores[k % K])
            ffn_toggle = 0
            for kind, t, fl in sc.profile():
                a = agg.setdefault(kind, [0.0, 0.0, 0])
                a[0] += t
                a[1] += fl
                a[2] += 1
                if kind == 0 and fl == ffn_flops:
                    name = ("ffn_in", "ffn_out_ln")[ffn_toggle]
                    ffn_toggle ^= 1
                else:
                    name = classes.get(fl, "other_gemm") if kind == 0 else {1: "attention", 2: "layernorm"}.get(kind, "rows")
                q = per.setdefault(name, [0.0, 0.0, 0])
                q[0] += t
                q[1] += fl
                q[2] += 1
        sc.set_profiling(False)
        peaks, ncu = {}, {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        try:
            ncu = json.load(open(os.path.join(ROOT, "profiles", "r01m_gemm_ncu_summary.json")))["kernels"]
        except Exception:
            pass
        peak = peaks.get("bf16_tflops_sustained") or 1400.0
        peak_src = ("MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)"
                    if peaks.get("bf16_tflops_sustained") else "fallback 1400 TFLOP/s sustained (B200_PROFILING.md)")
        step_ms = sum(a[0] for a in agg.values())
        kernels = {}
        for name, (t, fl, n) in sorted(per.items()):
            kernels[name] = {"launches_per_step": n // n_prof, "avg_launch_us": t / n * 1e3, "share_of_step": t / step_ms,
                             "tflops": (fl / (t * 1e-3) / 1e12) if fl else None,
                             "frac_of_peak": (fl / (t * 1e-3) / 1e12 / peak) if fl else None}
        dom = [per[k] for k in ("qkv", "ffn_in") if k in per] or [agg.get(0, [1e-9, 0.0, 1])]
        d_ms, d_fl, d_n = (sum(x[i] for x in dom) for i in range(3))
        achieved = d_fl / (d_ms * 1e-3) / 1e12
        g_ms, g_fl, g_n = agg.get(0, [1e-9, 0.0, 1])
        traffic = None
        if "qkv" in ncu and "ffn_in" in ncu:   # DRAM bytes per launch of the dominant kernel, one ncu --set full capture
            traffic = 0.5e6 * sum(ncu[k]["dram_read_MB"] + ncu[k]["dram_write_MB"] for k in ("qkv", "ffn_in"))
        roofline = {
            "bound": "tensor", "kernel": "gemm_pair16_kernel (QKV and FFN-in projections, 16-bit output, TMA-store epilogue)",
            "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic,
            "traffic_source": "profiles/r01m_gemm_ncu_summary.json (dram read+write, mean of the two shapes)" if traffic else None,
            "peak_source": peak_src, "launches_per_step": d_n // n_prof, "avg_launch_us": d_ms / d_n * 1e3,
            "flops_per_launch": d_fl / d_n,
            "timing": "CUDA events recorded by the library after every launch on the forward's stream, 5 profiled steps "
                      "(event records between launches suppress the programmatic-dependent-launch overlap: per-kernel "
                      "times are upper bounds; `value` is measured without them)",
            "all_gemm": {"tflops": g_fl / (g_ms * 1e-3) / 1e12, "frac_of_peak": g_fl / (g_ms * 1e-3) / 1e12 / peak,
                         "launches_per_step": g_n // n_prof},
            "kernels": kernels,
            "whole_step": {"algorithmic_tflops": flops_per_pair(cfg) * value / world / 1e12,
                           "frac_of_peak": flops_per_pair(cfg) * value / world / 1e12 / peak,
                           "frac_of_burst_peak": (flops_per_pair(cfg) * value / world / 1e12 / peaks["bf16_tflops"])
                           if peaks.get("bf16_tflops") else None},
        }

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sample = 32
        pps, s_per, cores = time_cpu_port(cfg, weights, sample, 1, 1)
        cpu = {"value": pps, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{sample} pairs of the same workload, one timed pass after one warm-up pass ({s_per:.1f} s), "
                         f"fp32 PyTorch restatement of the reference graph on {cores} host threads"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": workload_name(cfg, B), "pairs_per_step_per_gpu": B,
                       "l2_policy": f"{n_sets} rotating resident input sets (3 x 75.5 MB fp32 features) + 220 MB of "
                                    "weights streamed per step: working set larger than the 126 MB L2",
                       "flops_per_pair": flops_per_pair(cfg),
                       "arithmetic": f"{args.dtype} MMA operands, fp32 accumulate / residual stream / LayerNorm / softmax",
                       "collective": "one NCCL all-gather of fp32 scores inside the timed region" if world > 1 else None,
                       "launch": ("forward replayed from CUDA graphs (one per rotating input set, captured before the "
                                  "warm-up steps)" if sc.use_graphs else "69 eager launches per forward")},
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
