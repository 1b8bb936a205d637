#!/usr/bin/env python
"""Headline benchmark: (query, product) pairs scored per second (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--model imagebert_zk|imagebert_lds|lxmert] [--impl reference]
                    [--precision fast|strict] [--quick]

A step is one pass of the scoring hot path over one batch of 256 synthetic pairs at the BASELINE configs[1]
shapes (12-layer ImageBert, 32 query tokens x 36 regions x 2048-d; configs[2] with --model lxmert).  For N > 1 the
driver launches this file under torchrun; every rank scores its own 256-pair batches (the pair list shards with no
data-path collective) and the per-rank fp32 scores are concatenated by ONE NCCL all-gather inside the timed region.

Printed JSON (rank 0, one line):
  value         whole-job pairs/s with inputs resident in HBM (weak scaling: 256 pairs per rank per step)
  e2e           the same metric through MatchScorer.score with pinned HOST feeds (H2D + D2H inside the timed region)
  roofline      the tcgen05 GEMM kernel's achieved TFLOP/s against the measured cuBLAS bf16 peak -- the BURST figure when
                the timed region is shorter than 1 s at the maximum SM clock, else the sustained one; per-kernel times
                are shares of an event-profiled step applied to the step measured without events (they sum to it);
                roofline.sustained = >= 2.5 s of back-to-back replays against the sustained peak, with its own clocks
  cpu_baseline  the fp32 oracle port of the same model timed on the host cores
  other_models  configs[1] with imagebert_lds and configs[2] (LXMERT 9/5/5), same protocol (N = 1)
  strict        the strict-precision mode's throughput against the default (N = 1)
  cfg4 / cfg5   BASELINE configs[3] / [4]: the 30,000-pair candidate set and the 3-model ensemble over 10,000 pairs,
                from HOST feeds, sharded over the ranks (STRONG scaling: pairs, wall_ms, pairs_per_s, speedup_vs_n1)
--impl reference times the CPU port alone (TF-1.12 / Python 2 cannot run).
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
N_SETS = 3   # rotating resident input sets: 3 x 75.5 MB of fp32 features + ~220 MB of weights per step > the 126 MB L2
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
    ap.add_argument("--precision", default="fast", choices=["fast", "strict"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--quick", action="store_true",
                    help="headline + e2e + roofline only: no sustained run, other models, strict mode, cfg4 / cfg5")
    return ap.parse_args()


def workload_name(cfg, batch):
    if cfg.kind == LXMERT:
        return (f"LXMERT dual-stream ({cfg.n_layers} lang / {cfg.n_r_layers} vis / {cfg.n_x_layers} cross layers), "
                f"{cfg.lq} query tokens x {cfg.nbox} regions x {cfg.feat_dim}-d, batch={batch}")
    return (f"{cfg.n_layers}-layer ImageBert ({cfg.kind}), {cfg.lq} query tokens x {cfg.nbox} regions x "
            f"{cfg.feat_dim}-d, batch={batch}")


def config_record(cfg, B, args, world, graphs=True):
    """The `config` object of the JSON line: names the workload; identical for the GPU arm and the --impl reference arm."""
    return {"workload": workload_name(cfg, B), "pairs_per_step_per_gpu": B,
            "l2_policy": f"{N_SETS} rotating resident input sets (3 x 75.5 MB fp32 features) + 220 MB of "
                         "weights streamed per step: working set larger than the 126 MB L2",
            "flops_per_pair": flops_per_pair(cfg), "precision": args.precision,
            "arithmetic": f"{args.dtype} MMA operands, fp32 accumulate / residual stream / LayerNorm / softmax",
            "collective": "one NCCL all-gather of fp32 scores inside the timed region" if world > 1 else None,
            "last_block": "keys / values for all rows, attention + projections + FFN for the [CLS] rows only "
                          "(the poolers read sequence_output[:, 0]); FLOPs counted are the reference's",
            "launch": ("forward replayed from CUDA graphs (one per rotating input set, captured before the "
                       "warm-up steps)" if graphs else "eager launches"),
            "query_grouping": (f"LXMERT: the {cfg.n_layers} query-only language blocks run once per DISTINCT query of a "
                               f"batch ({max(1, B // 30)} queries per {B} pairs, the testB ratio of ~30 candidates per "
                               "query) and are expanded before the cross-modality blocks; FLOPs counted are the "
                               "reference's, which evaluates them per pair") if cfg.kind == LXMERT else None}


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
    """The reference arm: the reference's own algorithm for this path on the host cores (the fp32 port: TF-1.12 / py2
    cannot run), K steps after W warm-up steps, each step a bounded SAMPLE of the GPU arm's step (16 of its 256 pairs),
    same metric / unit / config."""
    if rank != 0:
        return
    sample = 16
    world = int(os.environ.get("WORLD_SIZE", "1"))
    weights = synth.make_weights(cfg, seed=synth.SEED0)
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    # keep the whole run within a few minutes: probe one step, then cap the step count
    pps, s_per_step, cores = time_cpu_port(cfg, weights, sample, 1, 1)
    budget_steps = max(1, int(150.0 / max(s_per_step, 1e-3)))
    steps_run = min(steps, budget_steps)
    warm_run = min(warmup, max(0, budget_steps // 4))
    pps, s_per_step, cores = time_cpu_port(cfg, weights, sample, steps_run, warm_run)
    line = {
        "impl": "reference", "metric": METRIC, "value": pps, "unit": UNIT, "n_gpus": args.gpus, "steps": steps_run,
        "warmup": warm_run, "ms_per_step": s_per_step * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_record(cfg, args.batch, args, world),
        "cpu_baseline": {"value": pps, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{sample} pairs per step of the same workload ({args.batch} pairs per step on the GPU "
                                   f"arm), fp32 PyTorch restatement of the reference graph (TF-1.12 / py2 not runnable), "
                                   f"{cores} threads"},
        "e2e": {"value": pps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------- GPU arm
def load_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def pick_peak(peaks, clocks, region_s):
    """Denominator of a tensor-bound fraction: the BURST cuBLAS figure for a region that ran (nearly) at the maximum SM
    clock and lasted under a second, the SUSTAINED one for a long or power-capped region."""
    burst, sus = peaks.get("bf16_tflops"), peaks.get("bf16_tflops_sustained")
    at_max = bool(clocks and clocks.get("sm_mhz") and clocks.get("sm_max_mhz")
                  and clocks["sm_mhz"] >= 0.97 * clocks["sm_max_mhz"])
    if region_s < 1.0 and at_max and burst:
        return burst, "MEASURED_PEAKS.json bf16_tflops (burst: timed region < 1 s at the maximum SM clock)"
    if sus:
        return sus, "MEASURED_PEAKS.json bf16_tflops_sustained (region >= 1 s or SM clock below maximum under load)"
    return 1400.0, "fallback 1400 TFLOP/s sustained (B200_PROFILING.md)"


class Env:
    def __init__(self, args):
        import torch.distributed as dist
        self.dist = dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)

    def barrier(self):
        torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
            torch.cuda.synchronize()

    def max_ms(self, ms):
        t = torch.tensor([ms], dtype=torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())



def resident_sets(sc, cfg, B, rank, dev):
    host = [sc.to_feeds(synth.make_inputs(cfg, B, seed=synth.SEED0 + 1000 * rank + s, n_queries=max(1, B // 30)))
            for s in range(N_SETS)]
    return host, [{k: v.to(dev) for k, v in hs.items()} for hs in host]


def time_resident(env, sc, dev_sets, B, K, W, sample_clocks=False, gather=True):
    """K forwards over rotating device-resident input sets, device-timed, every step's scores kept; at N > 1 one NCCL
    all-gather of them inside the timed region.  Returns (total ms = max over ranks, clocks, scores)."""
    dev, world = env.dev, env.world
    scores = torch.empty((K, B, 2), dtype=torch.float32, device=dev)
    gathered = torch.empty((world, K * B), dtype=torch.float32, device=dev) if (world > 1 and gather) else None
    # every input set scores into its own buffer, so that a forward sees the same pointers every N_SETS steps: the
    # scorer captures it into a CUDA graph the second time and replays it from then on
    outs = [torch.empty((B, 2), dtype=torch.float32, device=dev) for _ in range(N_SETS)]
    for i in range(2 * N_SETS):                 # set-up, not warm-up: first pass eager, second pass = the captures
        sc.forward_device(dev_sets[i % N_SETS], probs_out=outs[i % N_SETS])
    for i in range(W):
        sc.forward_device(dev_sets[i % N_SETS], probs_out=outs[i % N_SETS])
    if gathered is not None:
        env.dist.all_gather_into_tensor(gathered.view(-1), scores[:, :, 1].contiguous().view(-1))
    # NVML initialisation takes ~15 ms: before the barrier, or rank 0 would enter the timed region that much after
    # the other ranks and they would wait for it in the all-gather
    sampler = ClockSampler(env.local) if (sample_clocks and env.rank == 0) else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    env.barrier()
    if sampler:
        sampler.start()
    e0.record()
    for k in range(K):
        sc.forward_device(dev_sets[k % N_SETS], probs_out=outs[k % N_SETS])
        scores[k].copy_(outs[k % N_SETS], non_blocking=True)      # every step's scores are kept (and gathered below)
    if gathered is not None:
        env.dist.all_gather_into_tensor(gathered.view(-1), scores[:, :, 1].contiguous().view(-1))
    e1.record()
    env.barrier()
    clocks = sampler.stop() if sampler else None
    total_ms = env.max_ms(e0.elapsed_time(e1))
    if not torch.isfinite(scores).all():
        raise SystemExit("bench.py: non-finite scores")
    return total_ms, clocks, scores


def time_e2e(env, sc, host_sets, B, K):
    """Pinned host feeds -> H2D -> kernels -> D2H scores, every step, through MatchScorer.score."""
    Ke = min(K, 24)
    big = {k: torch.cat([host_sets[s % N_SETS][k] for s in range(Ke)]).pin_memory() for k in host_sets[0]}
    out_host = torch.empty((Ke * B, 2), dtype=torch.float32).pin_memory()
    sc.score({k: v[: 6 * B] for k, v in big.items()})  # warm the copy stream / slots (and their two CUDA graphs)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    env.barrier()
    t0 = time.perf_counter()
    e0.record()
    sc.score(big, out=out_host)
    e1.record()
    env.barrier()
    wall = time.perf_counter() - t0
    ms = env.max_ms(max(e0.elapsed_time(e1), wall * 1e3))
    h2d = sum(v[:B].numel() * v.element_size() for v in big.values())
    return {"value": env.world * Ke * B / (ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
            "d2h_bytes_per_step": B * 2 * 4, "steps": Ke}


def sustained_record(env, sc, cfg, dev_sets, B, peaks, seconds=2.5):
    """The sustained regime: >= `seconds` of back-to-back graph replays after the headline region, own clock / power
    samples, against the SUSTAINED cuBLAS peak (the headline region is a burst of a few dozen ms)."""
    outs = [torch.empty((B, 2), dtype=torch.float32, device=env.dev) for _ in range(N_SETS)]
    for i in range(2 * N_SETS):
        sc.forward_device(dev_sets[i % N_SETS], probs_out=outs[i % N_SETS])
    torch.cuda.synchronize()
    sampler = ClockSampler(env.local)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.start()
    t_end = time.perf_counter() + seconds
    steps = 0
    e0.record()
    while time.perf_counter() < t_end:
        for _ in range(30):
            sc.forward_device(dev_sets[steps % N_SETS], probs_out=outs[steps % N_SETS])
            steps += 1
        torch.cuda.synchronize()      # keeps the host at most 30 steps ahead, so the loop ends on time
    e1.record()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1)
    pps = steps * B / (ms * 1e-3)
    tf = flops_per_pair(cfg) * pps / 1e12
    sus = peaks.get("bf16_tflops_sustained") or 1400.0
    return {"seconds": ms * 1e-3, "steps": steps, "value_per_gpu": pps, "ms_per_step": ms / steps,
            "algorithmic_tflops": tf, "peak": sus, "frac_of_sustained_peak": tf / sus, "clocks": clocks}


def roofline_record(sc, cfg, dev_sets, scores, B, value_per_gpu, ms_per_step, clocks, region_s, peaks):
    """Per-launch CUDA events (recorded by the library on the forward's stream) inside profiled steps, classified by
    their algorithmic FLOPs.  Event records between launches suppress the programmatic-dependent-launch overlap, so
    the RAW event times of a step add up to more than the step measured without them; what is reported per kernel is
    its SHARE of the profiled step applied to the measured step (sum over kernels == ms_per_step), which is also what
    the ncu launch list under profiles/ is compared on."""
    M = B * cfg.seq_len
    H, I = cfg.hidden, cfg.intermediate
    classes = {2.0 * M * 3 * H * H: "qkv", 2.0 * M * H * H: "out_proj_ln"} if cfg.kind != LXMERT else {}
    ffn_flops = 2.0 * M * H * I if cfg.kind != LXMERT else -1.0
    sc.set_profiling(True)
    agg, per = {}, {}
    n_prof = 5
    for k in range(n_prof):
        sc.forward_device(dev_sets[k % N_SETS], probs_out=scores[k % scores.shape[0]])
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
    ncu = {}
    for name in ("r02_gemm_ncu_summary.json", "r01m_gemm_ncu_summary.json"):
        try:
            ncu = json.load(open(os.path.join(ROOT, "profiles", name)))["kernels"]
            ncu_src = f"profiles/{name}"
            break
        except Exception:
            pass
    peak, peak_src = pick_peak(peaks, clocks, region_s)
    raw_step_ms = sum(a[0] for a in agg.values()) / n_prof
    scale = ms_per_step / raw_step_ms          # < 1: the overlap the event records suppress
    kernels = {}
    for name, (t, fl, n) in sorted(per.items()):
        us = t / n * 1e3 * scale
        kernels[name] = {"launches_per_step": n // n_prof, "avg_launch_us": us, "raw_event_us": t / n * 1e3,
                         "share_of_step": (t / n_prof) / raw_step_ms,
                         "tflops": (fl / n / (us * 1e-6) / 1e12) if fl else None,
                         "frac_of_peak": (fl / n / (us * 1e-6) / 1e12 / peak) if fl else None}
    dom = [per[k] for k in ("qkv", "ffn_in") if k in per] or [agg.get(0, [1e-9, 0.0, 1])]
    d_ms, d_fl, d_n = (sum(x[i] for x in dom) for i in range(3))
    d_us = d_ms / d_n * 1e3 * scale
    achieved = d_fl / d_n / (d_us * 1e-6) / 1e12
    traffic = None
    if "qkv" in ncu and "ffn_in" in ncu:   # DRAM bytes per launch of the dominant kernel, one ncu --set full capture
        traffic = 0.5e6 * sum(ncu[k]["dram_read_MB"] + ncu[k]["dram_write_MB"] for k in ("qkv", "ffn_in"))
    whole_tf = flops_per_pair(cfg) * value_per_gpu / 1e12
    return {
        "bound": "tensor", "kernel": "gemm_pair16_kernel (QKV and FFN-in projections, 16-bit output, TMA-store epilogue)",
        "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic,
        "traffic_source": f"{ncu_src} (dram read+write, mean of the two shapes)" if traffic else None,
        "peak_source": peak_src, "launches_per_step": d_n // n_prof, "avg_launch_us": d_us,
        "flops_per_launch": d_fl / d_n,
        "timing": "share of the event-timed profiled step (5 steps, one CUDA event after every launch on the forward's "
                  "stream) x the step time measured without events; raw event times are kept as raw_event_us: they "
                  f"sum to {raw_step_ms:.3f} ms per step against {ms_per_step:.3f} ms measured",
        "kernels": kernels,
        "whole_step": {"algorithmic_tflops": whole_tf, "frac_of_peak": whole_tf / peak,
                       "frac_of_burst_peak": whole_tf / peaks["bf16_tflops"] if peaks.get("bf16_tflops") else None,
                       "frac_of_sustained_peak": whole_tf / peaks["bf16_tflops_sustained"]
                       if peaks.get("bf16_tflops_sustained") else None},
    }


# ---------------------------------------------------------------------------------------------- cfg4 / cfg5 workloads
CFG4_PAIRS, CFG4_QUERIES = 30000, 1000      # BASELINE configs[3]: 1k queries x 30 candidates, 12-layer, sharded
CFG5_PAIRS, CFG5_QUERIES = 10000, 400       # BASELINE configs[4]: 3-model ensemble over 10k pairs (25 candidates each)
BLOCK = 1536                                # distinct synthetic pairs behind a candidate set (6 chunks of 256)


def candidate_block(sc, cfg, seed):
    """Pinned host feeds of BLOCK (+ one chunk of wrap-around) synthetic pairs; pair i of a candidate set reads block
    row i % BLOCK, so a rank stages exactly its own pairs' bytes per chunk without the set being resident as one
    8.8 GB array (the same H2D traffic: 295 KB per pair)."""
    host = sc.to_feeds(synth.make_inputs(cfg, BLOCK, seed=seed, n_queries=BLOCK // 30))
    return {k: torch.cat([v, v[:256]]).pin_memory() for k, v in host.items()}


def block_fetch(block):
    def fetch(lo, hi):
        p = lo % BLOCK
        return {k: v[p:p + (hi - lo)] for k, v in block.items()}
    return fetch


def _n1_cache(name):
    import tempfile
    return os.path.join(tempfile.gettempdir(), f"mmr_bench_{name}_n1.json")


def _speedup_vs_n1(name, world, pps):
    """Strong scaling against the N = 1 run of the same workload on this box (the driver runs N = 1, 2, 4, 8 back to
    back); null when that run is not on record."""
    path = _n1_cache(name)
    if world == 1:
        try:
            json.dump({"pairs_per_s": pps, "when": time.time()}, open(path, "w"))
        except OSError:
            pass
        return 1.0
    try:
        rec = json.load(open(path))
        if time.time() - rec["when"] < 6 * 3600:
            return pps / rec["pairs_per_s"]
    except Exception:
        pass
    return None


def run_cfg4(env, sc, cfg):
    """30,000 candidate pairs (1,000 queries x 30), sharded contiguously over the ranks, every rank staging its own
    pairs from pinned host memory in 256-pair chunks (ragged last chunk), ONE all-gather of the fp32 scores, top-5 per
    query on every rank.  Strong scaling: the pair list is fixed, the ranks split it."""
    from kddcup_2020_multimodalitiesrecall_2nd_place_b200.scorer import sharded_score_stream
    block = candidate_block(sc, cfg, synth.SEED0 + 404)
    fetch = block_fetch(block)
    # warm-up = the workload itself, twice: MatchScorer runs a (batch, buffers) combination eagerly the first time it
    # sees it and captures it into a CUDA graph the second time; the timed pass then replays, as any pass of a long
    # scoring job after its first two does (a capture costs ~15 ms, a ragged tail chunk per rank stays cheap)
    for _ in range(2):
        sharded_score_stream(sc, CFG4_PAIRS, fetch, env.rank, env.world)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    env.barrier()
    t0 = time.perf_counter()
    e0.record()
    scores = sharded_score_stream(sc, CFG4_PAIRS, fetch, env.rank, env.world)
    top5 = scores.view(CFG4_QUERIES, -1).topk(5, dim=1).indices
    e1.record()
    wall = time.perf_counter() - t0
    env.barrier()
    ms = env.max_ms(max(e0.elapsed_time(e1), wall * 1e3))
    pps = CFG4_PAIRS / (ms * 1e-3)
    per_rank = -(-CFG4_PAIRS // env.world)
    return {"workload": "testB-shape candidate set: 1,000 queries x 30 candidates, 12-layer ImageBert (zk), 32 x 36 x "
                        "2048-d, host feeds, sharded over the ranks + one all-gather + top-5 per query on every rank",
            "pairs": CFG4_PAIRS, "queries": CFG4_QUERIES, "n_gpus": env.world, "pairs_per_rank": per_rank,
            "chunks_per_rank": -(-per_rank // 256), "wall_ms": ms, "pairs_per_s": pps, "scaling": "strong",
            "speedup_vs_n1": _speedup_vs_n1("cfg4", env.world, pps),
            "top5_checksum": int(top5.sum().item()), "finite": bool(torch.isfinite(scores).all())}


def run_cfg5(env, scorers, cfgs):
    """10,000 pairs scored by imagebert_zk, imagebert_lds and lxmert (each sharded over the ranks + one all-gather), then
    the main.py ensemble (zk feeds both of its ImageBertB slots) with the 0.92 product-uniqueness filter and top-5, on
    every rank."""
    from kddcup_2020_multimodalitiesrecall_2nd_place_b200 import ensemble
    from kddcup_2020_multimodalitiesrecall_2nd_place_b200.scorer import sharded_score_stream
    fetchers = {k: block_fetch(candidate_block(scorers[k], cfgs[k], synth.SEED0 + 505)) for k in scorers}
    for _ in range(2):                                        # warm-up = the workload itself, twice (see run_cfg4)
        for k, sc in scorers.items():
            sharded_score_stream(sc, CFG5_PAIRS, fetchers[k], env.rank, env.world)
    per_q = CFG5_PAIRS // CFG5_QUERIES
    qi = np.arange(CFG5_PAIRS, dtype=np.int64) // per_q
    pi = np.arange(CFG5_PAIRS, dtype=np.int64) % 6000           # products recur across queries: the filter has work
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    env.barrier()
    t0 = time.perf_counter()
    e0.record()
    s = {k: sharded_score_stream(sc, CFG5_PAIRS, fetchers[k], env.rank, env.world).numpy().astype(np.float64)
         for k, sc in scorers.items()}
    t_scored = time.perf_counter() - t0
    rows, merged = ensemble.select_flat(qi, pi, np.stack([s[ZK], s[ZK], s[LDS], s[LXMERT]]), 6000)
    e1.record()
    wall = time.perf_counter() - t0
    env.barrier()
    ms = env.max_ms(max(e0.elapsed_time(e1), wall * 1e3))
    pps = CFG5_PAIRS / (ms * 1e-3)
    return {"workload": "3-model ensemble (imagebert_zk + imagebert_lds + lxmert, full depth, 32 x 36 x 2048-d) over "
                        "10,000 pairs, host feeds, each model sharded over the ranks + all-gather, main.py merge + "
                        "0.92 uniqueness filter + top-5 on every rank",
            "pairs": CFG5_PAIRS, "queries": CFG5_QUERIES, "n_gpus": env.world, "wall_ms": ms,
            "ensemble_ms": (wall - t_scored) * 1e3, "pairs_per_s": pps, "model_pair_scores_per_s": 3 * pps,
            "scaling": "strong", "speedup_vs_n1": _speedup_vs_n1("cfg5", env.world, pps),
            "queries_written": len(rows), "finite": bool(np.isfinite(merged).all())}


def main():
    args = parse()
    cfg = baseline_cfg2(args.model)
    if args.impl == "reference":
        run_reference(args, cfg, int(os.environ.get("RANK", "0")))
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference for "
                         "the CPU arm)")
    from kddcup_2020_multimodalitiesrecall_2nd_place_b200.scorer import MatchScorer
    env = Env(args)
    rank, world, dev = env.rank, env.world, env.dev
    B, K, W = args.batch, args.steps, max(args.warmup, 3)
    peaks = load_peaks()

    weights = synth.make_weights(cfg, seed=synth.SEED0)
    sc = MatchScorer(cfg, weights, device=env.local, dtype=args.dtype, max_batch=B, precision=args.precision)
    host_sets, dev_sets = resident_sets(sc, cfg, B, rank, dev)

    # ---- headline: device-resident inputs
    total_ms, clocks, scores = time_resident(env, sc, dev_sets, B, K, W, sample_clocks=True)
    launches = sc.launches_per_forward() * K
    value = world * K * B / (total_ms * 1e-3)
    ms_per_step = total_ms / K

    # ---- end to end: pinned host feeds -> H2D -> kernels -> D2H scores, every step
    e2e = None if args.no_e2e else time_e2e(env, sc, host_sets, B, K)

    sustained = roofline = None
    if rank == 0:
        roofline = roofline_record(sc, cfg, dev_sets, scores, B, value / world, ms_per_step, clocks, total_ms * 1e-3, peaks)
    if not args.quick:
        sustained = sustained_record(env, sc, cfg, dev_sets, B, peaks)     # every rank runs it: same load on every GPU
        if roofline is not None:
            roofline["sustained"] = sustained

    # ---- the other scorers, the strict-precision mode, and the sharded candidate-set workloads
    other_models, strict, cfg4, cfg5 = {}, None, None, None
    if not args.quick:
        others = {}
        for kind in (ZK, LDS, LXMERT):
            if kind == cfg.kind:
                continue
            ocfg = baseline_cfg2(kind)
            osc = MatchScorer(ocfg, synth.make_weights(ocfg, seed=synth.SEED0), device=env.local, dtype=args.dtype,
                              max_batch=B)
            others[kind] = (osc, ocfg)
            if world == 1:
                _, osets = resident_sets(osc, ocfg, B, rank, dev)
                Ko = max(6, min(K, 20))
                oms, oclk, _ = time_resident(env, osc, osets, B, Ko, 3, sample_clocks=True, gather=False)
                pps = Ko * B / (oms * 1e-3)
                tf = flops_per_pair(ocfg) * pps / 1e12
                opeak, osrc = pick_peak(peaks, oclk, oms * 1e-3)
                other_models[kind] = {"workload": workload_name(ocfg, B), "value": pps, "unit": UNIT, "steps": Ko,
                                      "ms_per_step": oms / Ko, "flops_per_pair": flops_per_pair(ocfg),
                                      "algorithmic_tflops": tf, "peak": opeak, "peak_source": osrc,
                                      "frac_of_peak": tf / opeak, "launches_per_forward": osc.launches_per_forward(),
                                      "clocks": oclk}
                del osets
        if world == 1 and args.precision == "fast":
            ssc = MatchScorer(cfg, weights, device=env.local, dtype=args.dtype, max_batch=B, precision="strict")
            sms, sclk, sscores = time_resident(env, ssc, dev_sets, B, 6, 3, sample_clocks=True, gather=False)
            fast_p = sc.forward_device(dev_sets[5 % N_SETS]).clone()
            torch.cuda.synchronize()
            strict = {"what": "precision='strict': two-term split operands on every GEMM, fp32 attention, precise GELU",
                      "value": 6 * B / (sms * 1e-3), "unit": UNIT, "ms_per_step": sms / 6,
                      "throughput_vs_fast": (6 * B / (sms * 1e-3)) / (value / world),
                      "max_abs_dscore_vs_fast": float((sscores[5] - fast_p).abs().max().item()),
                      "launches_per_forward": ssc.launches_per_forward(), "clocks": sclk}
            ssc.close()
            del ssc
        if cfg.kind == ZK:
            cfg4 = run_cfg4(env, sc, cfg)
            scorers = {ZK: sc, LDS: others[LDS][0], LXMERT: others[LXMERT][0]}
            cfg5 = run_cfg5(env, scorers, {ZK: cfg, LDS: others[LDS][1], LXMERT: others[LXMERT][1]})
        for osc, _ in others.values():
            osc.close()

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
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.dtype, "data": "synthetic",
            "config": config_record(cfg, B, args, world, sc.use_graphs),
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "launches_per_forward": sc.launches_per_forward(),
            "roofline": roofline, "cpu_baseline": cpu, "other_models": other_models or None, "strict": strict,
            "cfg4": cfg4, "cfg5": cfg5,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        env.dist.destroy_process_group()


if __name__ == "__main__":
    main()
