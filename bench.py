#!/usr/bin/env python
"""bench.py -- throughput of the fnft_nsev hot path (BASELINE.json metric).

Workload (BASELINE.json configs[1]): fnft_nsev reflection coefficient,
2SPLIT4B, D = M = 16384, T = [-32, 32], XI = [-10, 10], kappa = +1, synthetic sech /
random-phase signals (seed 16384), 4096 signals per GPU per step ("weak" scaling: the
batch is independent signals, sharded by rank, no collective on the data path).

  python bench.py --gpus N --steps K --warmup W            our CUDA path
  python bench.py --impl reference --gpus N --steps K ...  the reference's CPU path
                                                           (oracle/_ref, all host cores)

One JSON line on stdout (rank 0).  `value` = device-resident signals/s (inputs already
in HBM), `e2e` = through the C-ABI fnft_nsev_batch with pinned HOST buffers (H2D + D2H
inside the timed region), `roofline` = product-tree algorithmic bytes / tree kernel
time (CUDA events per launch) against the measured HBM copy bandwidth, `cpu_baseline` =
the reference library timed on this box's host cores on a bounded sample.
"""
import argparse
import ctypes as C
import json
import math
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

D = 16384
M = 16384
TT = (-32.0, 32.0)
XI = (-10.0, 10.0)
KAPPA = +1
BATCH_PER_GPU = 4096
SEED = 16384
DEG0 = 2  # 2SPLIT4B


def tree_bytes_per_signal(d0=DEG0, dd=D):
    # one read + one write of every tree level
    return 16 * (8 * d0 * dd * int(math.log2(dd)) + 12 * dd - 12)


# ----------------------------------------------------------------------------------
# synthetic signals: parameters from numpy, samples from
# numpy (few signals, CPU legs) or torch (whole batch, on the GPU)
# ----------------------------------------------------------------------------------
def signal_params(B, seed=SEED):
    rng = np.random.default_rng(seed)
    return dict(
        A=rng.uniform(0.5, 5.4, B), lam0=rng.uniform(-3, 3, B), phi=rng.uniform(0, 2 * np.pi, B),
        w=rng.uniform(0.5, 2.0, B), c=rng.normal(0, 0.5, (B, 8)), psi=rng.uniform(0, 2 * np.pi, (B, 8)))


def signals_numpy(P, idx):
    t = np.linspace(TT[0], TT[1], D)
    out = np.empty((len(idx), D), dtype=np.complex128)
    for o, b in enumerate(idx):
        if b % 2 == 0:
            out[o] = P["A"][b] / np.cosh(t) * np.exp(-2j * P["lam0"][b] * t + 1j * P["phi"][b])
        else:
            th = np.zeros(D)
            for k in range(8):
                th += P["c"][b, k] * np.sin(2 * np.pi * (k + 1) * t / 64 + P["psi"][b, k])
            out[o] = P["A"][b] / np.cosh(t / P["w"][b]) * np.exp(1j * th)
    return out


def signals_torch(P, B, device):
    import torch
    t = torch.linspace(TT[0], TT[1], D, dtype=torch.float64, device=device)[None, :]
    tt = lambda a: torch.as_tensor(a, dtype=torch.float64, device=device)
    q = torch.empty((B, D), dtype=torch.complex128, device=device)
    step = 512
    for b0 in range(0, B, step):
        b1 = min(B, b0 + step)
        A = tt(P["A"][b0:b1])[:, None]
        even = A / torch.cosh(t) * torch.exp(1j * (-2 * tt(P["lam0"][b0:b1])[:, None] * t
                                                  + tt(P["phi"][b0:b1])[:, None]))
        th = torch.zeros((b1 - b0, D), dtype=torch.float64, device=device)
        for k in range(8):
            th += tt(P["c"][b0:b1, k])[:, None] * torch.sin(
                2 * math.pi * (k + 1) * t / 64 + tt(P["psi"][b0:b1, k])[:, None])
        odd = A / torch.cosh(t / tt(P["w"][b0:b1])[:, None]) * torch.exp(1j * th)
        is_even = (torch.arange(b0, b1, device=device) % 2 == 0)[:, None]
        q[b0:b1] = torch.where(is_even, even, odd)
    return q


# ----------------------------------------------------------------------------------
# CPU reference leg (oracle/_ref): process pool over the host cores
# ----------------------------------------------------------------------------------
def _ref_worker(args):
    sys.path.insert(0, ROOT)
    from oracle import ref_lib as R
    q, = args
    o = R.nsev_default_opts()
    t0 = time.perf_counter()
    ret, cs, *_ = R.nsev(q, TT, M, XI, KAPPA, o, K=0)
    return ret, time.perf_counter() - t0, cs


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_reference_sample(nsig, cores, P=None):
    """Times the reference on `nsig` signals with `cores` worker processes.
    Returns (signals_per_s, wall_s, outputs)."""
    import multiprocessing as mp
    if P is None:
        P = signal_params(max(nsig, 2))
    q = signals_numpy(P, list(range(nsig)))
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        pool.map(_ref_worker, [(q[0],)] * min(cores, nsig))  # warm-up: load lib, page in
        t0 = time.perf_counter()
        res = pool.map(_ref_worker, [(q[i],) for i in range(nsig)], chunksize=1)
        wall = time.perf_counter() - t0
    if any(r[0] != 0 for r in res):
        raise RuntimeError("reference returned an error code")
    return nsig / wall, wall, [r[2] for r in res]


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


# ----------------------------------------------------------------------------------
# clocks sampling (pynvml) during the timed region
# ----------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._th = None
        self.index = index
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4),
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.1)

    def start(self):
        if self.nv is not None:
            self._th = threading.Thread(target=self._loop, daemon=True)
            self._th.start()

    def stop(self):
        self._stop.set()
        if self._th is not None:
            self._th.join()
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ----------------------------------------------------------------------------------
def measured_hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def parse_report(txt):
    out = {}
    for line in txt.decode().splitlines():
        name, cnt, ms = line.split()
        out[name] = (int(cnt), float(ms))
    return out


_JSON_FD = None


def _claim_stdout():
    """Rank 0 must print exactly ONE line on stdout.  Libraries write there too (NCCL prints its
    version banner on stdout when NCCL_DEBUG is set), so file descriptor 1 is pointed at stderr
    for the rest of the run and the JSON line goes to a private duplicate of the real stdout."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def _emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH_PER_GPU, help="signals per GPU per step")
    ap.add_argument("--ref-signals", type=int, default=0, help="signals per reference step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cores = host_cores()
    config = {"workload": "fnft_nsev reflection coefficient, 2SPLIT4B, D=M=16384, T=[-32,32], "
                          "XI=[-10,10], kappa=+1, synthetic sech/random-phase signals (seed 16384), "
                          "%d signals per GPU per step" % args.batch,
              "D": D, "M": M, "batch_per_gpu": args.batch, "discretization": "2SPLIT4B",
              "l2": "inputs (%.0f MiB per step per GPU) larger than L2" % (args.batch * D * 16 / 2**20)}

    # ------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return 0
        nsig = args.ref_signals or max(8 * cores, 64)
        for _ in range(max(args.warmup, 0) and 1):
            run_reference_sample(min(nsig, cores), cores)
        t_tot, n_tot = 0.0, 0
        for _ in range(args.steps):
            rate, wall, _ = run_reference_sample(nsig, cores)
            t_tot += wall
            n_tot += nsig
        val = n_tot / t_tot
        sample = "%d signals per step (same generator as the GPU arm), %d steps, one signal per task" % (
            nsig, args.steps)
        line = {"impl": "reference", "metric": "fnft_nsev signals/sec at D=M=16384", "value": val,
                "unit": "signals/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": 1e3 * t_tot / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": val, "unit": "signals/s", "cores": cores,
                                 "kind": "reference", "sample": sample, "cpu": cpu_model(),
                                 "build": "oracle/_ref/libfnft_ref.so: unmodified FNFT 0.4.1 C sources, "
                                          "gcc -O3 -march=x86-64-v3, Kiss FFT"},
                "e2e": {"value": val, "unit": "signals/s", "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0}}
        _emit(line)
        return 0

    # ------------------------------------------------------------------ CPU baseline first
    # (before CUDA is initialised in this process: the pool is forked)
    cpu_baseline = None
    P_all = signal_params(args.batch)
    ref_out = None
    if rank == 0 and not args.no_cpu_baseline:
        from oracle import ref_lib as R
        if R.available():
            nsig = min(args.batch, max(8 * cores, 64))  # ~20 core-seconds of reference work
            rate, wall, ref_out = run_reference_sample(nsig, cores, P_all)
            cpu_baseline = {"value": rate, "unit": "signals/s", "cores": cores, "kind": "reference",
                            "sample": "first %d signals of the batch, one signal per task over %d "
                                      "worker processes, %.1f s wall" % (nsig, cores, wall),
                            "cpu": cpu_model()}

    # ------------------------------------------------------------------ GPU arm
    import torch
    import fnft_b200 as F
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the hot path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    L = F.lib()
    if L.fnft_b200_set_device(local_rank) != 0:
        raise SystemExit("fnft_b200_set_device failed")
    B = args.batch
    # every rank works on its own shard: signal parameters are offset by rank
    P = signal_params(B * world)
    P = {k: v[rank * B:(rank + 1) * B] for k, v in P.items()}
    q_dev = signals_torch(P, B, dev)
    out_dev = torch.zeros((B, M), dtype=torch.complex128, device=dev)
    Tarr = np.array(TT, dtype=np.float64)
    XIarr = np.array(XI, dtype=np.float64)
    opts = L.fnft_nsev_default_opts()  # 2SPLIT4B, reflection coefficient, normalisation on
    stream = torch.cuda.ExternalStream(L.fnft_b200_stream(), device=dev)

    def step_device():
        rc = L.fnft_nsev_batch(B, D, q_dev.data_ptr(), Tarr.ctypes.data, M, out_dev.data_ptr(),
                               XIarr.ctypes.data, None, 0, None, None, KAPPA, C.addressof(opts), None)
        if rc != 0:
            raise SystemExit("fnft_nsev_batch (device pointers) failed with code %d" % rc)

    def barrier():
        if world > 1:
            dist.barrier()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident throughput ("value")
    L.fnft_b200_set_device_pointers(1)
    for _ in range(args.warmup):
        step_device()
    L.fnft_b200_synchronize()
    torch.cuda.synchronize()
    sampler = ClockSampler(torch.cuda.current_device() if "CUDA_VISIBLE_DEVICES" not in os.environ
                           else local_rank)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    torch.cuda.synchronize()
    launches0 = L.fnft_b200_launch_count()
    sampler.start()
    ev0.record(stream)
    for _ in range(args.steps):
        step_device()
    ev1.record(stream)
    L.fnft_b200_synchronize()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    barrier()
    launches = int(L.fnft_b200_launch_count() - launches0)
    ms_dev = max_over_ranks(ev0.elapsed_time(ev1))
    value = world * B * args.steps / (ms_dev * 1e-3)

    # ---- parity spot check against the reference outputs computed for the CPU baseline
    parity = None
    if ref_out is not None:
        got = out_dev[:len(ref_out)].cpu().numpy()
        errs = [float(np.abs(got[i] - ref_out[i]).sum() / np.abs(ref_out[i]).sum())
                for i in range(len(ref_out))]
        parity = {"metric": "misc_rel_err(ours, reference) per signal, max over the sample",
                  "max": max(errs), "signals": len(errs), "bound": 1e-9}

    # ---- per-kernel timing (CUDA events around every launch) for the roofline
    L.fnft_b200_profile_enable(1)
    step_device()
    rep = parse_report(L.fnft_b200_profile_report())
    L.fnft_b200_profile_enable(0)
    tree_ms = sum(ms for k, (n, ms) in rep.items() if k.startswith("tree_"))
    tree_launches = sum(n for k, (n, ms) in rep.items() if k.startswith("tree_"))
    total_ms = sum(ms for k, (n, ms) in rep.items())
    peak, peak_src = measured_hbm_peak()
    bts = tree_bytes_per_signal()
    achieved = bts * B / (tree_ms * 1e-3) / 1e9
    # physical DRAM traffic of the tree kernels: per-signal bytes from the committed ncu capture
    # (profiles/tree_dram_bytes.json, dram__bytes_read.sum + dram__bytes_write.sum), times the
    # signals of one step -- "per launch set" like `achieved`
    traffic, traffic_src = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "tree_dram_bytes.json")) as f:
            tj = json.load(f)
        traffic = float(tj["tree_bytes_per_signal"]) * B
        traffic_src = tj.get("source")
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": "fmult2x2 product tree (all tree_* launches of one step)",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "peak_source": peak_src, "traffic": traffic, "traffic_source": traffic_src,
                "physical_gbs": (traffic / (tree_ms * 1e-3) / 1e9) if traffic else None,
                "algorithmic_bytes_per_signal": bts, "tree_ms_per_step": tree_ms,
                "tree_launches_per_step": tree_launches, "tree_share_of_step": tree_ms / total_ms,
                "kernel_ms_per_step": {k: round(ms, 4) for k, (n, ms) in sorted(rep.items())}}

    # ---- end to end through the C-ABI with pinned host buffers ("e2e")
    L.fnft_b200_set_device_pointers(0)
    q_host = torch.empty((B, D), dtype=torch.complex128, pin_memory=True)
    q_host.copy_(q_dev)
    out_host = torch.empty((B, M), dtype=torch.complex128, pin_memory=True)

    def step_host():
        rc = L.fnft_nsev_batch(B, D, q_host.data_ptr(), Tarr.ctypes.data, M, out_host.data_ptr(),
                               XIarr.ctypes.data, None, 0, None, None, KAPPA, C.addressof(opts), None)
        if rc != 0:
            raise SystemExit("fnft_nsev_batch (host pointers) failed with code %d" % rc)

    out_host.zero_()  # touch the pinned pages before the first DMA
    for _ in range(args.warmup):
        step_host()
    torch.cuda.synchronize()
    barrier()
    t0 = time.perf_counter()
    step_ms = []
    for _ in range(args.steps):
        t1 = time.perf_counter()
        step_host()
        step_ms.append(round((time.perf_counter() - t1) * 1e3, 2))
    torch.cuda.synchronize()
    ms_e2e = max_over_ranks((time.perf_counter() - t0) * 1e3)
    barrier()
    e2e = {"value": world * B * args.steps / (ms_e2e * 1e-3), "unit": "signals/s",
           "h2d_bytes_per_step": B * D * 16, "d2h_bytes_per_step": B * M * 16,
           "ms_per_step": ms_e2e / args.steps, "ms_each_step_rank0": step_ms,
           "api": "fnft_nsev_batch (C-ABI, libfnft_b200.so) with pinned host buffers"}

    if rank == 0:
        line = {"metric": "fnft_nsev signals/sec at D=M=16384", "value": value, "unit": "signals/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline,
                "cpu_baseline": cpu_baseline, "parity": parity}
        _emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
