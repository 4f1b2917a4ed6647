#!/usr/bin/env python
"""bench.py — DMEL forward+backward throughput (audio-seconds per second) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

One "step" = one fused forward + one fused backward (dL/dx AND dL/dlambda, log fused, eps 1e-10) over one batch of
synthetic clips of a BASELINE.json shape, ending with the path's one collective when N > 1: the sum of the scalar
dL/dlambda over the GPUs.  Prints ONE JSON line (rank 0).

Without --workload the headline is BASELINE.json configs[4] "large-batch throughput": 1024 clips of 10 s @ 16 kHz per GPU
(weak scaling; at 8 GPUs that is exactly the 8192-clip configuration), and the line carries one-shot sub-results for
configs[1] audio_mnist (256 clips per GPU) and configs[2] esc50 (128 clips split evenly over the ranks, i.e. strong
scaling as SURVEY.md §8d lays it out) under extra.workloads.  --workload NAME measures that one shape only.

  value      whole-job audio-s/s, inputs resident in HBM, C-ABI calls replayed from a CUDA graph; buffers rotate over
             a set > 2x L2
  e2e        same metric through the public nn.Module (MelSpectrogramLayer.forward + .backward) from PINNED HOST
             clips: H2D of x every step and D2H of the step's scalar result (dL/dlambda) inside the timed region
             (extras: e2e_pcm16 = int16 host clips, widened INSIDE the kernels; lambda_only = only lambd trains)
  roofline   dominant kernel: algorithmic bytes per launch / its mean CUDA-event duration vs the measured HBM copy
             peak; next to it what actually bounds these kernels (shared-memory data pipe, FP32 pipe) from the ncu
             captures under profiles/
  cpu_baseline  the CPU oracle port of the reference path (oracle/dmel_oracle.py: per-sample loop, torch.stft, filterbank
             rebuilt per sample, autograd backward) on this box's host cores, on a bounded sample — measured on rank 0
             BEFORE the process group exists, so no other rank spins meanwhile
  collective (N > 1) default: the peer-memory all-reduce kernel on a side branch of the step's graph, so that its flag
             exchange overlaps the next step's prep/forward; --collective fused puts it inside the backward's finalize
             kernel, --collective nccl uses the process group.  collective_check compares all three on one step.
`--impl reference` times only the CPU port (the reference is pure Python and /root/reference does not travel to the GPU
box; the port is pinned to the reference's outputs by tests/test_oracle_golden.py).
"""
import argparse
import ctypes
import json
import math
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "dmel_fwd_bwd_audio_seconds_per_second"
UNIT = "audio-s/s"
LOG_EPS = 1e-10
L2_BYTES = 126 * 1024 * 1024
FP32_PEAK_TFLOPS = 74.5          # 148 SMs x 128 lanes x 2 flop x 1.965 GHz
HEADLINE = "large_batch"


def workload(name, world=1):
    """BASELINE.json configs -> concrete synthetic shapes. B is the per-GPU batch."""
    if name == "tiny":             # CI-sized case (tests/test_bench_cli.py); not a BASELINE config
        return dict(B=4, N=1600, sr=8000, hop=80, M=16, lam=8000 * 0.035 / 6, n_fft=512, scaling="weak")
    if name == "audio_mnist":      # configs[1]: 1 s clips, batch 256, n_fft 512, 64 mels
        return dict(B=256, N=8000, sr=8000, hop=80, M=64, lam=8000 * 0.035 / 6, n_fft=512, scaling="weak")
    if name == "esc50":            # configs[2]: 5 s @ 44.1 kHz, batch 128, n_fft 2048, 128 mels
        return dict(B=128, N=220500, sr=44100, hop=441, M=128, lam=44100 * 0.035 / 6, n_fft=2048, scaling="weak")
    if name == "esc50_split":      # configs[2] at 1/2/4/8 GPUs: the 128 clips split evenly over the ranks
        return dict(B=max(1, 128 // world), N=220500, sr=44100, hop=441, M=128, lam=44100 * 0.035 / 6, n_fft=2048,
                    scaling="strong")
    if name == "large_batch":      # configs[4]: 8192 x 10 s @ 16 kHz over 8 GPUs = 1024 clips per GPU
        return dict(B=1024, N=160000, sr=16000, hop=160, M=128, lam=16000 * 0.035 / 6, n_fft=1024, scaling="weak")
    if name == "esc50_300ms":      # configs[2] lambda sweep, 300 ms window at 44.1 kHz -> n_fft 16384 (big-FFT path)
        return dict(B=128, N=220500, sr=44100, hop=441, M=128, lam=44100 * 0.300 / 6, n_fft=16384, scaling="weak")
    if name == "esc50_10ms":       # configs[2] lambda sweep, 10 ms window -> n_fft 512
        return dict(B=128, N=220500, sr=44100, hop=441, M=128, lam=44100 * 0.010 / 6, n_fft=512, scaling="weak")
    if name.startswith("sweep_"):  # configs[3]: batch 512, n_fft 128..4096, lambda = 0.75 n_fft / 6
        n_fft = int(name.split("_")[1])
        return dict(B=512, N=8000, sr=8000, hop=80, M=64, lam=0.75 * n_fft / 6, n_fft=n_fft, scaling="weak")
    raise SystemExit(f"unknown workload {name}")


def algorithmic_bytes(w, which):
    """compulsory fp32 traffic of a perfectly fused implementation, per batch."""
    N, M, T, B = w["N"], w["M"], w["N"] // w["hop"] + 1, w["B"]
    per_clip = {"fwd": 4 * (N + M * T), "bwd": 4 * (2 * N + M * T), "step": 4 * (3 * N + 2 * M * T)}[which]
    return per_clip * B


def useful_fft_flops(w, which):
    """2.5 n log2 n per real-FFT equivalent; forward 1, backward 2 (recompute + inverse) per frame."""
    T, n = w["N"] // w["hop"] + 1, w["n_fft"]
    per_frame = 2.5 * n * math.log2(n)
    return {"fwd": 1, "bwd": 2, "step": 3}[which] * per_frame * T * w["B"]


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_metrics(workload_name, kernel):
    """Per-kernel numbers read out of the committed ncu captures (profiles/ncu_metrics.json): DRAM bytes per launch,
    FP32-pipe and shared-memory-pipe utilisation.  None when no capture exists for that workload."""
    p = os.path.join(ROOT, "profiles", "ncu_metrics.json")
    try:
        return json.load(open(p)).get(workload_name, {}).get(kernel)
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed regions run."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={gpu_index}", f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            c = [v.strip() for v in line.split(",")]
            if len(c) < 7:
                continue
            try:
                sm.append(float(c[0]))
                mx.append(float(c[1]))
            except ValueError:
                continue
            for n, v in zip(names, c[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        if sm:
            # median over the samples taken under load (upper half: the idle gaps between regions read low)
            s = sorted(sm)
            out.update(sm_mhz=s[len(s) // 2], sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


# ------------------------------------------------------------------------------------------ CPU arm
def cpu_port_rate(w, budget_s, threads):
    """Oracle port (reference structure) fwd+bwd with dx and dlambda on a bounded sample; best of the
    repetitions that fit in ~budget_s.  Returns (audio-s/s, clips in the sample, repetitions, seconds/clip)."""
    import torch
    from oracle import dmel_oracle as orc
    torch.set_num_threads(threads)      # torchrun exports OMP_NUM_THREADS=1: the explicit call overrides it
    g = torch.Generator().manual_seed(1234)
    T = w["N"] // w["hop"] + 1
    kw = dict(hop=w["hop"], optimized=True, norm=False, n_mels=w["M"], sample_rate=w["sr"])

    def one_pass(x, G):
        xx = x.clone().requires_grad_(True)
        lam = torch.tensor(w["lam"], requires_grad=True)
        s = orc.port_forward(xx, lam, **kw)
        (torch.log(s + LOG_EPS) * G).sum().backward()
        return float(lam.grad)

    x1 = torch.randn(2, w["N"], generator=g)
    G1 = torch.randn(2, 1, w["M"], T, generator=g)
    one_pass(x1, G1)                                   # warm-up (thread pools, FFT plans)
    t0 = time.perf_counter()
    one_pass(x1, G1)
    per_clip = (time.perf_counter() - t0) / 2
    mem_cap = int(1.5e9 / (T * (w["n_fft"] // 2 + 1) * 8 * 3))      # autograd graph of the sample stays ~1.5 GB
    n = int(max(2, min(w["B"], (budget_s / 3.0) / max(per_clip, 1e-6), mem_cap)))
    x = torch.randn(n, w["N"], generator=g)
    G = torch.randn(n, 1, w["M"], T, generator=g)
    best, reps, t_start = float("inf"), 0, time.perf_counter()
    while reps < 5 and (reps == 0 or time.perf_counter() - t_start < budget_s):
        t0 = time.perf_counter()
        one_pass(x, G)
        best = min(best, time.perf_counter() - t0)
        reps += 1
    return n * (w["N"] / w["sr"]) / best, n, reps, best / n


def run_reference_arm(args, wname, w):
    """`--impl reference`: the CPU implementation of the path on the host cores, same metric/config keys."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from oracle import dmel_oracle as orc
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    T = w["N"] // w["hop"] + 1
    kw = dict(hop=w["hop"], optimized=True, norm=False, n_mels=w["M"], sample_rate=w["sr"])
    g = torch.Generator().manual_seed(1234)
    # size the per-step sample so that (steps + warmup) steps end within ~2 minutes
    rate, _, _, per_clip = cpu_port_rate(w, 4.0, cores)
    # ... and so that one step is a few seconds of work and its autograd graph (complex spectrogram + frames + power per
    # clip, kept for the backward) stays under a GB: the per-sample loop's cost is linear in clips, so the rate does not
    # depend on the sample size, but a sample that swaps would
    budget = min(5.0, 100.0 / max(1, args.steps + args.warmup))
    mem_cap = int(0.75e9 / (T * (w["n_fft"] // 2 + 1) * 8 * 3))
    n = int(max(1, min(w["B"], budget / per_clip, mem_cap)))
    x = torch.randn(n, w["N"], generator=g)
    G = torch.randn(n, 1, w["M"], T, generator=g)

    def step():
        xx = x.clone().requires_grad_(True)
        lam = torch.tensor(w["lam"], requires_grad=True)
        s = orc.port_forward(xx, lam, **kw)
        (torch.log(s + LOG_EPS) * G).sum().backward()

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = args.steps * n * (w["N"] / w["sr"]) / dt
    sample = f"{n} of {w['B']} clips per step (serial per-sample loop: cost linear in clips)"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": w["scaling"],
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(wname, w, extra={"device": "cpu", "sample": sample}),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def config_dict(wname, w, extra=None):
    c = {"workload": wname, "per_gpu_batch": w["B"], "n_points": w["N"], "sample_rate": w["sr"],
         "hop_length": w["hop"], "n_mels": w["M"], "n_fft": w["n_fft"], "lambda": w["lam"],
         "mode": "forward+backward (dL/dx and dL/dlambda), fused log eps=1e-10, optimized=True"}
    if extra:
        c.update(extra)
    return c


def default_steps(w):
    # ~1.5 s of device time at ~3e6 audio-s/s, so the clock sampler sees the run
    audio_s = w["B"] * w["N"] / w["sr"]
    return int(min(5000, max(20, 4.0e6 / audio_s)))


# ------------------------------------------------------------------------------------------ GPU arm
class Ctx:
    """Per-process state shared by the measurements."""
    pass


class Resident:
    """One workload with inputs resident in HBM: rotating buffer sets (> 2x L2), the step as C-ABI calls, optionally
    replayed from a CUDA graph."""

    def __init__(self, cx, wname, w, steps, collective):
        import torch
        from importlib import import_module
        self.cx, self.wname, self.w, self.collective = cx, wname, w, collective
        lib, dev, dm = cx.lib, cx.dev, cx.dm
        B, N, sr, hop, M, lam = w["B"], w["N"], w["sr"], w["hop"], w["M"], w["lam"]
        self.T = T = N // hop + 1
        self.geo = geo = dm.resolve_geometry(B, N, hop, lam, True, M)
        assert geo.n_fft == w["n_fft"], (geo.n_fft, w["n_fft"])
        self.bank = import_module("differentiable-mel-spectrogram_b200.melbank").device_bank(geo.n_freqs, 0.0, sr // 2, M, sr, dev)
        self.set_bytes = 4 * (2 * B * N + 2 * B * M * T)
        R = max(2, math.ceil(2.2 * L2_BYTES / self.set_bytes))
        R = min(R, steps)
        while steps % R:
            R += 1
        self.R = R
        self.ws_bytes = lib.dmel_workspace_bytes(B, geo.n_fft)
        gen = torch.Generator(device=dev).manual_seed(1234 + cx.rank)
        ggen = torch.Generator(device=dev).manual_seed(4321 + cx.rank)
        self.xs = [torch.randn(B, N, device=dev, generator=gen) for _ in range(R)]
        self.Gs = [torch.randn(B, 1, M, T, device=dev, generator=ggen) for _ in range(R)]
        self.ys = [torch.empty(B, 1, M, T, device=dev) for _ in range(R)]
        self.dxs = [torch.empty(B, N, device=dev) for _ in range(R)]
        self.dls = [torch.zeros((), device=dev) for _ in range(R)]
        self.wss = [torch.empty(self.ws_bytes, dtype=torch.uint8, device=dev) for _ in range(R)]
        # the side branch exists whenever the ring does: collective_check exercises it whatever the timed mode is
        self.side = torch.cuda.Stream(dev) if (cx.world > 1 and cx.ring is not None) else None
        self.graph = None

    def args_fwd(self, s, st):
        w, g = self.w, self.geo
        return (self.xs[s].data_ptr(), w["B"], w["N"], w["N"], w["lam"], g.win_length, g.n_fft, w["hop"], w["M"], self.bank.ref(),
                LOG_EPS, 0, self.ys[s].data_ptr(), self.wss[s].data_ptr(), self.ws_bytes, st)

    def args_bwd(self, s, st, dx=True):
        w, g = self.w, self.geo
        return (self.xs[s].data_ptr(), w["B"], w["N"], w["N"], w["lam"], g.win_length, g.n_fft, w["hop"], w["M"], self.bank.ref(),
                LOG_EPS, 0, self.ys[s].data_ptr(), self.Gs[s].data_ptr(), self.dxs[s].data_ptr() if dx else None,
                self.dls[s].data_ptr(), self.wss[s].data_ptr(), self.ws_bytes, 1, st)

    def step(self, i, collective=None):
        import torch
        import torch.distributed as dist
        cx, lib = self.cx, self.cx.lib
        mode = collective or self.collective
        s = i % self.R
        cur = torch.cuda.current_stream(cx.dev)
        st = int(cur.cuda_stream)
        rc = lib.dmel_forward_f32(*self.args_fwd(s, st))
        assert rc == 0, lib.dmel_status_string(rc)
        # the path's only collective is the sum of the scalar lambda gradient over the GPUs
        if cx.world > 1 and mode == "fused" and cx.ring is not None:
            rc = lib.dmel_backward_allreduce_f32(*self.args_bwd(s, st), cx.ring.bases.data_ptr(), cx.rank, cx.world)
            assert rc == 0, lib.dmel_status_string(rc)
            return
        rc = lib.dmel_backward_f32(*self.args_bwd(s, st))
        assert rc == 0, lib.dmel_status_string(rc)
        if cx.world == 1:
            return
        if mode == "side" and cx.ring is not None:
            # fork: the one-warp exchange kernel waits for the peers on a side stream while this stream goes on with the
            # next step's prep / forward; joined by join_side() before the slot's gradient could be read or reused
            self.side.wait_stream(cur)
            rc = lib.dmel_peer_allreduce_f32(self.dls[s].data_ptr(), cx.ring.bases.data_ptr(), cx.rank, cx.world,
                                             int(self.side.cuda_stream))
            assert rc == 0, lib.dmel_status_string(rc)
        else:
            dist.all_reduce(self.dls[s])            # NCCL

    def join_side(self):
        import torch
        if self.side is not None:
            torch.cuda.current_stream(self.cx.dev).wait_stream(self.side)

    def rotation(self):
        for i in range(self.R):
            self.step(i)
        self.join_side()

    def capture(self):
        import torch
        cx = self.cx
        try:
            side = torch.cuda.Stream(cx.dev)
            side.wait_stream(torch.cuda.current_stream(cx.dev))
            with torch.cuda.stream(side):
                self.rotation()
            torch.cuda.current_stream(cx.dev).wait_stream(side)
            torch.cuda.synchronize(cx.dev)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                self.rotation()
            graph.replay()
            torch.cuda.synchronize(cx.dev)
            self.graph = graph
        except Exception as e:  # capture of an NCCL call can fail on some stacks: run eagerly, and say so
            if cx.rank == 0:
                print(f"[bench] CUDA graph capture failed ({type(e).__name__}: {e}); eager launches", file=sys.stderr)
            self.graph = None
            torch.cuda.synchronize(cx.dev)

    def loop(self, k):
        if self.graph is not None:
            for _ in range(k // self.R):
                self.graph.replay()
        else:
            for i in range(k):
                self.step(i)
            self.join_side()

    def timed(self, steps, warmup, use_graph):
        """W warm-up steps, then exactly `steps` steps between barrier + synchronize; max over ranks."""
        import torch
        import torch.distributed as dist
        cx = self.cx
        for i in range(warmup):
            self.step(i)
        self.join_side()
        torch.cuda.synchronize(cx.dev)
        if use_graph:
            self.capture()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        cx.barrier()
        ev0.record()
        self.loop(steps)
        ev1.record()
        cx.barrier()
        ms = ev0.elapsed_time(ev1)
        if cx.world > 1:
            t = torch.tensor([ms], device=cx.dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t[0])
        return ms

    def kernel_times(self, steps):
        """Per-kernel durations: eager steps, every main kernel bracketed by CUDA events on its stream (these brackets
        add launch gaps: the kernels' SHARES are exact, their sum is a few % above the graph-replayed step)."""
        import torch
        lib, cx = self.cx.lib, self.cx
        lib.dmel_profile_enable(1)
        for i in range(steps):
            self.step(i, collective="none" if cx.world == 1 else None)
        self.join_side()
        torch.cuda.synchronize(cx.dev)
        kern = {}
        for kind, name in ((0, "prep"), (1, "fwd"), (2, "bwd"), (3, "finalize"), (4, "zero")):
            tot, cnt = ctypes.c_double(), ctypes.c_longlong()
            assert lib.dmel_profile_read(kind, ctypes.byref(tot), ctypes.byref(cnt)) == 0
            kern[name] = (tot.value / max(cnt.value, 1), cnt.value)
        lib.dmel_profile_enable(0)
        return kern

    def launches_per_step(self):
        import torch
        lib = self.cx.lib
        c0 = lib.dmel_launch_count(-1)
        self.step(0)
        self.join_side()
        torch.cuda.synchronize(self.cx.dev)
        return int(lib.dmel_launch_count(-1) - c0)

    def roofline(self, kern, step_ms):
        w = self.w
        peak, peak_src = measured_peaks()
        dom = "bwd" if kern["bwd"][0] >= kern["fwd"][0] else "fwd"
        dom_ms = kern[dom][0]
        achieved = algorithmic_bytes(w, dom) / (dom_ms * 1e-3) / 1e9
        nm = ncu_metrics(self.wname.replace("_split", ""), dom) or {}
        return {"bound": "hbm", "kernel": f"dmel2_{dom}_kernel<{w['n_fft']}>", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": nm.get("dram_bytes"), "peak_source": peak_src,
                "algorithmic_bytes_per_launch": algorithmic_bytes(w, dom),
                "step_frac": algorithmic_bytes(w, "step") / (step_ms * 1e-3) / 1e9 / peak,
                # what binds these kernels is on-chip (profiles/): the FFT's exchanges saturate the shared-memory data pipe
                "fp32_frac": useful_fft_flops(w, dom) / (dom_ms * 1e-3) / 1e12 / FP32_PEAK_TFLOPS,
                "fp32_pipe_pct": nm.get("fp32_pipe_pct"), "l1tex_pct": nm.get("l1tex_pct"),
                "limiter": "shared-memory data pipe (l1tex) per the ncu captures in profiles/; kernels_ms are eager "
                           "event brackets, ~5-20 % above their share of the graph-replayed step, so frac is slightly pessimistic"}


def collective_check(cx, res):
    """One step three ways on every rank — exchange fused into finalize, exchange on the side stream, NCCL — plus the
    rank-local values summed on the host; and a small sharded batch against the same batch on one rank."""
    import torch
    import torch.distributed as dist
    dev, lib, dm = cx.dev, cx.lib, cx.dm
    out = {}
    vals = {}
    for mode in ("fused", "side", "nccl"):
        if mode != "nccl" and cx.ring is None:
            continue
        res.dls[0].zero_()
        res.step(0, collective=mode)
        res.join_side()
        torch.cuda.synchronize(dev)
        vals[mode] = res.dls[0].clone()
    # rank-local gradient, gathered and summed on the host in rank order
    st = int(torch.cuda.current_stream(dev).cuda_stream)
    assert lib.dmel_forward_f32(*res.args_fwd(0, st)) == 0
    assert lib.dmel_backward_f32(*res.args_bwd(0, st)) == 0
    torch.cuda.synchronize(dev)
    local = [torch.zeros((), device=dev) for _ in range(cx.world)]
    dist.all_gather(local, res.dls[0].clone())
    host_sum = 0.0
    for v in local:
        host_sum += float(v)
    ref = float(vals["nccl"])
    for mode in ("fused", "side"):
        if mode in vals:
            out[f"rel_err_{mode}_vs_nccl"] = abs(float(vals[mode]) - ref) / max(abs(ref), 1e-30)
            g = [torch.zeros((), device=dev) for _ in range(cx.world)]
            dist.all_gather(g, vals[mode])
            out[f"{mode}_bit_identical_across_ranks"] = all(torch.equal(g[0], t) for t in g)
    out["rel_err_nccl_vs_host_sum"] = abs(ref - host_sum) / max(abs(host_sum), 1e-30)
    # N ranks on shards of one small batch == one rank on the whole batch (what tests/test_multi_gpu.py asserts)
    w = res.w
    Bs = 2 * cx.world
    g = torch.Generator().manual_seed(99)
    x = torch.randn(Bs, min(w["N"], 32000), generator=g)
    G = torch.randn(Bs, 1, w["M"], x.shape[1] // w["hop"] + 1, generator=g)
    layer = dm.MelSpectrogramLayer(torch.tensor(w["lam"]), n_mels=w["M"], n_points=x.shape[1], sample_rate=w["sr"],
                                   hop_length=w["hop"], device=str(dev), optimized=True, log_eps=LOG_EPS).to(dev)
    lo, hi = dm.shard_bounds(Bs, cx.world, cx.rank)
    xin = x[lo:hi].to(dev).requires_grad_(True)
    (layer(xin) * G[lo:hi].to(dev)).sum().backward()
    dm.allreduce_lambda_grad([layer.lambd], ring=cx.ring)
    sharded = float(layer.lambd.grad)
    layer.lambd.grad = None
    xall = x.to(dev).requires_grad_(True)
    (layer(xall) * G.to(dev)).sum().backward()
    whole = float(layer.lambd.grad)
    out["rel_err_sharded_vs_single_rank"] = abs(sharded - whole) / max(abs(whole), 1e-30)
    out["timed_out"] = bool(cx.ring.timed_out()) if cx.ring is not None else False
    return out


def measure_e2e(cx, res, steps, warmup):
    """Through the public module from pinned host memory: fp32 clips (the headline e2e) and int16 PCM clips that the
    kernels widen themselves."""
    import torch
    dev, dm, w = cx.dev, cx.dm, res.w
    B, N, sr, hop, M, lam = w["B"], w["N"], w["sr"], w["hop"], w["M"], w["lam"]
    layer = dm.MelSpectrogramLayer(torch.tensor(lam), n_mels=M, n_points=N, sample_rate=sr, hop_length=hop,
                                   device=str(dev), optimized=True, log_eps=LOG_EPS).to(dev)
    Rh = min(res.R, 3)
    hgen = torch.Generator().manual_seed(1234 + cx.rank)
    host_x = [torch.randn(B, N, generator=hgen).pin_memory() for _ in range(Rh)]
    host_out = torch.zeros(2).pin_memory()
    host_slot = [host_out[k] for k in range(2)]
    # three device buffers and two copy streams: uploads i and i+1 overlap each other and the compute of batch i-1
    NB = 3
    copy_streams = [torch.cuda.Stream(dev) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(NB)]
    free = [torch.cuda.Event() for _ in range(NB)]
    main_stream = torch.cuda.current_stream(dev)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def pipelined(k, upload, compute):
        for i in range(k + 2):
            if i < k:                      # upload clip batch i while batches i-1 / i-2 upload / compute
                cs = copy_streams[i % 2]
                with torch.cuda.stream(cs):
                    cs.wait_event(free[i % NB])
                    upload(i)
                    ready[i % NB].record(cs)
            j = i - 2
            if j >= 0:
                main_stream.wait_event(ready[j % NB])
                layer.lambd.grad = None
                compute(j)
                if cx.world > 1:
                    dm.allreduce_lambda_grad([layer.lambd], ring=cx.ring)
                free[j % NB].record(main_stream)
                # the step's result goes back to the host: dL/dlambda
                host_slot[j % 2].copy_(layer.lambd.grad, non_blocking=True)

    def run(upload, compute):
        for e in free:
            e.record(main_stream)
        pipelined(max(3, min(warmup, 10)), upload, compute)
        cx.barrier()
        ev0.record()
        pipelined(steps, upload, compute)
        ev1.record()
        cx.barrier()
        return ev0.elapsed_time(ev1)

    xdev = [torch.empty(B, N, device=dev) for _ in range(NB)]
    ms_f32 = run(lambda i: xdev[i % NB].copy_(host_x[i % Rh], non_blocking=True),
                 lambda j: layer(xdev[j % NB].detach().requires_grad_(True)).backward(res.Gs[j % res.R]))
    # the datasets are 16-bit PCM.  Same loop with pinned int16 host clips; the layer hands the int16
    # tensor to the kernels as it is (dmel.h DMEL_FLAG_X_I16: widened while the signal is staged — half the H2D bytes,
    # no cast pass, no fp32 copy of x).  x is data here (the reference's training case), so the step produces dL/dlambda.
    host_pcm = [(h * 8192.0).clamp_(-32768, 32767).to(torch.int16).pin_memory() for h in host_x]
    del xdev
    pdev = [torch.empty(B, N, device=dev, dtype=torch.int16) for _ in range(NB)]
    ms_pcm = run(lambda i: pdev[i % NB].copy_(host_pcm[i % Rh], non_blocking=True),
                 lambda j: layer(pdev[j % NB]).backward(res.Gs[j % res.R]))
    return ms_f32, ms_pcm


def lambda_only_times(cx, res, steps):
    """The reference's training case (x is data, only lambd trains): forward-mode kernel + dot-product backward vs
    forward + recomputing backward without dL/dx.  Eager launches, per GPU."""
    import torch
    lib, dev, w, geo = cx.lib, cx.dev, res.w, res.geo
    B, M, T = w["B"], w["M"], res.T
    dys = [torch.empty(B, 1, M, T, device=dev) for _ in range(min(res.R, 2))]
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def one(i, jvp):
        s = i % res.R
        st = int(torch.cuda.current_stream(dev).cuda_stream)
        if jvp:
            d = dys[i % len(dys)]
            a = res.args_fwd(s, st)
            rc = lib.dmel_forward_jvp_f32(*a[:13], d.data_ptr(), *a[13:])
            assert rc == 0, lib.dmel_status_string(rc)
            rc = lib.dmel_lambda_grad_f32(res.Gs[s].data_ptr(), d.data_ptr(), B * M * T, w["lam"], res.dls[s].data_ptr(),
                                          res.wss[s].data_ptr(), res.ws_bytes, st)
        else:
            lib.dmel_forward_f32(*res.args_fwd(s, st))
            rc = lib.dmel_backward_f32(*res.args_bwd(s, st, dx=False))
        assert rc == 0, lib.dmel_status_string(rc)

    out = {}
    for tag, jvp in (("jvp", True), ("recompute", False)):
        if jvp and not lib.dmel_jvp_supported(geo.n_fft):
            continue
        for i in range(3):
            one(i, jvp)
        torch.cuda.synchronize(dev)
        ev0.record()
        for i in range(steps):
            one(i, jvp)
        ev1.record()
        torch.cuda.synchronize(dev)
        out[tag] = ev0.elapsed_time(ev1) / steps
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--workload", default=None, help=f"one BASELINE shape only (default: {HEADLINE} + sub-results)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-graph", action="store_true", help="launch the resident loop eagerly instead of from a CUDA graph")
    ap.add_argument("--collective", default="side", choices=["side", "fused", "nccl"],
                    help="N > 1: peer-memory all-reduce on a side branch (default) / fused into finalize / NCCL")
    ap.add_argument("--nccl-only", action="store_true", help="same as --collective nccl")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the audio_mnist / esc50 sub-results of the default run")
    ap.add_argument("--cpu-budget", type=float, default=12.0, help="seconds of CPU work for cpu_baseline")
    args = ap.parse_args()
    if args.nccl_only:
        args.collective = "nccl"
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    wname = args.workload or HEADLINE
    w = workload(wname, world)
    audio_s = w["B"] * w["N"] / w["sr"]
    if args.steps is None:
        args.steps = default_steps(w)
    if args.warmup is None:
        args.warmup = max(3, min(50, args.steps // 10))
    args.warmup = max(3, args.warmup)
    if args.impl == "reference":
        return run_reference_arm(args, wname, w)

    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")

    # ---- CPU baseline first, on rank 0 alone: the process group does not exist yet, so the other ranks sit in its
    # TCP rendezvous (asleep, GPUs idle) instead of spinning in an NCCL barrier next to the timed CPU code
    cpu_line = None
    if rank == 0 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        rate, n, reps, _ = cpu_port_rate(w, args.cpu_budget, cores)
        cpu_line = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                    "sample": f"{n} of {w['B']} clips, best of {reps} passes, fwd+bwd (dx+dlambda), torch threads={cores}, "
                              "measured before the process group was created"}

    import torch.distributed as dist
    import dmel_b200 as dm
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    cx = Ctx()
    cx.rank, cx.world, cx.dev, cx.dm, cx.lib = rank, world, dev, dm, dm.load()
    cx.ring = None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
    cx.barrier = barrier

    if world > 1 and args.collective != "nccl":
        cx.ring = dm.PeerRing.create(dev)
        if rank == 0 and cx.ring is None:
            print("[bench] peer-memory all-reduce unavailable or failed its self-test: using NCCL", file=sys.stderr)
    collective = args.collective if (world > 1 and cx.ring is not None) else ("nccl" if world > 1 else "none")

    sampler = ClockSampler(local) if rank == 0 else None
    res = Resident(cx, wname, w, args.steps, collective)
    ms = res.timed(args.steps, args.warmup, not args.no_graph)
    step_ms = ms / args.steps
    per_step_launches = res.launches_per_step()
    kprof_steps = min(args.steps, 100)
    kern = res.kernel_times(kprof_steps)
    ms_lambda = lambda_only_times(cx, res, min(kprof_steps, 50))
    check = collective_check(cx, res) if world > 1 else None

    e2e = None
    if not args.no_e2e:
        e2e_steps = max(10, min(args.steps, int(2.0e6 / audio_s)))
        ms_f32, ms_pcm = measure_e2e(cx, res, e2e_steps, args.warmup)
        if world > 1:
            t = torch.tensor([ms_f32, ms_pcm], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_f32, ms_pcm = float(t[0]), float(t[1])
        e2e = (ms_f32, ms_pcm, e2e_steps)
    launch_mode = "cuda_graph" if res.graph is not None else "eager"
    l2_policy = f"inputs rotate over {res.R} buffer sets ({res.R * res.set_bytes / 2**20:.0f} MiB > 2x L2)"
    headline_roofline = res.roofline(kern, step_ms)
    del res
    torch.cuda.empty_cache()

    # ---- one-shot sub-results for the other BASELINE shapes (default run only)
    extras = {}
    if args.workload is None and not args.no_extras:
        for xname in ("audio_mnist", "esc50_split"):
            xw = workload(xname, world)
            xsteps = max(20, min(default_steps(xw), 400))
            r2 = Resident(cx, xname, xw, xsteps, collective)
            xms = r2.timed(xsteps, 5, not args.no_graph)
            xk = r2.kernel_times(min(xsteps, 50))
            x_audio = xw["B"] * xw["N"] / xw["sr"]
            extras[xname.replace("_split", "")] = {
                "value": world * x_audio * xsteps / (xms * 1e-3), "unit": UNIT, "ms_per_step": xms / xsteps,
                "steps": xsteps, "scaling": xw["scaling"], "per_gpu_batch": xw["B"], "n_fft": xw["n_fft"],
                "kernels_ms": {k: round(v[0], 6) for k, v in xk.items()},
                "roofline": r2.roofline(xk, xms / xsteps),
                "launch": "cuda_graph" if r2.graph is not None else "eager"}
            del r2
            torch.cuda.empty_cache()

    clocks = sampler.stop() if sampler else None

    if rank == 0:
        value = world * audio_s * args.steps / (ms * 1e-3)
        ms_lambda_only = ms_lambda.get("jvp", ms_lambda["recompute"])
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": w["scaling"], "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": config_dict(wname, w, extra={
                "l2_policy": l2_policy, "launch": launch_mode,
                "collective": {"none": "none (1 GPU)",
                               "side": "all-reduce of dL/dlambda (1 x fp32): peer-memory kernel over NVLink on a side branch "
                                       "of the step's graph (overlaps the next step's prep/forward)",
                               "fused": "all-reduce of dL/dlambda (1 x fp32) fused into the backward's finalize kernel over "
                                        "NVLink peer memory",
                               "nccl": "nccl all_reduce(dL/dlambda, 1 x fp32) per step"}[collective]}),
            "gpu_launches": per_step_launches * args.steps,
            "kernels_ms": {k: round(v[0], 6) for k, v in kern.items()},
            "lambda_only": {"ms_per_step": ms_lambda_only, "value": audio_s / (ms_lambda_only * 1e-3), "unit": UNIT,
                            "ms_per_step_by_method": {k: round(v, 6) for k, v in ms_lambda.items()},
                            "note": "per GPU, eager launches: the reference's training case (x without grad, only lambd "
                                    "trains, models.py:19 / main.py:27): forward-mode kernel + dot-product backward (jvp) "
                                    "vs forward + recomputing backward without dL/dx (recompute)"},
            "roofline": headline_roofline,
            "clocks": clocks,
        }
        if check is not None:
            line["collective_check"] = check
        if e2e:
            B, N = w["B"], w["N"]
            line["e2e"] = {"value": world * audio_s * e2e[2] / (e2e[0] * 1e-3), "unit": UNIT,
                           "h2d_bytes_per_step": 4 * B * N, "d2h_bytes_per_step": 4, "steps": e2e[2],
                           "api": "MelSpectrogramLayer.forward + Tensor.backward (eager, dL/dx and dL/dlambda), pinned host x, "
                                  "uploads pipelined over two copy streams; the 4 bytes read back are the step's dL/dlambda "
                                  "(y and dL/dx stay on the device, as in training)"}
            line["e2e_pcm16"] = {"value": world * audio_s * e2e[2] / (e2e[1] * 1e-3), "unit": UNIT,
                                 "h2d_bytes_per_step": 2 * B * N, "d2h_bytes_per_step": 4,
                                 "note": "same loop with int16 PCM host clips (the datasets' native format) handed to the "
                                         "layer as int16 and widened inside the kernels (DMEL_FLAG_X_I16); x is data, the "
                                         "step yields dL/dlambda (the reference's training case); extra, not the headline"}
        if extras:
            line["extra"] = {"workloads": extras}
        if cpu_line is not None:
            line["cpu_baseline"] = cpu_line
        print(json.dumps(line), flush=True)
    sys.stdout.flush()
    sys.stderr.flush()
    # ---- orderly teardown: nothing captured in a graph may outlive the objects it points into
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
        if cx.ring is not None:
            cx.ring.close()
            cx.ring = None
        torch.cuda.synchronize(dev)
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
