#!/usr/bin/env python
"""bench.py -- Chorowski TIMIT forward+backward(+gradient step) throughput on 1..8 B200.

Contract (driver):  python bench.py --gpus N --steps K --warmup W [--impl reference]
  N > 1 is launched by torch.distributed.run (one rank per GPU, NCCL).  Rank 0 prints ONE JSON line.

Workload ("step"): one pass of the training hot path over one synthetic minibatch of BASELINE.json
configs[1] ("Chorowski TIMIT baseline batch 32, location-aware attention + GRU encoder/decoder"):
timit/model_chorowski_baseline.lua with hybridAttendFeatureMaps = 16, filter 10 (`cfg2loc`), batch 32 PER GPU (weak
scaling), L = 300 log-mel frames of D = 123, T = 50 labels of V = 62:
zero grads -> encoder/decoder forward -> per-utterance NLL -> backward -> [sum of the flat gradient over the ranks:
the C ABI's NCCL plane, bucketed under the backward pass] -> /B, clip, adadelta, row-norm constraint (timit/timit.lua:233-348).
  value   : frames/s with inputs resident in HBM (CUDA events on the launching stream, per-step events,
            L2 flushed between steps by a 256 MiB write outside the event pairs; max over ranks)
  e2e     : the same metric through the public host API with HOST buffers (pinned): H2D of the
            batch and D2H of the per-utterance NLL inside the timed region, as a training loop's input
            pipeline (the next step's copies on a copy stream, the result read one step late)
  roofline: the kernel class with the largest share of the step, timed live with CUDA events by the
            library's profiling hook on an instrumented extra pass (s2s_ctx_profile)
  variants: the same contract (K steps, L2 flush, device events, max over ranks) for cfg2 (the shipped default, content-only
            attention K = 0), cfg3 (dropout model + AdaptiveWeightNoise) and, for N > 1, strong scaling at GLOBAL batch 32
  cpu_baseline: the CPU oracle (a C restatement of the reference; Torch7 cannot run here) on the same 32-utterance
            minibatch, all host threads
`--config cfg2` / `cfg3` / `cfg4` run one of the other training configurations of BASELINE.json as the headline (content-only;
dropout + AdaptiveWeightNoise; librispeech/model_vgg.lua with its VGG front-end); `--config cfg5` is the attention-step sweep.
`--impl reference` times that CPU oracle alone (the reference's own implementation is Lua/Torch7 and
cannot be installed in this image: see DESIGN.md).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG = dict(D=123, H=256, NL=3, S=512, ST=256, V=62, K=16, KF=10, M=64, MW=7)   # headline: location-aware (hybridAttendFeatureMaps = 16, filter 10)
B_PER_GPU, L, T = 32, 300, 50
METRIC = "chorowski_timit_fwd_bwd_frames_per_sec"
WORKLOAD_CFG3 = ("cfg3: timit/model_chorowski_baseline_dropout.lua (cfg2loc model + Dropout(0.5) on {s,c}) with AdaptiveWeightNoise "
                 "(lambda=1, sigma_init=0.075), batch 32/GPU, L=300, T=50; step = AWN sample (one per shard) + dropout mask + zero-grad + "
                 "fwd + NLL + bwd + [all-reduce] + /B + clip + AWN forward/backward + adadelta over {mu, log sigma^2} + row-norm")
WORKLOAD = ("cfg2loc: timit/model_chorowski_baseline.lua (3x biGRU-256 encoder, location-aware attention K=16 k=10, GRU-256 decoder, "
            "maxout 64x7), batch 32/GPU, L=300, D=123, T=50, V=62; step = zero-grad + fwd + NLL + bwd + "
            "[all-reduce] + /B + clip + adadelta + row-norm")


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), float(p["bf16_tflops"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, 1590.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(cls):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the kernel class, from the committed `ncu --set full`
    capture (profiles/r02_ncu_full.json; round-1 capture for kernels not re-captured); None when there is no capture for it."""
    names = {"attn_fwd": ("attn_fwd_kernel",), "attn_bwd": ("attn_bwd_kernel",), "attn_dvh": ("attn_dvh_kernel",),
             "gru_fwd": ("gru3_fwd_kernel", "gru_seq_fwd"), "gru_bwd": ("gru3_bwd_kernel", "gru_seq_bwd"), "gemm": ("gemm_tc_kernel",),
             "gemm_side": ("gemm_tc_kernel",), "dense_small": ("dense_small_kernel",), "dec_fwd": ("dec_cluster_fwd_kernel",),
             "dec_bwd": ("dec_cluster_bwd_kernel",)}
    try:
        rows = []
        for f in ("r02_ncu_full.json", "r01_ncu_full.json"):
            path = os.path.join(ROOT, "profiles", f)
            if not rows and os.path.exists(path):
                rows = [r for r in json.load(open(path)) if r["kernel"].startswith(names[cls])]
        unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        vals = [r["dram__bytes_read.sum"] * unit[r["dram__bytes_read.sum.unit"]] + r["dram__bytes_write.sum"] * unit[r["dram__bytes_write.sum.unit"]]
                for r in rows]
        return sum(vals) / len(vals) if vals else None
    except Exception:
        return None


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def synth(seed, B):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((B, L, CFG["D"])).astype(np.float32)
    labels = rng.integers(0, CFG["V"] - 1, (B, T)).astype(np.int32)
    labels[:, -1] = CFG["V"] - 1
    lengths = np.full(B, L, np.int32)
    tlens = np.full(B, T, np.int32)
    return X, labels, lengths, tlens


# -------------------------------------------------------------------------------------------------
def cpu_oracle_run(nutt, nthreads, reps=1):
    """frames/s of the CPU oracle (fp32, OpenMP over utterances) on `nutt` utterances of the workload"""
    from oracle.oracle import Oracle, build, init_params
    build()
    o = Oracle("f32")
    P = init_params(CFG, seed=1234, dtype=np.float32, oracle=Oracle("f64"))
    X, labels, lengths, tlens = synth(99, nutt)
    tot = 0.0
    for _ in range(reps):
        t0 = time.perf_counter()
        o.model_fwdbwd(CFG, P, X, lengths, labels, tlens, nthreads=nthreads, want=())
        tot += time.perf_counter() - t0
    return nutt * L / (tot / reps), tot / reps


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_threads()
    nthreads = max(1, min(cores, 32))
    nutt = B_PER_GPU  # the GPU arm's step: the same 32-utterance minibatch (OpenMP over utterances)
    for _ in range(min(args.warmup, 1)):
        cpu_oracle_run(nutt, nthreads)
    times = []
    for _ in range(args.steps):
        _, dt = cpu_oracle_run(nutt, nthreads)
        times.append(dt)
    ms = 1e3 * float(np.mean(times))
    val = nutt * L / (ms / 1e3)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "global_batch": B_PER_GPU, "parallelism": "cpu",
                   "reference_arm": "CPU oracle (oracle/s2s_oracle.c, C restatement of the Torch7 path; "
                   "the Lua reference cannot be installed: no LuaJIT/Torch7 in the image)"},
        "cpu_baseline": {"value": val, "unit": "frames/s", "cores": nthreads, "kind": "port",
                         "sample": f"{nutt} utterances (L={L}, T={T}) per step, fwd+bwd, OpenMP over utterances"},
        "e2e": {"value": val, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit_json(json.dumps(line))


# -------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, dev):
        self.dev = dev
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.dev)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# -------------------------------------------------------------------------------------------------
# BASELINE.json configs[3]: librispeech/model_vgg.lua -- VGG front-end + attention decoder (two-stage Maxout MLP), long
# utterances.  Composed from the C-ABI pieces: s2s_vgg_forward -> s2s_attention_forward -> s2s_nll_grad_seed ->
# s2s_attention_backward -> s2s_vgg_backward, then the gradient step.  Builder-chosen where BASELINE.json leaves it open
# (SURVEY 8d): X [B,3,1600,40], L = 796, V = 29, T = 250 labels, batch 8 per GPU.
def run_cfg4(args):
    import torch
    import torch.distributed as dist
    import s2s_b200 as s2s

    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    ctx = s2s.Context(local)
    if world > 1:
        s2s.dp.init(ctx, rank, world)             # the C ABI's NCCL plane (csrc/dp_nccl.cu)
    B, Tin, F, Tdec, V = 8, 1600, 40, 250, 29
    vcfg = s2s.VGG_LIBRISPEECH
    dcfg = dict(D=F, H=vcfg["OUT"] // 2, NL=0, S=512, ST=256, V=V, K=0, KF=10, M=64, MW=7, MLP=2)   # model_vgg.lua:57-69
    Lenc = (Tin - 8) // 2
    rng = np.random.default_rng(4000 + rank)
    ne = s2s.vgg_param_count(vcfg, F); nd = s2s.param_count(dcfg)
    prng = np.random.default_rng(1234)
    P = torch.cat([torch.from_numpy((prng.uniform(-1, 1, ne) * 0.03).astype(np.float32)), torch.from_numpy(s2s.init_params(dcfg, seed=1234))]).to(dev)
    G = torch.zeros_like(P); vs = torch.zeros_like(P); as_ = torch.zeros_like(P)
    Pe, Pd, Ge, Gd = P[:ne], P[ne:], G[:ne], G[ne:]
    Xh = rng.standard_normal((B, 3, Tin, F)).astype(np.float32)
    yh = rng.integers(0, V - 1, (B, Tdec)).astype(np.int32); yh[:, -1] = V - 1
    X = torch.from_numpy(Xh).to(dev); y = torch.from_numpy(yh).to(dev)
    nll = torch.zeros(B, device=dev); dlogp = torch.zeros(B, Tdec, V, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def fwdbwd(Xd, yd):
        h = s2s.vgg_forward(ctx, vcfg, Pe, Xd)
        logp = s2s.attention_forward(ctx, dcfg, Pd, h, yd)
        s2s.nll_grad_seed(ctx, logp, yd, flags=s2s.NORMALIZE_NLL, nll=nll, dlogp=dlogp)
        dh = s2s.attention_backward(ctx, dcfg, Pd, Gd, h, yd, dlogp)
        s2s.vgg_backward(ctx, vcfg, Pe, Xd, dh, dP=Ge)
        return h, logp, dh                       # kept alive: a captured graph writes to these buffers

    graph = {"id": None, "keep": None}

    def step(Xd, yd):
        G.zero_()
        if graph["id"] is not None and Xd is X and yd is y:
            ctx.graph_launch(graph["id"])        # the whole forward + backward as one replayed CUDA graph
        else:
            fwdbwd(Xd, yd)
        if world > 1:
            s2s.dp.allreduce(ctx, G)
        s2s.grad_finalize(ctx, G, P, B * world, 1e20, want_norm=False)
        s2s.adadelta(ctx, P, G, vs, as_)
        s2s.model_rownorm_constraint(ctx, dcfg, Pd, 1.0)
        return nll

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    step(X, y)                                   # eager once: sizes every workspace
    torch.cuda.synchronize()
    if not os.environ.get("S2S_BENCH_NOGRAPH"):
        ctx.graph_begin()
        graph["keep"] = fwdbwd(X, y)
        graph["id"] = ctx.graph_end()
    for _ in range(max(args.warmup, 3)):
        step(X, y)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    l0 = ctx.launches
    barrier()
    for a, b in ev:
        flush.fill_(1); a.record(); step(X, y); b.record()
    barrier()
    launches = ctx.launches - l0
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([sum(a.elapsed_time(b) for a, b in ev) / args.steps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    frames = B * Tin * world
    # e2e: pinned host inputs -> H2D -> step -> D2H(nll)
    Xp = torch.from_numpy(Xh).pin_memory(); yp = torch.from_numpy(yh).pin_memory()
    nll_h = torch.empty(B).pin_memory()

    def e2e_step():
        X.copy_(Xp, non_blocking=True); y.copy_(yp, non_blocking=True)      # the step's device input buffers
        step(X, y)
        nll_h.copy_(nll, non_blocking=True)
        torch.cuda.synchronize()

    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    t = torch.tensor([(time.perf_counter() - t0) * 1e3 / args.steps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item())
    # per-class breakdown of one extra pass
    hbm, tf, how = peaks()
    classes = {}
    if graph["id"] is not None:
        ctx.graph_destroy(graph["id"]); graph["id"] = None           # the instrumented pass runs eagerly
    if rank == 0:
        ctx.profile(True)
    ev2 = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
    torch.cuda._sleep(int(60e6)); ev2[0].record(); step(X, y); ev2[1].record(); torch.cuda.synchronize()
    total_ms = ev2[0].elapsed_time(ev2[1])
    roof = None
    if rank == 0:
        prof = ctx.profile_read(); ctx.profile(False)
        for k, (kms, cnt, work) in prof.items():
            if cnt:
                tensor = k in ("gemm", "gemm_side")
                ach = work / (kms * 1e-3)
                classes[k] = {"bound": "tensor" if tensor else "hbm", "ms_per_step": kms, "launches_per_step": cnt,
                              "achieved": ach / (1e12 if tensor else 1e9), "peak": tf if tensor else hbm, "unit": "TFLOP/s" if tensor else "GB/s",
                              "frac": ach / (1e12 if tensor else 1e9) / (tf if tensor else hbm), "share": kms / total_ms}
        top = max((k for k in classes if k != "gemm_side"), key=lambda k: classes[k]["ms_per_step"])     # largest class ON the critical path
        c = classes[top]
        roof = {"kernel": top, "bound": c["bound"], "achieved": c["achieved"], "peak": c["peak"], "unit": c["unit"], "frac": c["frac"], "traffic": None,
                "peak_source": how, "share_of_step": c["share"], "instrumented_step_ms": total_ms,
                "note": "fp32-equivalent FLOPs of the 3xTF32 tcgen05 GEMMs (3 tensor-core passes per product) against the measured bf16 peak; "
                        "the unfold/fold/ReLU/pooling kernels of the first VGG path are not in a profiled class"}
        line = {"metric": "librispeech_vgg_fwd_bwd_frames_per_sec", "value": frames / (ms / 1e3), "unit": "frames/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic",
                "config": {"workload": "cfg4: librispeech/model_vgg.lua (4 conv3x3 + 2 poolings + 4 1x1 layers -> 796 annotations of 512; content attention, "
                                       "GRU-256 decoder, Maxout-Linear-Maxout-Linear MLP), X [8,3,1600,40] per GPU, 250 labels of 29 classes; step = zero-grad + "
                                       "fwd + NLL + bwd + [all-reduce] + /B + clip + adadelta + row-norm (decoder)",
                           "global_batch": B * world, "parallelism": f"dp{world}", "l2": "flushed between timed steps (256 MiB write)"},
                "e2e": {"value": frames / (e2e_ms / 1e3), "unit": "frames/s", "ms_per_step": e2e_ms,
                        "h2d_bytes_per_step": Xp.numel() * 4 + yp.numel() * 4, "d2h_bytes_per_step": B * 4},
                "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "kernels": classes,
                "cpu_baseline": None}
        emit_json(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# -------------------------------------------------------------------------------------------------
# BASELINE.json configs[4] and the metric's second half ("attention HBM GB/s"): the attention-step sweep over encoder length and
# batch (scoring + softmax + context, forward and backward), L2 flushed between launches, content (K = 0) and location-aware (k = 10)
def run_cfg5(args):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    sys.path.insert(0, os.path.join(ROOT, "benchmarks"))
    import attn_sweep
    import s2s_b200 as s2s
    ctx = s2s.Context(int(os.environ.get("LOCAL_RANK", "0")))
    sampler = ClockSampler(int(os.environ.get("LOCAL_RANK", "0")))
    sampler.start()
    l0 = ctx.launches
    res = {kf: attn_sweep.run_sweep(kf, quick=False, verbose=False, ctx=ctx) for kf in (0, 10)}
    launches = ctx.launches - l0
    clocks = sampler.stop()
    hbm, tf, how = peaks()

    def best(rows, key):
        r = max(rows, key=lambda r: r[key])
        return {"B": r["B"], "L": r["L"], "GB/s": r[key], "frac": r[key] / hbm}
    loc, con = res[10]["rows"], res[0]["rows"]
    top = max(loc, key=lambda r: (r["mbytes"] * 2) / (r["fwd_us"] + r["bwd_us"]))
    value = 1e-3 * (4.0 * top["B"] * (2 * top["L"] * 1024 + 8 * top["L"] + 3 * 1024)) / ((top["fwd_us"] + top["bwd_us"]) * 1e-6) / 1e6
    line = {"metric": "attention_step_hbm_gbs", "value": value, "unit": "GB/s", "n_gpus": 1, "steps": 10, "warmup": 3,
            "ms_per_step": (top["fwd_us"] + top["bwd_us"]) * 1e-3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": "cfg5: attention-step sweep (scoring + softmax + context, forward + backward), S = A = 512, L in {100..2000} x B in {1..256}, "
                                   "location-aware k = 10 (headline: best forward+backward point) and content-only; algorithmic bytes A_f + A_b,min (SURVEY 8d)",
                       "l2": "flushed before every timed launch (256 MiB write)", "headline_point": {"B": top["B"], "L": top["L"]}},
            "roofline": {"bound": "hbm", "achieved": value, "peak": hbm, "unit": "GB/s", "frac": value / hbm, "traffic": None, "peak_source": how},
            "best": {"location_fwd": best(loc, "fwd_gbs"), "location_bwd": best(loc, "bwd_gbs"), "content_fwd": best(con, "fwd_gbs"), "content_bwd": best(con, "bwd_gbs")},
            "points_at_or_above_60pct": {"location_fwd": sum(r["fwd_frac"] >= 0.6 for r in loc), "location_bwd": sum(r["bwd_frac"] >= 0.6 for r in loc),
                                         "content_fwd": sum(r["fwd_frac"] >= 0.6 for r in con), "content_bwd": sum(r["bwd_frac"] >= 0.6 for r in con), "of": len(loc)},
            "sweep": {"location_k10": loc, "content": con}, "gpu_launches": int(launches), "clocks": clocks, "cpu_baseline": None,
            "e2e": None}
    emit_json(json.dumps(line))


def dbg(msg):
    if os.environ.get("S2S_BENCH_DEBUG"):
        print(f"[bench rank {os.environ.get('RANK', '0')}] {msg}", file=sys.stderr, flush=True)


def run_ours(args):
    if os.environ.get("S2S_BENCH_DEBUG"):
        import faulthandler
        faulthandler.dump_traceback_later(int(os.environ["S2S_BENCH_DEBUG"]), exit=True)
    import torch
    import torch.distributed as dist
    import s2s_b200 as s2s

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dbg("process group up")
    ctx = s2s.Context(local)
    # the data path's collective is the C ABI's own NCCL plane (csrc/dp_nccl.cu); torch.distributed only ships the 128-byte id and carries
    # the barrier / max-over-ranks of the timing.  Default: the bucketed overlap inside s2s_model_fwdbwd -- the gradient buckets are reduced
    # on the low-priority side stream under the remaining backward pass, as nodes of the replayed CUDA graph (N=8: 8.15 vs
    # 8.20 ms per step).  S2S_BENCH_DP_OVERLAP=0: one s2s_dp_allreduce of the flat gradient after the graph.
    dp_overlap = os.environ.get("S2S_BENCH_DP_OVERLAP", "1") not in ("0", "")
    if world > 1:
        s2s.dp.init(ctx, rank, world, overlap=dp_overlap)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    class Workload:
        """one configuration: parameters, optimiser state, a resident synthetic batch and its step function"""

        def __init__(self, cfg, B, awn_dropout=False, seed=1000):
            self.cfg, self.B, self.awn = dict(cfg), B, awn_dropout
            n = s2s.param_count(cfg)
            self.n = n
            self.P = torch.from_numpy(s2s.init_params(cfg, seed=1234)).to(dev)
            self.G = torch.zeros(n, device=dev)
            self.v = torch.zeros(n, device=dev); self.a = torch.zeros(n, device=dev)
            self.host = synth(seed + rank, B)
            Xh, yh, lh, th = self.host
            self.X = torch.from_numpy(Xh).to(dev); self.y = torch.from_numpy(yh).to(dev)
            self.ln = torch.from_numpy(lh).to(dev); self.tl = torch.from_numpy(th).to(dev)
            self.nll = torch.zeros(B, device=dev)
            if awn_dropout:
                # AdaptiveWeightNoise.lua:8-25: weight = {mu, s = log sigma^2}, s initialised to log(sigma_init^2) (timit.lua:35-36,198-205)
                self.W2 = torch.cat([self.P, torch.full((n,), float(np.log(0.075 ** 2)), device=dev)])
                self.gW2 = torch.zeros(2 * n, device=dev)
                self.v2 = torch.zeros(2 * n, device=dev); self.a2 = torch.zeros(2 * n, device=dev)
                self.Pn = torch.empty(n, device=dev)
                self.mask = torch.empty(B, T, cfg["ST"] + 2 * cfg["H"], device=dev)
                self.counter = 0

        def step(self, Xd=None, yd=None, lnd=None, tld=None):
            Xd = self.X if Xd is None else Xd; yd = self.y if yd is None else yd
            lnd = self.ln if lnd is None else lnd; tld = self.tl if tld is None else tld
            cfg, G, Bg = self.cfg, self.G, self.B * world
            if self.awn:
                self.counter += 1
                seed = (self.counter << 8) | rank                                  # a different sample per rank and step
                s2s.awn_sample(ctx, self.W2, seed=seed, out=self.Pn)               # parameters:copy(AWN:Sample())   (timit.lua:247-253)
                s2s.dropout_mask(ctx, self.mask.shape, 0.5, seed=seed, out=self.mask)   # nn.Dropout on {s,c}  (model_chorowski_baseline_dropout.lua:56)
                G.zero_()
                s2s.model_fwdbwd(ctx, cfg, self.Pn, G, Xd, yd, lengths=lnd, tlens=tld, dropmask=self.mask, flags=s2s.NORMALIZE_NLL, nll=self.nll)
                if world > 1 and not dp_overlap:
                    s2s.dp.allreduce(ctx, G)
                s2s.grad_finalize(ctx, G, self.Pn, Bg, 1e20, want_norm=False)
                s2s.awn_accgrad(ctx, self.W2, G, 1.0, out=self.gW2)                # AWN:backward(nll, gradients)    (timit.lua:318-327)
                s2s.adadelta(ctx, self.W2, self.gW2, self.v2, self.a2)             # optimMethod(optimfunc, adaparameters, ...)  (:336)
                s2s.model_rownorm_constraint(ctx, cfg, self.W2[:self.n], 1.0)      # the graph's weights are views of mu here (:346-348)
                return self.nll
            G.zero_()                                                             # zeroGradParameters (timit.lua:233)
            s2s.model_fwdbwd(ctx, cfg, self.P, G, Xd, yd, lengths=lnd, tlens=tld, flags=s2s.NORMALIZE_NLL, nll=self.nll)
            if world > 1 and not dp_overlap:                                      # (with overlap the library reduced the buckets under the backward pass)
                s2s.dp.allreduce(ctx, G)                                          # data-parallel gradient sum over NVLink: s2s_dp_allreduce
            s2s.dp.gradient_step(ctx, s2s, cfg, self.P, G, self.v, self.a, Bg)    # /B, clip, adadelta, row-norm (timit.lua:292-348)
            return self.nll

        def timed(self, steps, warmup):
            """K steps, per-step CUDA events on the launching stream, L2 flushed between steps; max over ranks"""
            for i in range(max(warmup, 3)):
                self.step()
                torch.cuda.synchronize()
            barrier()
            ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
            l0 = ctx.launches
            barrier()
            for a, b in ev:
                flush.fill_(1)
                a.record()
                self.step()
                b.record()
            barrier()
            launches = ctx.launches - l0
            ms = max_over_ranks(sum(a.elapsed_time(b) for a, b in ev) / steps)
            return ms, launches

    K16 = dict(CFG, K=16)
    K0 = dict(CFG, K=0)
    headline_cfg = dict(CFG)
    cfg3 = args.config == "cfg3"
    main_w = Workload(headline_cfg, B_PER_GPU, awn_dropout=cfg3)
    B = B_PER_GPU
    dbg("workload built")

    # ---- timed region of the headline configuration ------------------------------------------------------------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms, launches = main_w.timed(args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    dbg("timed region done")
    frames = B * L * world
    value = frames / (ms / 1e3)

    # ---- e2e: host buffers -> H2D -> step -> D2H(nll), wall clock with device sync, max over ranks ------
    Xh, yh, lh, th = main_w.host
    X, y, ln, tl, nll = main_w.X, main_w.y, main_w.ln, main_w.tl, main_w.nll
    Xp = torch.from_numpy(Xh).pin_memory(); yp = torch.from_numpy(yh).pin_memory()
    lp = torch.from_numpy(lh).pin_memory(); tp = torch.from_numpy(th).pin_memory()
    Xd = torch.empty_like(X); yd = torch.empty_like(y); lnd = torch.empty_like(ln); tld = torch.empty_like(tl)
    # Input pipeline of a training loop: step i+1's host buffers are copied H2D on a copy stream into a staging set while
    # step i computes (the step then moves them into its fixed input buffers with a device copy), and step i's result is
    # read on the host while step i+1 runs.  Every step still pays its own H2D and D2H inside the timed region.
    cs = torch.cuda.Stream(device=dev)
    cur = torch.cuda.current_stream(dev)
    stg = [tuple(torch.empty_like(a) for a in (X, y, ln, tl)) for _ in range(2)]
    nll_hs = [torch.empty(B, dtype=torch.float32).pin_memory() for _ in range(2)]
    ev_in = [torch.cuda.Event() for _ in range(2)]; ev_free = [torch.cuda.Event() for _ in range(2)]; ev_out = [torch.cuda.Event() for _ in range(2)]

    def prefetch(i):
        k = i & 1
        with torch.cuda.stream(cs):
            cs.wait_event(ev_free[k])
            for dst, src in zip(stg[k], (Xp, yp, lp, tp)):
                dst.copy_(src, non_blocking=True)
            ev_in[k].record(cs)

    def e2e_run(n):
        tot = 0.0
        for k in range(2):
            ev_free[k].record(cur)
        prefetch(0)
        for i in range(n):
            k = i & 1
            if i + 1 < n:
                prefetch(i + 1)
            cur.wait_event(ev_in[k])
            for dst, src in zip((Xd, yd, lnd, tld), stg[k]):
                dst.copy_(src, non_blocking=True)
            ev_free[k].record(cur)
            main_w.step(Xd, yd, lnd, tld)
            nll_hs[k].copy_(nll, non_blocking=True)
            ev_out[k].record(cur)
            if i > 0:
                ev_out[k ^ 1].synchronize()
                tot += float(nll_hs[k ^ 1].sum())
        ev_out[(n - 1) & 1].synchronize()
        tot += float(nll_hs[(n - 1) & 1].sum())
        torch.cuda.synchronize()
        return tot

    e2e_run(3)
    barrier()
    t0 = time.perf_counter()
    e2e_run(args.steps)
    barrier()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3 / args.steps)
    h2d = Xp.numel() * 4 + yp.numel() * 4 + lp.numel() * 4 + tp.numel() * 4
    d2h = nll_hs[0].numel() * 4

    # ---- roofline: instrumented extra pass (graphs off), per-kernel-class CUDA events ------------------
    roof = None
    classes = {}
    hbm, tf, how = peaks()
    ctx.set_graphs(False)
    if rank == 0:
        ctx.profile(True)
    ev2 = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
    nprof = 3
    # Eager launches of microsecond kernels are host-bound: an event pair around such a launch would also time the
    # host gap.  A spin kernel keeps the GPU behind the host while the step is enqueued, so every (event, kernel,
    # event) triple executes back to back and the events bracket GPU execution only.  Every rank runs the pass
    # (the step contains the all-reduce); only rank 0 records.
    total_ms = 0.0
    for _ in range(nprof):
        torch.cuda._sleep(int(60e6))          # ~30 ms of spinning, not inside any event pair
        ev2[0].record()
        main_w.step()
        ev2[1].record()
        torch.cuda.synchronize()
        total_ms += ev2[0].elapsed_time(ev2[1]) / nprof
    ctx.set_graphs(True)
    if rank == 0:
        prof = ctx.profile_read()
        ctx.profile(False)
        for k, (kms, cnt, work) in prof.items():
            if cnt == 0:
                continue
            per_launch_ms = kms / cnt
            ach = work / cnt / (per_launch_ms * 1e-3)
            if k in ("gemm", "gemm_side"):
                classes[k] = {"bound": "tensor", "ms_per_step": kms / nprof, "launches_per_step": cnt / nprof, "achieved": ach / 1e12,
                              "peak": tf, "unit": "TFLOP/s", "frac": ach / 1e12 / tf, "share": kms / nprof / total_ms,
                              "note": "fp32-equivalent FLOPs; large products run 3xTF32 on tcgen05 (3 tensor-core passes per product), small ones exact-fp32 SIMT"
                                      + ("; issued on the low-priority side stream with a grid limited to the 36 SMs the cluster kernels leave idle: "
                                         "overlapped with the recurrences, NOT on the critical path of the step" if k == "gemm_side" else
                                         "; the products on the critical path of the step (forward projections, data gradients, layer 0's weight gradients)")}
            else:
                classes[k] = {"bound": "hbm", "ms_per_step": kms / nprof, "launches_per_step": cnt / nprof, "achieved": ach / 1e9,
                              "peak": hbm, "unit": "GB/s", "frac": ach / 1e9 / hbm, "share": kms / nprof / total_ms}
        top = max((k for k in classes if k != "gemm_side"), key=lambda k: classes[k]["ms_per_step"])     # largest class ON the critical path
        c = classes[top]
        roof = {"kernel": top, "bound": c["bound"], "achieved": c["achieved"], "peak": c["peak"], "unit": c["unit"], "frac": c["frac"],
                "traffic": ncu_traffic(top), "traffic_source": "profiles/ ncu --set full capture (bytes per launch)",
                "algorithmic_bytes_per_launch": prof[top][2] / max(prof[top][1], 1),
                "peak_source": how, "share_of_step": c["share"], "instrumented_step_ms": total_ms,
                "note": "bound/achieved/peak are the HBM view the contract asks for (algorithmic bytes / launch time); the recurrence kernels are "
                        "bound by the fp32 FMA issue rate and the DSMEM exchange chain, not by HBM: see kernels.gru_*.fp32_fma "
                        "(DESIGN.md 3.5, profiles/r02_gru_trace.txt); attention kernels: see `kernels` and `--config cfg5`"}
        # the recurrences' own roofline: FMAs of the three H x H mat-vecs per frame-step against the issue rate of 3-register FFMA
        # (one warp instruction per 2 cycles per SM sub-partition = 64 FMA/clk/SM; B300_MICROARCH.md) on the 112 SMs the one-wave
        # cluster launch occupies (14 clusters of 8)
        mhz = (clocks or {}).get("sm_mhz") or 1965.0
        for k in ("gru_fwd", "gru_bwd"):
            if k in classes and classes[k]["launches_per_step"]:
                fma = float(B_PER_GPU) * L * 2 * 3 * CFG["H"] * CFG["H"]                         # per launch (one layer, both directions)
                sec = classes[k]["ms_per_step"] * 1e-3 / classes[k]["launches_per_step"]
                per_clk_sm = fma / sec / (112 * mhz * 1e6)
                classes[k]["fp32_fma"] = {"fma_per_launch": fma, "sms": 112, "sm_mhz": mhz, "achieved_fma_per_clk_per_sm": per_clk_sm,
                                          "issue_limit_fma_per_clk_per_sm": 64.0, "frac": per_clk_sm / 64.0,
                                          "us_per_frame_step": sec * 1e6 / L}
    kc = ctx.kernel_counts()

    # ---- the other configurations of BASELINE.json, device-timed in the same run (same contract: K steps, L2 flush, max over ranks) ----
    variants = {}
    if not args.no_variants and args.config == "cfg2loc":
        vsteps = max(3, min(args.steps, 10))
        del main_w
        specs = [("cfg2", K0, B_PER_GPU, False, "the shipped default hybridAttendFeatureMaps = 0 (content-only attention), batch 32/GPU"),
                 ("cfg3", K16, B_PER_GPU, True, "model_chorowski_baseline_dropout.lua + AdaptiveWeightNoise (lambda 1, sigma_init 0.075), batch 32/GPU")]
        if world > 1 and B_PER_GPU % world == 0:
            specs.append(("cfg2loc_strong_b32", K16, B_PER_GPU // world, False, f"strong scaling: GLOBAL batch 32 = {B_PER_GPU // world}/GPU"))
        for name, vcfg, vB, vawn, note in specs:
            w = Workload(vcfg, vB, awn_dropout=vawn)
            vms, vl = w.timed(vsteps, 3)
            variants[name] = {"ms_per_step": vms, "value": vB * L * world / (vms / 1e3), "unit": "frames/s", "steps": vsteps,
                              "global_batch": vB * world, "scaling": "strong" if "strong" in name else "weak", "gpu_launches": int(vl), "note": note}
            del w

    # ---- CPU baseline (rank 0, N = 1 only): bounded sample of the same workload -------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = host_threads()
        nthreads = max(1, min(cores, 32))
        nutt = 32
        reps = 4 if nthreads >= 16 else 2                             # ~10-20 s of CPU work
        val, dt = cpu_oracle_run(nutt, nthreads, reps=reps)
        cpu = {"value": val, "unit": "frames/s", "cores": nthreads, "kind": "port",
               "sample": f"{reps} passes of fwd+bwd over the same {nutt}-utterance minibatch (L={L}, T={T}), fp32 C oracle, OpenMP over utterances, {dt:.1f} s per pass (mean)"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD_CFG3 if cfg3 else WORKLOAD, "global_batch": B * world, "parallelism": f"dp{world}",
                       "l2": "flushed between timed steps (256 MiB write outside the per-step event pairs)",
                       "e2e_pipeline": "step i+1's H2D on a copy stream during step i, result of step i read during step i+1",
                       "collective": ("none (1 rank)" if world == 1 else
                                      ("s2s_dp_allreduce buckets inside s2s_model_fwdbwd (NCCL via the C ABI, side stream, under the backward pass)"
                                       if dp_overlap else "s2s_dp_allreduce after s2s_model_fwdbwd (NCCL via the C ABI)"))},
            "e2e": {"value": frames / (e2e_ms / 1e3), "unit": "frames/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches),
            "kernel_launch_counts": kc,
            "clocks": clocks,
            "roofline": roof,
            "kernels": classes,
            "variants": variants,
            "cpu_baseline": cpu,
        }
        emit_json(json.dumps(line))
    if world > 1:
        s2s.dp.destroy(ctx)
        dist.destroy_process_group()


def main():
    # stdout carries exactly ONE line (the JSON); anything libraries print there (e.g. NCCL's version banner) goes to stderr
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        os.write(real_stdout, (line + "\n").encode())

    globals()["emit_json"] = emit
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-variants", action="store_true", help="skip the extra configurations reported under `variants`")
    ap.add_argument("--config", default="cfg2loc", choices=["cfg2", "cfg2loc", "cfg3", "cfg4", "cfg5"],
                    help="cfg2loc = the metric's configuration (default): BASELINE.json configs[1], location-aware attention (hybridAttendFeatureMaps = 16 "
                         "as timit/timit.lua:130, filter 10 as model_chorowski_baseline.lua:39); its line also carries cfg2 / cfg3 / strong scaling under "
                         "`variants`.  cfg2 = the shipped default hybridAttendFeatureMaps = 0 (content-only); cfg3 = dropout model + AdaptiveWeightNoise "
                         "(configs[2]); cfg4 = librispeech/model_vgg.lua, VGG front-end + attention decoder (configs[3]); cfg5 = attention-step sweep (configs[4])")
    args = ap.parse_args()
    if args.config == "cfg2":      # model.hybridAttendFeatureMaps = opt.hybridAttendFeatureMaps or 0 (model_chorowski_baseline.lua:40)
        CFG["K"] = 0
        globals()["WORKLOAD"] = WORKLOAD.replace("cfg2loc:", "cfg2:").replace("location-aware attention K=16 k=10", "content attention K=0")
    if args.impl == "reference":
        run_reference(args)
    elif args.config == "cfg4":
        run_cfg4(args)
    elif args.config == "cfg5":
        run_cfg5(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
