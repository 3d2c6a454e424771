"""bench.py -- headline benchmark of the B200-native 2048-PPO hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one pass of g2048_step over config C2 of BASELINE.json: 2^20 random boards x 4
moves = 4,194,304 full Game2048.step transitions (move, merge points, shaping record, Philox
spawn, legal mask / done) per GPU.  Prints ONE JSON line (see DESIGN.md "Measurement").

  value     env-steps/s, inputs resident in HBM, CUDA-event timed on the launching stream
  e2e       same metric through the public host API (g2048.env.step) with HOST pinned buffers,
            host<->device copies inside the timed region
  roofline  algorithmic bytes (30 B / transition) / measured kernel time vs the measured HBM peak
  cpu_baseline  the C port of the reference (oracle/) on the host cores, bounded sample

`--impl reference` times the reference's CPU algorithm (the oracle port; the reference itself is
pure Python and does not travel to the GPU box) with all host threads on the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "2048-ppo_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

N_BOARDS = 1 << 20
MOVES = 4
N_TRANS = N_BOARDS * MOVES
BYTES_PER_TRANSITION = 30      # g2048_step on (board, action) pairs: 8 board in + 1 action + 8 board out + 4 points + 1 flags + 8 shaping
BYTES_PER_BOARD4 = 92          # g2048_step4 (the C2 form, one board -> its four transitions): 8 board in + 4 x (8 + 4 + 1 + 8) out = 23 B / transition
# dram__bytes_read.sum + dram__bytes_write.sum of one step_kernel_dense launch on this workload, from the
# committed `ncu --set full` capture (profiles/r02_step_dense_ncu.txt: 38.0 MB read + 33.2 MB written; the rest of
# the 88 MB of outputs is still dirty in the 126 MB L2 when the kernel ends)
NCU_TRAFFIC_BYTES_PER_LAUNCH = 71.2e6
# the same for one step4_kernel_dense launch (profiles/r02_step4_ncu.txt)
NCU_TRAFFIC_BYTES_PER_LAUNCH_STEP4 = 37.5e6     # 8.7 MB read + 28.9 MB written (of 96 MB algorithmic: the outputs are still dirty in L2)
METRIC = "env_steps_per_sec"
UNIT = "env-steps/s"
WORKLOAD = ("c2_env_step: 2^20 boards x 4 moves = 4194304 full Game2048.step transitions per GPU per step "
            "(move+merge points+shaping+Philox spawn+legal/done)")


def python_reference(seconds: float = 2.0, train_steps: int = 12):
    """BASELINE.md section 3 B1-B5 on the UNMODIFIED Python reference staged under baseline/_ref (own process: its worker
    pool must not fork a CUDA context; B5 = `train_steps` steps of the config #1 command on CPU).  None when the reference is not staged."""
    import subprocess
    script = os.path.join(ROOT, "baseline", "time_reference.py")
    if not os.path.exists(os.path.join(ROOT, "baseline", "_ref", "game.py")):
        return {"unavailable": "baseline/_ref is not staged on this box"}
    env = dict(os.environ)
    for k in ("OMP_NUM_THREADS", "MKL_NUM_THREADS"):     # torchrun exports OMP_NUM_THREADS=1
        env.pop(k, None)
    try:
        out = subprocess.run([sys.executable, script, "--seconds", str(seconds), "--train-steps", str(train_steps)], capture_output=True,
                             text=True, timeout=420, env=env)
        return json.loads(out.stdout.strip().splitlines()[-1])
    except Exception as e:          # a baseline that cannot run is reported, not fatal
        return {"unavailable": f"{type(e).__name__}: {e}"}


def c2_boards(seed: int) -> np.ndarray:
    """SURVEY 8(d) C2 input: exponents 1..11, each cell emptied with probability 0.30."""
    import torch
    g = torch.Generator().manual_seed(seed)
    e = torch.randint(1, 12, (N_BOARDS, 16), generator=g, dtype=torch.int64)
    e[torch.rand((N_BOARDS, 16), generator=g) < 0.30] = 0
    sh = torch.arange(16, dtype=torch.int64) * 4
    return (e << sh).sum(dim=1).numpy()


def c2_transitions(seed: int):
    """(boards int64[4M], actions uint8[4M]): transition d*2^20+i = board i, move d."""
    b = c2_boards(seed)
    boards = np.tile(b, MOVES)
    actions = np.repeat(np.arange(MOVES, dtype=np.uint8), N_BOARDS)
    return boards, actions


def bf16_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return float(json.load(open(path))["bf16_tflops"])
    return 1590.0


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return float(d["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """Polls SM clock / throttle reasons through NVML while the timed region runs."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception:
            self.nv = None

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    r = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    r = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                break
            time.sleep(0.002 if len(self.samples) < 200 else 0.05)     # dense over the timed step region, sparse over the later sections

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        med = int(np.median(self.samples)) if self.samples else None
        busy = [c for c in self.samples if self.max_mhz is None or c >= 0.5 * self.max_mhz]     # idle gaps between sections clock down
        med = int(np.median(busy)) if busy else med
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples), "sm_mhz_min": int(min(self.samples)) if self.samples else None}


def cpu_baseline(min_seconds: float = 10.0):
    """C port of Game2048.step (oracle/) on all host threads, bounded sample of the C2 workload."""
    from oracle import oracle as O
    O.set_threads(os.cpu_count() or 1)        # torchrun exports OMP_NUM_THREADS=1; the baseline gets every core
    boards, actions = c2_transitions(2048)
    n = 1 << 21
    b, a = boards[:n].view(np.uint64), actions[:n]
    # interleave the four moves so that the sample is representative
    idx = (np.arange(n) % MOVES) * N_BOARDS + (np.arange(n) // MOVES)
    b, a = np.ascontiguousarray(boards.view(np.uint64)[idx]), np.ascontiguousarray(actions[idx])
    O.step_batch(b[:4096], a[:4096], seed=1, env0=0, ctr=1)
    t0 = time.perf_counter()
    done = 0
    reps = 0
    while time.perf_counter() - t0 < min_seconds:
        O.step_batch(b, a, seed=1, env0=0, ctr=1 + reps)
        done += n
        reps += 1
    dt = time.perf_counter() - t0
    return {"value": done / dt, "unit": UNIT, "cores": O.num_threads(), "kind": "port",
            "sample": f"{reps} x {n} C2 transitions ({dt:.1f} s) through oracle/oracle2048.c orc_step_batch (OpenMP)"}


def run_reference(args):
    """The reference's CPU algorithm for the path on every host thread: the C port (oracle/oracle2048.c, OpenMP) steps the
    same 4 194 304 C2 transitions per step as our arm; the UNMODIFIED Python reference (baseline/_ref) is timed beside it on
    a bounded sample (it runs ~4e3 env-steps/s/core: a full step would take minutes)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as O
    O.set_threads(os.cpu_count() or 1)        # torchrun exports OMP_NUM_THREADS=1; the reference arm gets every core
    boards, actions = c2_transitions(2048)
    b, a = np.ascontiguousarray(boards.view(np.uint64)), np.ascontiguousarray(actions)
    n = N_TRANS
    for w in range(args.warmup):
        O.step_batch(b, a, seed=1, env0=0, ctr=w)
    t0 = time.perf_counter()
    for k in range(args.steps):
        O.step_batch(b, a, seed=1, env0=0, ctr=100 + k)
    dt = time.perf_counter() - t0
    val = n * args.steps / dt
    sample = f"each step = all {N_TRANS} C2 transitions, oracle/oracle2048.c orc_step_batch, {O.num_threads()} OpenMP threads"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "boards": N_BOARDS, "moves": MOVES},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": O.num_threads(), "kind": "port", "sample": sample},
        "python_reference": python_reference(2.0),
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def rollout_section(args, dev, world, rank, barrier):
    """BASELINE configs C3 / C4 / C5.
      C3 (N = 1 only): GameMLP h=196 L=2, 65536 envs x 512 steps on one GPU, Kaiming init with non-zero heads, README reward
          flags: the fused rollout kernel alone (the fp32-grade split-fp16 tcgen05 kernel that `auto` selects, with the fp32
          FFMA kernel and the bf16 tcgen05 kernel as labelled variants) and the full rollout + advantage + update train step.
      C4 (every N): 2^20 envs x 512 steps in total, sharded contiguously over the N ranks (N = 1: the one-GPU time of the same
          job, the denominator of the strong-scaling figure), rollout + update with the flat-gradient and moment all-reduces;
          the same step with the collectives switched off (each rank on its own shard: what one GPU does with that per-GPU
          env count) and the gradient all-reduce timed on its own.
      C5: GameURM rollout."""
    import torch
    import torch.distributed as dist

    from g2048 import dp, trainer as tr

    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    flops = 2 * (48 * 196 + 2 * 196 * 196 + 5 * 196)

    def max_over_ranks(*vals):
        if world == 1:
            return vals if len(vals) > 1 else vals[0]
        v = torch.tensor(list(vals), device=dev, dtype=torch.float64)
        dist.all_reduce(v, op=dist.ReduceOp.MAX)
        out = tuple(float(x) for x in v.tolist())
        return out if len(out) > 1 else out[0]

    def timed(fn, reps):
        fn()                       # warm-up (allocations, first-launch costs)
        barrier()
        ev0.record()
        for _ in range(reps):
            r = fn()
        ev1.record()
        barrier()
        return ev0.elapsed_time(ev1) / reps, r

    def measure(envs_total, horizon, reps, variants):
        # the model the reference builds (train.py:1522: MLPConfig default, Dropout(0.1) active in the update forward, train.py:483)
        cfg = tr.TrainConfig(hidden_dim=196, num_layers=2, envs=envs_total, horizon=horizon, zero_heads=False, seed=2048, dropout=0.1)
        t = tr.Trainer(cfg, dev)
        n_global = envs_total * horizon
        res = {"envs_total": envs_total, "envs_per_gpu": t.B, "horizon": horizon, "timing_reps": reps}

        def time_rollout(precision, r):
            cfg.rollout_precision = precision
            ms, _ = timed(t.collect, r)
            return max_over_ranks(ms)

        ro_ms = time_rollout("auto", reps)
        res["rollout_ms"] = ro_ms
        res["env_steps_per_sec"] = n_global / (ro_ms * 1e-3)
        res["rollout_kernel"] = ("rollout_mlp_x3_kernel<208> (tcgen05.mma on split-fp16 operands: hi*hi + hi*lo + lo*hi, fp32 accumulate "
                                 "in TMEM; log-probs / values within 2e-5 of the fp32 policy)")
        res["rollout_model_tflops"] = res["env_steps_per_sec"] * flops / 1e12
        res["rollout_mma_tflops"] = 3 * res["rollout_model_tflops"]
        res["rollout_frac_of_bf16_peak_mma"] = 3 * (res["env_steps_per_sec"] / world) * flops / 1e12 / bf16_peak()
        if variants:
            ms = time_rollout("fp32", max(3, reps // 3))
            res["fp32_ffma_rollout_variant"] = {"ms": ms, "env_steps_per_sec": n_global / (ms * 1e-3), "kernel": "rollout_mlp_kernel<208>"}
            ms = time_rollout("bf16", reps)
            res["bf16_rollout_variant"] = {"ms": ms, "env_steps_per_sec": n_global / (ms * 1e-3), "kernel": "rollout_mlp_tc_kernel<208>",
                                           "note": "bf16-rounded operands: log-probs ~1e-2 off the fp32 policy -- NOT reference precision, never selected by `auto`"}
        cfg.rollout_precision = "auto"
        step_ms, stats = timed(t.train_step, reps)
        times = t.times
        step_ms = max_over_ranks(step_ms)
        res["train_step_ms"] = step_ms
        res["rollout_update_steps_per_sec"] = n_global / (step_ms * 1e-3)
        res["phase_ms_rank0"] = {"rollout": times.rollout_ms, "advantage": times.advantage_ms, "update": times.update_ms,
                                 "grad_allreduce_last_step": times.grad_allreduce_ms, "moments_allreduce": times.allreduce_ms}
        res["loss"] = stats["loss"]
        res["update_dropout"] = cfg.dropout
        if variants:
            cfg.dropout = 0.0                     # information only: the same step on a model without dropout
            ms0, _ = timed(t.train_step, reps)
            res["train_step_ms_without_dropout"] = max_over_ranks(ms0)
            res["update_ms_rank0_without_dropout"] = t.times.update_ms
            cfg.dropout = 0.1
            cfg.kl_stats = True                   # + the KL(old || new) statistic the reference logs (train.py:577-597): a second,
            ms1, st1 = timed(t.train_step, reps)  # forward-only pass over the batch with the updated weights
            res["train_step_ms_with_kl_statistic"] = max_over_ranks(ms1)
            res["kl_average"] = st1.get("kl_average")
            cfg.kl_stats = False
        if world > 1:
            # the same train step without the collectives: every rank alone on its shard
            dp.COLLECTIVES = False
            local_ms, _ = timed(t.train_step, reps)
            dp.COLLECTIVES = True
            local_ms = max_over_ranks(local_ms)
            res["train_step_ms_without_collectives"] = local_ms
            res["weak_scaling_efficiency_vs_same_per_gpu_envs"] = local_ms / step_ms
            ar_ms, _ = timed(lambda: t.bucket.allreduce(), 50)
            res["grad_allreduce_ms"] = max_over_ranks(ar_ms)
            res["grad_allreduce_bytes"] = t.bucket.numel() * 4
        if variants and args.update_variants:
            for name in ("tf32", "fp32", "x3"):
                cfg.update_matmul = name
                ms, _ = timed(t.train_step, 2)
                res[f"{name}_autograd_update_variant"] = {"train_step_ms": max_over_ranks(ms)}
            cfg.update_matmul = "fused"
        del t
        torch.cuda.empty_cache()
        return res

    out = {"model_flops_per_env_step": flops,
           "update": "update_mlp_x3_kernel (one pipelined tcgen05 kernel: GameMLP forward + PPO loss + backward-data, split-fp16 GEMMs, "
                     "three products per k-step in both directions, Philox Dropout(0.1) masks as the reference's update forward has them) + "
                     "x3_wgrad_kernel weight gradients on the fp16 operand images, Muon+AdamW"}
    if world == 1:
        c3 = measure(args.rollout_envs, args.rollout_steps, 10, True)
        c3["workload"] = (f"c3: GameMLP h=196 L=2 fused rollout, {args.rollout_envs} envs x {args.rollout_steps} steps on one GPU, "
                          "auto-reset, Philox seed 2048")
        out["c3"] = c3
    if args.c4_envs > 0:
        per_gpu = args.c4_envs // world
        c4 = measure(args.c4_envs, args.rollout_steps, 3 if per_gpu >= (1 << 19) else 5, False)
        c4["workload"] = (f"c4: {args.c4_envs} envs x {args.rollout_steps} steps in total, sharded over {world} GPU(s) "
                          f"({per_gpu} envs per GPU), rollout + advantage + fused update + flat-gradient all-reduce (NCCL) + Muon/AdamW")
        out["c4"] = c4
    # ---- C5: GameURM rollout (default config) at the configured env count
    if args.urm_envs > 0:
        from g2048 import env as genv, rollout as gro
        from g2048.policy import GameURM, GameURMConfig
        torch.manual_seed(5)
        um = GameURM(GameURMConfig(dropout=0.0)).to(dev).eval()
        upol = gro.pack_policy(um)
        ub = genv.reset(args.urm_envs, device=dev, seed=5, env0=rank * args.urm_envs, ctr=0)
        ubuf = gro.RolloutBuffers.allocate(args.urm_steps, args.urm_envs, dev)
        ctr = [1]

        prec = ["x3"]

        def urm_once():
            gro.rollout(upol, ub, args.urm_steps, seed=5, env0=rank * args.urm_envs, ctr0=ctr[0], out=ubuf, precision=prec[0])
            ctr[0] += args.urm_steps
        ums, _ = timed(urm_once, 2)
        ums = max_over_ranks(ums)
        # projections: 2 * 16 tokens * (64*192 + 64*64 + 64*240 + 120*64) per block application, attention 2 * 2 * 4 heads * 16 * 16 * 16
        urm_flops = um.config.num_loops * um.config.num_layers * (2 * 16 * (64 * 192 + 64 * 64 + 64 * 240 + 120 * 64) + 4 * 4 * 16 * 16 * 16)
        rate = world * args.urm_envs * args.urm_steps / (ums * 1e-3)
        out["urm"] = {"workload": f"c5: GameURM (hidden 64, 2 layers, 4 heads, 4 loops) fused rollout, {args.urm_envs} envs x {args.urm_steps} steps per GPU",
                      "env_steps_per_sec": rate, "ms": ums, "model_flops_per_env_step": urm_flops,
                      "model_tflops": rate * urm_flops / 1e12 / world, "mma_tflops": rate * 3 * (urm_flops - um.config.num_loops * um.config.num_layers * 4 * 4 * 16 * 16 * 16) / 1e12 / world,
                      "kernel": "rollout_urm_x3_kernel (tcgen05.mma projections on split-fp16 operands, three products per k-step; fp32 K/V, attention, "
                                "norms and SwiGLU on CUDA cores; log-probs / values within 2e-5 of the fp32 model)"}
        prec[0] = "fp16"
        vms, _ = timed(urm_once, 2)
        vms = max_over_ranks(vms)
        out["urm"]["fp16_variant"] = {"ms": vms, "env_steps_per_sec": world * args.urm_envs * args.urm_steps / (vms * 1e-3), "kernel": "rollout_urm_kernel",
                                      "note": "single fp16 operands and fp16 K/V: log-probs ~1e-2 off the fp32 model -- NOT reference precision, never selected by default"}
        del ubuf, ub
        # GameURM train step: fused rollout + advantage + update on this library's kernels (g2048/urm_ops.py),
        # with the torch mirror's own forward / backward (ATen + cuBLAS) timed beside it
        if args.urm_train_envs > 0:
            utr = {}
            for mode in ("ops", "autograd"):
                ucfg = tr.TrainConfig(model_type="urm", envs=args.urm_train_envs * world, horizon=args.urm_train_steps, zero_heads=False,
                                      warmup_steps=0, seed=5, urm_update=mode, urm_chunk=(1 << 17) if mode == "ops" else (1 << 15))
                ut = tr.Trainer(ucfg, dev)
                ms, _ = timed(ut.train_step, 2 if mode == "ops" else 1)
                ms = max_over_ranks(ms)
                utr[mode] = {"train_step_ms": ms, "update_ms_rank0": ut.times.update_ms, "rollout_ms_rank0": ut.times.rollout_ms,
                             "rollout_update_steps_per_sec": world * args.urm_train_envs * args.urm_train_steps / (ms * 1e-3)}
                del ut
            out["urm"]["train_step"] = {
                "workload": f"GameURM rollout + advantage + update, {args.urm_train_envs} envs x {args.urm_train_steps} steps per GPU, Muon+AdamW",
                "update": "projections on x3_gemm_kernel / x3_wgrad_kernel (tcgen05), attention / ConvSwiGLU / RMS-norm forward and backward on "
                          "g2048_urm_train.cu kernels; torch autograd only as the tape",
                **utr["ops"], "aten_autograd_variant": utr["autograd"]}
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist

    from g2048 import env

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    env.init(dev)
    env.lut(dev)

    # ring of buffer sets larger than L2 (126 MB): 6 x 120 MiB touched round-robin
    RING = 6
    sets = []
    for s in range(RING):
        b, a = c2_transitions(2048 + 97 * s + 1009 * rank)
        sets.append(dict(
            boards=torch.from_numpy(b).to(dev), actions=torch.from_numpy(a).to(dev),
            out=dict(boards=torch.empty(N_TRANS, dtype=torch.int64, device=dev),
                     points=torch.empty(N_TRANS, dtype=torch.int32, device=dev),
                     flags=torch.empty(N_TRANS, dtype=torch.uint8, device=dev),
                     shaping=torch.empty(N_TRANS, dtype=torch.int64, device=dev))))
    env0 = rank * N_TRANS

    def one_step(k):
        s = sets[k % RING]
        env.step(s["boards"], s["actions"], seed=2048, env0=env0, ctr=1 + k, shaping=True, out=s["out"])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def capture(fn, count):
        """The K launches of the timed region as one CUDA graph: the kernels are tens of microseconds,
        shorter than a Python + ctypes call, so launching them one by one would time the host."""
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            fn(0)
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            for k in range(count):
                fn(k)
        graph.replay()     # untimed: the first launch of a graph also uploads it to the device (~2 % of a 200-node replay)
        return graph

    for w in range(max(args.warmup, 3)):
        one_step(w)
    barrier()
    step_graph = capture(one_step, args.steps)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    step_graph.replay()               # exactly args.steps launches of g2048_step
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    pairs_ms_per_step = ms / args.steps           # g2048_step on the 4 Mi (board, action) pairs
    pairs_value = world * N_TRANS * args.steps / (ms * 1e-3)

    # ---- the headline: the same 4 194 304 transitions in C2's own form, g2048_step4 (one thread plays the four moves of a board
    # and shares their common work; bit for bit the transitions g2048_step returns for the pairs -- tests/test_env_gpu.py).  The
    # [n,4] outputs reuse the pair buffers.
    s4_in = [s["boards"][:N_BOARDS] for s in sets]
    s4_out = [{k: v.view(N_BOARDS, 4) for k, v in s["out"].items()} for s in sets]

    def one_step4(k):
        env.step4(s4_in[k % RING], seed=2048, env0=env0, ctr=1 + k, shaping=True, out=s4_out[k % RING])

    for w in range(max(args.warmup, 3)):
        one_step4(w)
    barrier()
    s4_graph = capture(one_step4, args.steps)
    barrier()
    ev0.record()
    s4_graph.replay()                 # exactly args.steps launches of g2048_step4
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    value = world * N_TRANS * args.steps / (ms * 1e-3)
    global PAIRS_FORM
    PAIRS_FORM = {"ms_per_launch": pairs_ms_per_step, "env_steps_per_sec": pairs_value, "bytes_per_transition": BYTES_PER_TRANSITION,
                  "achieved_gbs": BYTES_PER_TRANSITION * N_TRANS / (pairs_ms_per_step * 1e-3) / 1e9, "kernel": "step_kernel_dense<true, false>",
                  "traffic": NCU_TRAFFIC_BYTES_PER_LAUNCH,
                  "note": "g2048_step on 4 194 304 independent (board, action) pairs: the general form of the same transitions"}

    # ---- the same step without the shaping record (shaping == NULL: 22 B / transition)
    ns_out = [dict(boards=s["out"]["boards"], points=s["out"]["points"], flags=s["out"]["flags"]) for s in sets]
    ns_graph = capture(lambda k: env.step(sets[k % RING]["boards"], sets[k % RING]["actions"], seed=2048, env0=env0, ctr=1 + k,
                                          shaping=False, out=ns_out[k % RING]), args.steps)
    barrier()
    ev0.record()
    ns_graph.replay()
    ev1.record()
    barrier()
    ns_ms = ev0.elapsed_time(ev1) / args.steps
    no_shaping = {"ms_per_launch": ns_ms, "bytes_per_transition": 22, "env_steps_per_sec": world * N_TRANS / (ns_ms * 1e-3),
                  "achieved_gbs": 22 * N_TRANS / (ns_ms * 1e-3) / 1e9, "kernel": "step_kernel_dense<false, false>"}

    # ---- C2 with the spawn draws REPLAYED from a buffer (config 2 as worded: "bit-exact vs game.py on replayed spawns"):
    # the same kernel reads (u0, u1) per transition -- 8 more input bytes -- instead of running Philox (36 instructions)
    gen = torch.Generator(device=dev).manual_seed(4096 + rank)
    rp = [torch.randint(-2 ** 31, 2 ** 31 - 1, (N_BOARDS, 4, 2), generator=gen, device=dev, dtype=torch.int64).to(torch.int32) for _ in range(RING)]
    rp_graph = capture(lambda k: env.step4(s4_in[k % RING], seed=2048, env0=env0, ctr=1 + k, shaping=True, out=s4_out[k % RING],
                                           replay=rp[k % RING]), args.steps)
    barrier()
    ev0.record()
    rp_graph.replay()
    ev1.record()
    barrier()
    rp_ms = ev0.elapsed_time(ev1) / args.steps
    replayed = {"ms_per_launch": rp_ms, "env_steps_per_sec": world * N_TRANS / (rp_ms * 1e-3), "kernel": "step4_kernel_dense<true, true>",
                "bytes_per_transition": 23, "achieved_gbs": 23 * N_TRANS / (rp_ms * 1e-3) / 1e9,
                "bytes_per_transition_with_the_draws": 31, "achieved_gbs_with_the_draws": 31 * N_TRANS / (rp_ms * 1e-3) / 1e9,
                "note": "g2048_step4 with replay != NULL: two u32 draws per transition read from HBM (the reference's own draws in the parity "
                        "tests) instead of the Philox stream; SURVEY 8(d) counts no bytes for the draws (first pair of figures), the kernel "
                        "does read them (second pair)"}
    del rp_graph, rp

    # ---- C2 in its 4-move expansion form (g2048_expand4: all four pre-spawn successors per board)
    ex_sets = [dict(succ=torch.empty((N_BOARDS, 4), dtype=torch.int64, device=dev),
                    points=torch.empty((N_BOARDS, 4), dtype=torch.int32, device=dev),
                    legal=torch.empty(N_BOARDS, dtype=torch.uint8, device=dev), max_tile=None) for _ in range(RING * 2)]
    # 12 distinct C2 board sets (the same distribution as the step input: exponents <= 11)
    ex_in = [s["boards"][:N_BOARDS] for s in sets] + [torch.from_numpy(c2_boards(5000 + 31 * s + 1009 * rank)).to(dev)
                                                      for s in range(RING)]
    for k in range(3):
        env.expand4(ex_in[k], out=ex_sets[k])
    barrier()
    ex_steps = max(args.steps, 20)
    ex_graph = capture(lambda k: env.expand4(ex_in[k % len(ex_in)], out=ex_sets[k % len(ex_sets)]), ex_steps)
    barrier()
    ev0.record()
    ex_graph.replay()
    ev1.record()
    barrier()
    ex_ms = ev0.elapsed_time(ev1) / ex_steps
    if world > 1:
        t = torch.tensor([ex_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ex_ms = float(t.item())
    # second generation: the boards g2048_step just produced (a few % hold a 4096 tile and take the L2 table path)
    ex2_in = [s["out"]["boards"][:N_BOARDS] for s in sets]
    ex2_graph = capture(lambda k: env.expand4(ex2_in[k % len(ex2_in)], out=ex_sets[k % len(ex_sets)]), ex_steps)
    barrier()
    ev0.record()
    ex2_graph.replay()
    ev1.record()
    barrier()
    ex2_ms = ev0.elapsed_time(ev1) / ex_steps
    expand = {"no_shaping_step": no_shaping, "replayed_step": replayed, "ms_per_launch": ex_ms, "boards_per_launch": N_BOARDS, "ms_per_launch_on_post_step_boards": ex2_ms,
              "transitions_per_sec": world * N_BOARDS * 4 / (ex_ms * 1e-3),
              "bytes_per_board": 57, "achieved_gbs": 57 * N_BOARDS / (ex_ms * 1e-3) / 1e9,
              "l2": "12 distinct input / output sets (57 MiB each) used round-robin"}

    # the same kernel on 2^23 boards per launch: the fixed per-launch cost (~9 us) no longer hides the kernel
    big_n = 1 << 23
    g0 = torch.Generator(device=dev).manual_seed(7 + rank)
    big_in = []
    for _ in range(2):
        e = torch.randint(1, 12, (big_n, 16), generator=g0, device=dev, dtype=torch.int64)
        e[torch.rand((big_n, 16), generator=g0, device=dev) < 0.30] = 0
        big_in.append((e << (torch.arange(16, device=dev) * 4)).sum(1))
        del e
    big_out = [dict(succ=torch.empty((big_n, 4), dtype=torch.int64, device=dev),
                    points=torch.empty((big_n, 4), dtype=torch.int32, device=dev),
                    legal=torch.empty(big_n, dtype=torch.uint8, device=dev), max_tile=None) for _ in range(2)]
    for k in range(3):
        env.expand4(big_in[k % 2], out=big_out[k % 2])
    barrier()
    ev0.record()
    for k in range(10):
        env.expand4(big_in[k % 2], out=big_out[k % 2])
    ev1.record()
    barrier()
    big_ms = ev0.elapsed_time(ev1) / 10
    expand["large_batch"] = {"boards_per_launch": big_n, "ms_per_launch": big_ms,
                             "transitions_per_sec": world * big_n * 4 / (big_ms * 1e-3),
                             "achieved_gbs": 57 * big_n / (big_ms * 1e-3) / 1e9,
                             "l2": "2 input / output sets of 456 MiB each, alternating"}
    del big_in, big_out
    torch.cuda.empty_cache()

    e2e_value = e2e_ms = None
    e2e_steps = 0
    h2d = N_BOARDS * 8
    d2h = N_TRANS * (8 + 4 + 1 + 8)
    if not args.no_e2e:
        e2e_value, e2e_ms, e2e_steps = e2e_section(args, dev, world, rank, barrier, sets, env0)

    # ---- C3 / C4: fused MLP rollout (65536 envs x 512 steps per GPU) and rollout+update
    ro = None
    if not args.no_rollout:
        ro = rollout_section(args, dev, world, rank, barrier)

    clocks = sampler.stop()      # sampled from the timed step region to the end of the GPU sections (the K timed launches alone last ~1.5 ms)
    if rank == 0:
        finish(args, world, value, ms_per_step, clocks, e2e_value, e2e_ms, e2e_steps, h2d, d2h, ro, RING, expand)
    if world > 1:
        dist.destroy_process_group()


def e2e_section(args, dev, world, rank, barrier, sets, env0):
    """Same metric through the public host-buffer API (g2048.env.HostStepper4, and HostStepper for the pairs form) with HOST pinned
    buffers: every step copies the inputs host->device, runs g2048_step4 and copies every output device->host (2^20-transition
    chunks on 2 streams so the two PCIe directions and the kernel overlap; tools/pcie_bw.py / tools/time_e2e.py: the device->host
    direction is the ceiling for every chunking)."""
    import torch
    import torch.distributed as dist

    from g2048 import env
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    hb, ha = c2_transitions(4242 + rank)
    old_affinity = env.bind_host_thread_to_gpu(dev)      # pinned buffers next to the GPU (restored below)
    print(f"[bench] rank {rank}: host thread bound to {len(os.sched_getaffinity(0))} of {os.cpu_count()} cpus for the e2e leg"
          f" ({'NVML affinity' if old_affinity else 'unbound'})", file=sys.stderr)
    h_boards, h_actions = torch.from_numpy(hb).pin_memory(), torch.from_numpy(ha).pin_memory()
    h_out = dict(boards=torch.empty(N_TRANS, dtype=torch.int64).pin_memory(),
                 points=torch.empty(N_TRANS, dtype=torch.int32).pin_memory(),
                 flags=torch.empty(N_TRANS, dtype=torch.uint8).pin_memory(),
                 shaping=torch.empty(N_TRANS, dtype=torch.int64).pin_memory())
    stepper = env.HostStepper(N_TRANS, device=dev)
    stepper4 = env.HostStepper4(N_BOARDS, device=dev)
    h_boards4 = torch.from_numpy(hb[:N_BOARDS].copy()).pin_memory()
    h_out4 = {k: v.view(N_BOARDS, 4) for k, v in h_out.items()}

    def e2e_step(k):
        stepper4.step(h_boards4, h_out4, seed=2048, env0=env0, ctr=1000 + k)

    def e2e_step_pairs(k):
        stepper.step(h_boards, h_actions, h_out, seed=2048, env0=env0, ctr=1000 + k)

    e2e_steps = max(3, min(args.steps, 20))

    def run(fn, after=None):
        for w in range(3):
            fn(w)
        if after:
            after()
        barrier()
        ev0.record()
        for k in range(e2e_steps):
            fn(k)
        if after:
            after()                    # the current stream waits for every enqueued step: the downloads end inside the timed region
        ev1.record()
        barrier()
        ms = ev0.elapsed_time(ev1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    pairs_ms = run(e2e_step_pairs)
    e2e_ms = run(e2e_step)
    # the same steps enqueued without a join between them (HostStepper4.step(join=False) ... join()): the upload and kernel of step
    # k + 1 run under the download of step k; two sets of host output buffers, as a consumer of a streaming API would hold
    h_out4b = {k: torch.empty_like(v).pin_memory() for k, v in h_out4.items()}
    pipe_ms = run(lambda k: stepper4.step(h_boards4, (h_out4, h_out4b)[k % 2], seed=2048, env0=env0, ctr=3000 + k, join=False), after=stepper4.join)
    del h_out4b
    joined_ms, e2e_ms = e2e_ms, pipe_ms          # headline: the streaming use of the API; the per-step-joined call beside it
    e2e_value = world * N_TRANS * e2e_steps / (e2e_ms * 1e-3)
    # information only: the same call without the shaping record (13 B instead of 21 B per transition device->host) -- the
    # device->host direction is what bounds this leg
    stepper_ns = env.HostStepper(N_TRANS, device=dev, shaping=False)
    h_out_ns = {k: v for k, v in h_out.items() if k != "shaping"}
    for w in range(2):
        stepper_ns.step(h_boards, h_actions, h_out_ns, seed=2048, env0=env0, ctr=2000 + w)
    barrier()
    ev0.record()
    for k in range(e2e_steps):
        stepper_ns.step(h_boards, h_actions, h_out_ns, seed=2048, env0=env0, ctr=2100 + k)
    ev1.record()
    barrier()
    ns_ms = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([ns_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ns_ms = float(t.item())
    global E2E_EXTRA
    E2E_EXTRA = {"api": "g2048.env.HostStepper4.step(join=False) x K, then join() (g2048_step4 on pinned host buffers, 2^18-board chunks on 2 streams; "
                        "every step uploads its boards and downloads all its outputs; consecutive steps overlap on the stepper's streams, two sets "
                        "of host output buffers)",
                 "pcie_d2h_gbs_per_gpu": N_TRANS * 21 * e2e_steps / (e2e_ms * 1e-3) / 1e9,
                 "pcie_h2d_gbs_per_gpu": N_BOARDS * 8 * e2e_steps / (e2e_ms * 1e-3) / 1e9,
                 "joined_per_step": {"value": world * N_TRANS * e2e_steps / (joined_ms * 1e-3), "ms_per_step": joined_ms / e2e_steps,
                                     "pcie_d2h_gbs_per_gpu": N_TRANS * 21 * e2e_steps / (joined_ms * 1e-3) / 1e9,
                                     "api": "HostStepper4.step(join=True): the caller's stream waits for every step before the next one is enqueued"},
                 "pairs_form": {"value": world * N_TRANS * e2e_steps / (pairs_ms * 1e-3), "h2d_bytes_per_step": N_TRANS * 9, "d2h_bytes_per_step": N_TRANS * 21,
                                "api": "g2048.env.HostStepper.step (g2048_step on (board, action) pairs)"},
                 "without_shaping_record": {"value": world * N_TRANS * e2e_steps / (ns_ms * 1e-3), "d2h_bytes_per_step": N_TRANS * 13,
                                            "pcie_d2h_gbs_per_gpu": N_TRANS * 13 * e2e_steps / (ns_ms * 1e-3) / 1e9,
                                            "note": "information only: the headline e2e returns every output of the step"}}
    if old_affinity:
        os.sched_setaffinity(0, old_affinity)            # the cpu_baseline leg uses every core
    return e2e_value, e2e_ms, e2e_steps


E2E_EXTRA = {}
PAIRS_FORM = {}


def finish(args, world, value, ms_per_step, clocks, e2e_value, e2e_ms, e2e_steps, h2d, d2h, ro, RING, expand):
    if True:
        peak, which = peaks()
        expand["frac_of_hbm_peak"] = expand["achieved_gbs"] / peak
        no_shaping = expand.pop("no_shaping_step")
        replayed = expand.pop("replayed_step")
        replayed["frac_of_hbm_peak"] = replayed["achieved_gbs"] / peak
        replayed["frac_of_hbm_peak_with_the_draws"] = replayed["achieved_gbs_with_the_draws"] / peak
        no_shaping["frac_of_hbm_peak"] = no_shaping["achieved_gbs"] / peak
        expand["large_batch"]["frac_of_hbm_peak"] = expand["large_batch"]["achieved_gbs"] / peak
        achieved = BYTES_PER_BOARD4 * N_BOARDS / (ms_per_step * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "boards": N_BOARDS, "moves": MOVES, "kernel_launches_per_step": 1,
                       "entry_point": "g2048_step4: the four moves of every board from one thread, each transition with its own Philox spawn (env id 4 b + m)", "l2": f"ring of {RING} distinct 120 MiB buffer sets (> 126 MB L2) used round-robin",
                       "spawn": "philox4x32-10", "launch": "the K timed launches are one replay of a CUDA graph (after one untimed replay that uploads it)"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "ms_per_step": (e2e_ms / e2e_steps) if e2e_steps else None, **E2E_EXTRA},
            "gpu_launches": args.steps,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": NCU_TRAFFIC_BYTES_PER_LAUNCH_STEP4 or None, "algorithmic_bytes_per_launch": BYTES_PER_BOARD4 * N_BOARDS,
                         "peak_source": which, "kernel": "step4_kernel_dense<true, false>",
                         "bytes_per_unit": BYTES_PER_BOARD4 / MOVES, "units_per_launch": N_TRANS},
        }
        if PAIRS_FORM:
            PAIRS_FORM["frac_of_hbm_peak"] = PAIRS_FORM["achieved_gbs"] / peak
            line["step_pairs_form"] = PAIRS_FORM
        line["step_without_shaping"] = no_shaping
        line["step_replayed_draws"] = replayed
        line["expand4"] = expand
        if ro is not None:
            line["rollout"] = ro
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline(args.cpu_seconds)
            line["cpu_baseline"]["python_reference"] = python_reference(2.0)
        print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--no-rollout", action="store_true", help="skip the C3 rollout / update section")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer e2e leg (profiling runs only)")
    ap.add_argument("--rollout-envs", type=int, default=65536)
    ap.add_argument("--rollout-steps", type=int, default=512)
    ap.add_argument("--c4-envs", type=int, default=1 << 20, help="config #4: total env count sharded over the GPUs (0 = skip)")
    ap.add_argument("--update-variants", action="store_true", help="also time the autograd update variants (cuBLAS fp32 / TF32, x3 GEMMs)")
    ap.add_argument("--urm-envs", type=int, default=262144, help="config #5 env count per GPU (0 = skip)")
    ap.add_argument("--urm-steps", type=int, default=64)
    ap.add_argument("--urm-train-envs", type=int, default=65536, help="GameURM train step: envs per GPU (0 = skip)")
    ap.add_argument("--urm-train-steps", type=int, default=4)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
