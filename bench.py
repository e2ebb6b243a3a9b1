#!/usr/bin/env python
"""SMPL forward benchmark (BASELINE.json metric: SMPL forward bodies/sec; LBS % HBM, GEMM % TC).

    python bench.py --gpus N --steps K --warmup W [--impl reference]

One "step" = one SMPL forward (k2 chain -> blendshapes -> LBS + projection) over one batch of
synthetic per-person parameters; the workload is BASELINE.json configs[2] ("batch 4096 ...
tcgen05 GEMM regime ... fp32 LBS") per GPU.  With N > 1 (torchrun, one process per GPU) every rank
runs its own 4096-body shard (weak scaling, no data-path collective) and every step exchanges the
joints + kp2d rows of all ranks (480 B/body, configs[3]) -- by peer stores over NVLink on a side
stream (sharding.PeerExchange), in BOTH the `value` and the `e2e` leg.

Printed JSON (rank 0, one line):
  value       bodies/s, whole job, inputs resident in HBM, CUDA-event timed, max over ranks
  e2e         same metric through the C-ABI host entry (smplb200_forward_host): pinned host inputs
              H2D + forward + D2H of the WHOLE result (vertices + joints + kp2d) every step;
              `e2e_small_outputs` is the same call returning joints + kp2d only (vertices stay on the
              device, as they do for the reference's GPU module) -- quote both
  roofline    dominant kernel against the measured HBM peak; roofline_kernels lists all, plus
              the blendshape GEMM against TF32 / bf16 tensor peaks measured in this run
  accuracy    max abs error of every precision mode against the fp32 CPU oracle, MEASURED in this run
  configs_extra  BASELINE.json configs[1] (64 bodies, fp32 FMA regime) and configs[4] (DLA-34 -> decode
              -> SMPL for 32 x 32 people, parity asserted in the run)
  cpu_baseline   the CPU oracle (oracle/smpl_ref.py, fp32 eager PyTorch) timed on the host cores

`--impl reference` times the reference arm.  The reference snapshot contains no SMPL layer and
no compilable source for this path, so the arm runs the oracle restatement of the
eager PyTorch layer on the host cores ("kind": "port") on the SAME workload: 4096 bodies per step
(evaluated in 512-body chunks to bound the eager T[N,V,4,4] intermediate).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# The contract is ONE JSON line on stdout.  Libraries print there too (NCCL writes its version banner
# to fd 1), so fd 1 is pointed at stderr for the whole run and the JSON goes to the saved descriptor.
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(line: dict) -> None:
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


BODIES_PER_GPU = 4096
BYTES_E2E = 83_500          # per body: 340 in + 83,160 out
BYTES_K1 = 83_020           # write vposed 82,680 + read coefficients 340
BYTES_K2 = 2_100
BYTES_K3 = 166_992          # read vposed 82,680 + A 1,152; write verts 82,680 + joints/kp2d 480
BYTES_FUSED = 84_652        # fused blendshapes+LBS: read coef 340 + A 1,152, write verts 82,680 + joints/kp2d 480
FLOPS_K1 = 8_970_780        # 2 * 217 * 20670 (algorithmic; the MMA executes K = 224)
FLOPS_BODY = 14_080_000     # whole forward, SURVEY.md §8d
NCU_SUMMARY = "r02_ncu_summary.json"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained"), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


def ncu_traffic(kernel: str):
    """dram bytes (read+write) per launch of `kernel` from the committed `ncu --set full` summary of this
    round (bench.py cannot run under ncu; the capture is of this same command, see profiles/README.md)."""
    for name in (NCU_SUMMARY, "r01_ncu_summary.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                return json.load(f)["kernels"][kernel]["dram_bytes_per_launch"], f"profiles/{name}"
        except (OSError, KeyError, ValueError):
            continue
    return None, None


def bench_config(n: int, world: int, args, strong: bool) -> dict:
    """The `config` object: BOTH arms print exactly this (same workload, same batch per step)."""
    return {
        "workload": f"SMPL forward + 2D keypoint projection, batch {n} per GPU "
                    f"(BASELINE.json configs[2]), synthetic SMPL-shaped model seed 0 "
                    f"(6890 verts, 24 joints, 10 betas, 207 posedirs, {args.weights} weights), "
                    f"outputs vertices[N,6890,3] + joints[N,24,3] + kp2d[N,24,2] fp32",
        "bodies_per_step_per_gpu": n,
        "inputs": "synthetic seed 1+rank: betas N(0,1) clipped +-3, pose N(0,0.3^2) rad (+ Rodrigues edge rows), cam",
        "parallelism": f"batch-sharded x{world}, no data-path collective"
                       + ("; joints+kp2d rows of every rank exchanged each step (value and e2e legs alike)" if world > 1 else ""),
        "scaling_mode": "strong" if strong else "weak",
        "l2": "per step >= 0.34 GB of vertices (+ the vposed intermediate) stream through HBM >> 126 MB L2: every "
              "timed iteration runs on inputs/outputs larger than L2; model tensors (19 MB) are L2-resident by "
              "design and excluded from algorithmic bytes",
    }


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled in the background with host timestamps."""
    FIELDS = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int, period_ms: int = 20):
        self.samples = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", str(period_ms), "-i", str(index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        import datetime
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) >= 8:
                try:  # nvidia-smi's own timestamp: the pipe is block-buffered, read time is useless
                    t = datetime.datetime.strptime(parts[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                except ValueError:
                    t = time.time()
                self.samples.append((t, parts[1:]))

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self, t0: float, t1: float) -> dict:
        inside = [p for (t, p) in self.samples if t0 <= t <= t1]
        note = "sampled inside the timed region"
        if len(inside) < 3:
            inside = [p for (_, p) in self.samples]
            note = "timed region shorter than 3 sampling periods: samples span warm-up + timed + kernel loops"
        if not inside:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "note": "nvidia-smi unavailable"}
        mhz, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for p in inside:
            try:
                mhz.append(float(p[0])); mx.append(float(p[1]))
            except ValueError:
                continue
            for name, flag in zip(names, p[3:7]):
                if flag.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(mhz) if mhz else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(inside), "note": note}


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_oracle_throughput(model, bodies: int, warm: int, timed: int, seed: int = 1, min_seconds: float = 0.0,
                          chunk: int = 512):
    """bodies/s of the CPU oracle (fp32, all host threads) on `bodies` bodies per step (chunks of `chunk`)."""
    import torch
    from human_3d_reconstruction_b200 import synthetic
    from oracle.smpl_ref import smpl_forward_chunked
    cores = host_cores()
    torch.set_num_threads(cores)
    betas, pose, cam = synthetic.make_inputs(bodies, seed)
    tm = {k: torch.as_tensor(v) for k, v in model.items()}
    times = []
    with torch.no_grad():
        i = 0
        while i < warm + timed or sum(times) < min_seconds:
            t0 = time.perf_counter()
            smpl_forward_chunked(tm, betas, pose, cam, chunk=chunk, dtype=torch.float32)
            dt = time.perf_counter() - t0
            if i >= warm:
                times.append(dt)
            i += 1
    return bodies / statistics.median(times), cores, times


def run_reference(args):
    """Reference arm: the oracle port of the eager PyTorch SMPL layer on the host cores, same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import torch
    from human_3d_reconstruction_b200 import synthetic
    world = int(os.environ.get("WORLD_SIZE", "1"))
    model = synthetic.make_model(0, weights=args.weights)
    n = args.bodies_per_gpu
    # the full batch per step unless warm-up + steps would run past ~4 minutes (then a bounded sample)
    bps, cores, _ = cpu_oracle_throughput(model, 256, 1, 2)
    budget_s = 240.0
    fit = int(bps * budget_s / max(1, args.steps + args.warmup))
    bodies = n if fit >= n else max(512, fit // 512 * 512)
    _, cores, times = cpu_oracle_throughput(model, bodies, args.warmup, args.steps)
    total = sum(times)
    value = bodies * len(times) / total
    sample = (f"{bodies} bodies per step" + (" (the whole batch)" if bodies == n else f" (bounded sample of the {n}-body batch)")
              + f" in 512-body chunks, {len(times)} timed steps, oracle/smpl_ref.py, torch {torch.__version__} fp32, {cores} threads")
    line = {
        "impl": "reference", "metric": "smpl_forward_bodies_per_sec", "value": value, "unit": "bodies/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": bench_config(n, world, args, False),
        "note": "reference snapshot has no SMPL layer (SURVEY.md F1): arm = CPU oracle port of the eager PyTorch layer",
        "cpu_baseline": {"value": value, "unit": "bodies/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "bodies/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


def time_loop(fn, iters: int, torch):
    """CUDA-event time of `iters` back-to-back calls on the current stream, in seconds."""
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    start.record()
    for _ in range(iters):
        fn()
    end.record()
    torch.cuda.synchronize()
    return start.elapsed_time(end) * 1e-3


def measure_peaks(torch, capi, dev):
    """TF32 / bf16 tensor throughput (cuBLAS, 8192^3) and fp32 FMA throughput (own probe kernel), in this run."""
    out = {}
    m = 8192
    for name, dt_, tf32 in (("tf32_tflops", torch.float32, True), ("bf16_tflops", torch.bfloat16, False)):
        a = torch.randn(m, m, device=dev, dtype=dt_)
        b = torch.randn(m, m, device=dev, dtype=dt_)
        prev = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        try:
            for _ in range(3):
                torch.matmul(a, b)
            best = min(time_loop(lambda: torch.matmul(a, b), 3, torch) / 3 for _ in range(4))
        finally:
            torch.backends.cuda.matmul.allow_tf32 = prev
        out[name] = 2.0 * m ** 3 / best * 1e-12
        del a, b
    import ctypes as C
    scratch = torch.zeros(64, device=dev)
    flop = C.c_double()
    lib = capi.lib()
    s = torch.cuda.current_stream(dev).cuda_stream

    def probe():
        capi.check(lib.smplb200_probe_fp32_fma(dev.index, 20000, scratch.data_ptr(), C.byref(flop), s), "probe")

    probe()
    best = min(time_loop(probe, 2, torch) / 2 for _ in range(3))
    out["fp32_fma_tflops"] = flop.value / best * 1e-12
    out["how"] = ("torch.matmul 8192^3 (cuBLAS; allow_tf32 for the fp32 operands), best of 4 x 3; fp32 FMA: "
                  "smplb200_probe_fp32_fma (8 CTAs/SM x 256 threads x 8 independent FMA chains), best of 3 x 2")
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--bodies-per-gpu", type=int, default=BODIES_PER_GPU)
    ap.add_argument("--total-bodies", type=int, default=0,
                    help="strong scaling: fix the whole job's batch (SURVEY C4: 65536) and shard it over the ranks")
    ap.add_argument("--precision", default="f16", choices=["fp32", "bf16", "tf32", "bf16x3", "f16x3", "auto", "f16"],
                    help="blendshape MMA operands.  f16 (default) = the FUSED blendshapes+skinning kernel, fp16 operands "
                         "(BASELINE configs[2] is the reduced-precision tensor-core regime: 'TF32/BF16 blendshapes ... fp32 "
                         "LBS'; f16 is 10x / 70x more accurate than those, error measured in the run); auto = f16x3, the "
                         "near-fp32 split-fp16 mode (unfused kernels); bf16x3 = round 1's split-bf16 mode")
    ap.add_argument("--lbs", default="auto", choices=["fma", "tc", "dense", "auto"])
    ap.add_argument("--weights", default="sparse", choices=["sparse", "dense"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip variants / configs_extra / next rows (profiling runs)")
    ap.add_argument("--exchange", default="auto", choices=["auto", "peer", "dma", "collective", "off"])
    ap.add_argument("--kernel-iters", type=int, default=50)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from human_3d_reconstruction_b200 import SMPL, GraphedSMPL, StaticSMPL, capi, synthetic, sharding
    from human_3d_reconstruction_b200 import smpl as ops
    from human_3d_reconstruction_b200.smpl import HostRunner

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback for the product path)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    n = args.bodies_per_gpu
    strong = args.total_bodies > 0
    if strong:      # contiguous shards of a fixed job
        if args.total_bodies % world:
            raise SystemExit("--total-bodies must be divisible by the number of ranks")
        n = sharding.shard_bounds(args.total_bodies, world, 0)[1]
    peaks = load_peaks()
    extras = rank == 0 and not strong and not args.no_extras

    model = synthetic.make_model(0, weights=args.weights)
    layer = SMPL(model, precision=args.precision, lbs=args.lbs).to(dev)
    betas, pose, cam = synthetic.make_inputs(n, 1 + rank)
    tb, tp, tc = (torch.from_numpy(x).to(dev) for x in (betas, pose, cam))
    n_total = n * world
    exchange = None
    if world > 1 and args.exchange != "off":
        exchange = sharding.PeerExchange(n_total, dev, transport=args.exchange)

    def new_ready():
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(dev))       # creates the CUDA handle the library re-records
        return ev

    # two batches in flight on two streams: a serving loop keeps the GPU busy across the
    # kernel-to-kernel bubbles of one forward (same structure as the e2e leg below).  Each slot is a
    # StaticSMPL runner (static device buffers, one C call per step): with the general SMPL.forward the
    # loop is host-bound at this step time (0.14 ms).
    streams = [torch.cuda.Stream(device=dev) for _ in range(2)]
    runners, readies = [], []
    for q in range(2):
        bq, pq, cq = synthetic.make_inputs(n, 1 + rank + 100 * q)
        r = StaticSMPL(layer, n, dev)
        r.betas.copy_(torch.from_numpy(bq)); r.pose.copy_(torch.from_numpy(pq)); r.cam.copy_(torch.from_numpy(cq))
        runners.append(r)
        readies.append(new_ready())
    inputs = [(r.betas, r.pose, r.cam) for r in runners]
    counter = {"i": 0}
    last = {}

    def step():
        q = counter["i"] & 1
        counter["i"] += 1
        if exchange is None:
            last["out"] = runners[q].run(stream=streams[q])
            return
        v, j, k = runners[q].run(stream=streams[q], joints_ready=readies[q])
        # every rank's joints + kp2d on every rank (configs[3]); vertices stay sharded.  Runs on the
        # exchange's side stream as soon as k2 has produced the rows -- never on the compute stream.
        last["out"] = (v,) + exchange.exchange(j, k, readies[q])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def join_streams(main, strs):
        for st in strs:
            main.wait_stream(st)
        if exchange is not None:
            main.wait_stream(exchange.stream)

    sampler = ClockSampler(local_rank) if rank == 0 else None
    with torch.no_grad():
        for _ in range(args.warmup):
            step()
        barrier()
        t_wall0 = time.time()
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        main_stream = torch.cuda.current_stream(dev)
        start.record(main_stream)
        for st in streams:
            st.wait_stream(main_stream)
        for _ in range(args.steps):
            step()
        join_streams(main_stream, streams)
        end.record(main_stream)
        barrier()
        t_wall1 = time.time()
        elapsed = start.elapsed_time(end) * 1e-3
        if world > 1:
            t = torch.tensor([elapsed], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            elapsed = float(t.item())
        value = n_total * args.steps / elapsed
        if exchange is not None:      # the gathered rows really are everyone's: check this rank's view once
            _, j_all, k_all, _ = last["out"]
            q = (counter["i"] - 1) & 1
            ref = layer(*inputs[q])
            torch.cuda.synchronize()
            lo = rank * n
            assert torch.equal(j_all[lo:lo + n], ref[1]) and torch.equal(k_all[lo:lo + n], ref[2]), "exchange: own rows"
            chk = torch.stack([j_all.double().sum(), k_all.double().sum()])
            allc = [torch.empty_like(chk) for _ in range(world)]
            dist.all_gather(allc, chk)
            assert all(torch.equal(c, allc[0]) for c in allc), "exchange: ranks disagree on the gathered rows"

        # ---- e2e through the C-ABI host entry point (pinned host buffers) -----------------
        def e2e_rate(with_vertices: bool, iters: int):
            # two runners on two streams, alternated: every step still does its own H2D of the inputs
            # and D2H of the results, but step i+1's copies overlap step i's kernels (what a serving
            # loop over smplb200_forward_host does with two staging arenas)
            runners = [HostRunner(layer, n, dev, with_vertices=with_vertices, with_cam=True) for _ in range(2)]
            strs = [torch.cuda.Stream(device=dev) for _ in range(2)]
            evs = [new_ready() for _ in range(2)]
            for r in runners:
                r.betas.copy_(torch.from_numpy(betas)); r.pose.copy_(torch.from_numpy(pose))
                r.cam.copy_(torch.from_numpy(cam))
            state = {"i": 0}

            def one():
                k = state["i"] & 1
                state["i"] += 1
                if exchange is None:
                    runners[k].run(stream=strs[k])
                else:       # same exchange as the `value` leg, fed from the staging arena's device rows
                    runners[k].run(stream=strs[k], joints_ready=evs[k])
                    exchange.exchange(runners[k].joints_dev, runners[k].kp2d_dev, evs[k])

            for _ in range(4):
                one()
            barrier()
            main = torch.cuda.current_stream(dev)
            start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            start.record(main)
            for st in strs:
                st.wait_stream(main)
            for _ in range(iters):
                one()
            join_streams(main, strs)
            end.record(main)
            torch.cuda.synchronize()
            dt = start.elapsed_time(end) * 1e-3
            runner = runners[0]
            if world > 1:
                t = torch.tensor([dt], device=dev, dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dt = float(t.item())
            return n_total * iters / dt, runner.h2d_bytes, runner.d2h_bytes, dt / iters

        e2e_val, h2d, d2h, e2e_dt = e2e_rate(True, max(5, min(args.steps, 20)))
        e2e_s_val, h2d_s, d2h_s, _ = e2e_rate(False, max(10, min(args.steps, 200)))

        # ---- per-kernel timing (rank 0) for the roofline ------------------------------------
        kernels = {}
        fused = layer.flags & capi.PREC_MASK == capi.PREC_F16
        if rank == 0:
            it = args.kernel_iters
            lib = capi.lib()
            s = torch.cuda.current_stream(dev).cuda_stream
            # the unfused tensor-core kernels (k1 split-fp16, k3 3xTF32): timed for every run so the table is complete
            ulayer = layer if not fused else SMPL(model, precision="f16x3", lbs="tc").to(dev)
            flags = ulayer.flags
            coef, A, joints = ops.pose_chain(ulayer, tb, tp)
            vposed = ops.blendshapes(ulayer, coef, flags=flags)
            h = ulayer.handle(dev)
            verts = torch.empty((n, layer.num_verts, 3), device=dev)
            kp = torch.empty((n, 24, 2), device=dev)
            ws = torch.empty(max(h.workspace_bytes(n, flags), 256), dtype=torch.uint8, device=dev)
            wsb = int(lib.smplb200_blendshapes_workspace_bytes(h.ptr, n, flags))
            wsl = int(lib.smplb200_lbs_workspace_bytes(h.ptr, n, flags))
            wsf = int(lib.smplb200_blend_skin_workspace_bytes(h.ptr, n))
            wsf_t = torch.empty(max(wsf, 256), dtype=torch.uint8, device=dev)

            def k2():
                capi.check(lib.smplb200_pose_chain(h.ptr, tb.data_ptr(), tp.data_ptr(), n, coef.data_ptr(),
                                                   A.data_ptr(), joints.data_ptr(), flags, s), "k2")

            def k1():  # note: the stand-alone entry also runs the small operand pack kernel
                capi.check(lib.smplb200_blendshapes(h.ptr, coef.data_ptr(), n, vposed.data_ptr(), ws.data_ptr(),
                                                    wsb, flags, s), "k1")

            def k3():
                capi.check(lib.smplb200_lbs(h.ptr, vposed.data_ptr(), A.data_ptr(), n, verts.data_ptr(),
                                            joints.data_ptr(), tc.data_ptr(), kp.data_ptr(), ws.data_ptr(),
                                            wsl, flags, s), "k3")

            def kf():  # fused kernel ALONE: operand images packed once below, reused (coef = A = NULL)
                capi.check(lib.smplb200_blend_skin(h.ptr, None, None, n, verts.data_ptr(), wsf_t.data_ptr(), wsf, s), "fused")

            def whole():    # the forward alone on ONE stream (no second batch in flight)
                layer(tb, tp, tc)

            todo = [("k2_pose_chain", k2, BYTES_K2), ("k1_blendshapes", k1, BYTES_K1), ("k3_lbs", k3, BYTES_K3),
                    ("forward_one_stream", whole, BYTES_E2E)]
            if wsf:
                capi.check(lib.smplb200_blend_skin(h.ptr, coef.data_ptr(), A.data_ptr(), n, verts.data_ptr(),
                                                   wsf_t.data_ptr(), wsf, s), "fused (pack)")
                todo.insert(3, ("fused_blend_skin", kf, BYTES_E2E))
            for name, fn, byts in todo:
                for _ in range(3):
                    fn()
                dt = time_loop(fn, it, torch) / it
                kernels[name] = {"us": dt * 1e6, "gbs": byts * n / dt * 1e-9}

        # ---- accuracy of every precision mode vs the fp32 CPU oracle, measured HERE ---------------
        accuracy, variants = {}, {}
        if extras:
            from oracle.smpl_ref import smpl_forward_chunked
            na = min(n, 512)
            ref = smpl_forward_chunked(model, betas[:na], pose[:na], cam[:na], chunk=256, dtype=torch.float32)
            bound = {"fp32": 1e-6, "f16x3": 4e-6, "bf16x3": 1e-5, "auto": 1e-5, "f16": 5e-5, "tf32": 5e-4, "bf16": 4e-3}
            for prec in ("f16", "auto", "fp32", "f16x3", "bf16x3", "tf32", "bf16"):
                lay = layer if prec == args.precision else SMPL(model, precision=prec, lbs=args.lbs if prec != "f16" else "auto").to(dev)
                v, j, k = lay(tb[:na], tp[:na], tc[:na])
                ev = (v.cpu() - ref[0]).abs().max().item()
                ej = (j.cpu() - ref[1]).abs().max().item()
                ek = (k.cpu() - ref[2]).abs().max().item()
                accuracy[prec] = {"vertices_max_abs_err_m": ev, "joints_max_abs_err_m": ej, "kp2d_max_abs_err": ek,
                                  "stated_vertex_bound_m": bound[prec] if prec != "fp32" else "rtol 1e-5 / atol 1e-6"}
                okv = torch.allclose(v.cpu(), ref[0], rtol=1e-5, atol=bound[prec])
                okj = torch.allclose(j.cpu(), ref[1], rtol=1e-5, atol=1e-6) and torch.allclose(k.cpu(), ref[2], rtol=1e-5, atol=2e-6)
                accuracy[prec]["within_bound"] = bool(okv and okj)
                assert okv and okj, f"parity check failed in the bench run for precision {prec}: {accuracy[prec]}"
                if prec in (args.precision,):
                    continue
                # same workload with these operands (short loop)
                for _ in range(3):
                    lay(tb, tp, tc)
                it = 10 if prec == "fp32" else 50
                dt = time_loop(lambda: lay(tb, tp, tc), it, torch) / it
                variants[prec] = {"bodies_per_s": n / dt, "us_per_step": dt * 1e6}
            accuracy["how"] = f"first {na} bodies of the timed batch vs oracle/smpl_ref.py (fp32, CPU), max abs error, this run"

        # ---- peaks measured in this run -------------------------------------------------------------
        run_peaks = measure_peaks(torch, capi, dev) if rank == 0 and not strong else None

        # ---- configs_extra: BASELINE.json configs[1] and configs[4] --------------------------------
        configs_extra, next_rows = {}, {}
        if extras:
            configs_extra["n64_fp32"] = bench_n64(torch, SMPL, GraphedSMPL, synthetic, model, dev, run_peaks)
            configs_extra["dla34_decode_smpl"] = bench_config5(torch, SMPL, synthetic, model, dev)
            next_rows = bench_next_rows(torch, SMPL, capi, synthetic, model, dev, n, tb, tp, tc)

    clocks = sampler.summary(t_wall0, t_wall1) if sampler else None
    if sampler:
        sampler.stop()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        bps, cores, times = cpu_oracle_throughput(model, 1024, 1, 3, min_seconds=10.0)
        cpu_model = ""
        try:
            with open("/proc/cpuinfo") as fh:
                cpu_model = next((ln.split(":", 1)[1].strip() for ln in fh if ln.startswith("model name")), "")
        except OSError:
            pass
        # single-body latency of the eager layer on the host (all threads, then one thread)
        import torch as _t
        from oracle.smpl_ref import smpl_forward as _oracle_fwd
        b1_, p1_, c1_ = synthetic.make_inputs(1, 3)

        def _lat():
            for _ in range(3):
                _oracle_fwd(model, b1_, p1_, c1_)
            t0 = time.perf_counter()
            for _ in range(20):
                _oracle_fwd(model, b1_, p1_, c1_)
            return (time.perf_counter() - t0) / 20 * 1e6

        with _t.no_grad():
            lat_all = _lat()
            nt_prev = _t.get_num_threads()
            _t.set_num_threads(1)
            lat_one = _lat()
            _t.set_num_threads(nt_prev)
        cpu = {"value": bps, "unit": "bodies/s", "cores": cores, "kind": "port",
               "cpu_model": cpu_model, "os_cpu_count": os.cpu_count(), "affinity": host_cores(),
               "n1_latency_us": {"all_threads": lat_all, "one_thread": lat_one},
               "sample": f"oracle/smpl_ref.py fp32, 1024-body steps in 512-body chunks (4 steps = the 4096-body workload), "
                         f"1 warm-up + {len(times)} timed steps (median), {sum(times):.1f} s of CPU work"}

    if rank == 0:
        k3 = kernels["k3_lbs"]; k1 = kernels["k1_blendshapes"]
        hbm = peaks["hbm_gbs"]
        tf_k1 = FLOPS_K1 * n / (k1["us"] * 1e-6) * 1e-12
        k1_roof = {"bound": "hbm", "achieved": k1["gbs"], "peak": hbm, "unit": "GB/s",
                   "frac": k1["gbs"] / hbm, "us": k1["us"], "bytes_per_body": BYTES_K1, "operands": "f16x3 (unfused path, what 'auto' runs)",
                   "tensor_tflops": tf_k1, "tensor_frac_of_bf16_burst": tf_k1 / peaks["bf16_tflops"],
                   "note": "K=217: write-bound; algorithmic flops (the split modes execute 2-3x as many MMAs)"}
        if run_peaks:
            k1_roof["tensor_frac_of_tf32_peak"] = tf_k1 / run_peaks["tf32_tflops"]
            k1_roof["tensor_frac_of_bf16_peak_this_run"] = tf_k1 / run_peaks["bf16_tflops"]
        one = kernels["forward_one_stream"]
        roof_k = {
            "k3_lbs": {"bound": "hbm", "achieved": k3["gbs"], "peak": hbm, "unit": "GB/s",
                       "frac": k3["gbs"] / hbm, "us": k3["us"], "bytes_per_body": BYTES_K3, "note": "unfused path (k_lbs_tc)"},
            "k1_blendshapes": k1_roof,
            "k2_pose_chain": {"bound": "latency", "us": kernels["k2_pose_chain"]["us"],
                              "achieved": kernels["k2_pose_chain"]["gbs"], "unit": "GB/s"},
            "forward_one_stream": {"bound": "hbm", "achieved": one["gbs"], "peak": hbm, "unit": "GB/s",
                                   "frac": one["gbs"] / hbm, "us": one["us"], "bytes_per_body": BYTES_E2E,
                                   "note": "one smplb200_forward after another on ONE stream"},
            "whole_step": {"bound": "hbm", "achieved": BYTES_E2E * value / world * 1e-9, "peak": hbm,
                           "unit": "GB/s", "frac": BYTES_E2E * value / world * 1e-9 / hbm,
                           "bytes_per_body": BYTES_E2E, "note": "the `value` leg (two batches in flight)"},
        }
        if "fused_blend_skin" in kernels:
            kf_ = kernels["fused_blend_skin"]
            tf_f = (FLOPS_K1 + 3_968_640) * n / (kf_["us"] * 1e-6) * 1e-12
            roof_k["fused_blend_skin"] = {
                "bound": "hbm", "achieved": kf_["gbs"], "peak": hbm, "unit": "GB/s", "frac": kf_["gbs"] / hbm,
                "us": kf_["us"], "bytes_per_body": BYTES_E2E,
                "tensor_tflops_algorithmic": tf_f,
                "note": "k_fused_tc alone (blendshapes + skinning in one kernel); algorithmic bytes = the whole step's "
                        "83,500 B/body (340 in + 83,160 out): the kernel has no other mandatory HBM traffic"}
            if run_peaks:
                roof_k["fused_blend_skin"]["tensor_frac_of_bf16_peak_this_run"] = tf_f / run_peaks["bf16_tflops"]
        dom = "fused_blend_skin" if fused else "k3_lbs"
        dom_kernel = "k_fused_tc" if fused else ("k_lbs_tc" if n >= capi.TC_LBS_MIN_BATCH else "k_lbs_fma")
        dom_bytes = BYTES_E2E if fused else BYTES_K3
        launches_per_step = layer.launch_count(n, True, dev) + (2 if (exchange is not None and exchange.transport == "peer") else 0)   # 'dma' exchange: no kernel
        traffic, traffic_src = ncu_traffic(dom_kernel) if n == BODIES_PER_GPU else (None, None)
        pcie_gbs = (h2d + d2h) / e2e_dt * 1e-9
        config = bench_config(n, world, args, strong)
        line = {
            "metric": "smpl_forward_bodies_per_sec", "value": value, "unit": "bodies/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * elapsed / args.steps, "higher_is_better": True,
            "scaling": "strong" if strong else "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config,
            "impl_config": {
                "blendshape_operands": args.precision, "lbs": args.lbs, "accumulate": "fp32",
                "resolved_flags": int(layer.flags),
                "streams": "2 batches in flight on 2 CUDA streams (value and e2e legs); per-kernel roofline "
                           "times are single-stream, back to back",
                "exchange": None if exchange is None else
                            {"transport": exchange.transport, "why_not_peer": exchange.why_not_peer,
                             "what": "joints+kp2d rows of all ranks on every rank, on a side stream behind the "
                                     "joints-ready event of each forward"},
            },
            "e2e": {"value": e2e_val, "unit": "bodies/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "pcie_gbs_per_gpu": pcie_gbs,
                    "note": "smplb200_forward_host: pinned host betas/pose/cam H2D, forward, D2H of vertices + joints + kp2d "
                            "(the whole result); PCIe-bound"},
            "e2e_small_outputs": {"value": e2e_s_val, "unit": "bodies/s", "h2d_bytes_per_step": h2d_s,
                                  "d2h_bytes_per_step": d2h_s,
                                  "note": "same call returning joints + kp2d only; vertices stay device-resident"},
            "gpu_launches": launches_per_step * args.steps,
            "gpu_launches_per_step": launches_per_step,
            "roofline": {**roof_k[dom], "kernel": dom_kernel,
                         "peak_source": peaks["source"] + " (MEASURED_PEAKS.json hbm_gbs)",
                         "algorithmic_bytes_per_launch": dom_bytes * n,
                         "traffic": traffic, "traffic_source": traffic_src},
            "peaks_this_run": run_peaks,
            "variants_same_workload": variants,
            "accuracy": accuracy,
            "configs_extra": configs_extra,
            "next_rows": next_rows,
            "roofline_kernels": roof_k,
            "cpu_baseline": cpu,
            "clocks": clocks,
        }
        emit(line)
    if world > 1:
        # rank 0 is still busy with the extras while the others are done: nobody unmaps the symmetric exchange
        # buffers (or leaves the process group) before everyone has arrived here
        dist.barrier()
        exchange = None
        dist.destroy_process_group()
    return 0


# ---------------------------------------------------------------------------------------------------
def bench_n64(torch, SMPL, GraphedSMPL, synthetic, model, dev, run_peaks):
    """BASELINE.json configs[1]: 64 bodies (one image's detected people), fp32 FMA regime, CUDA-graph replay."""
    from oracle.smpl_ref import smpl_forward
    n = 64
    b, p, c = synthetic.make_inputs(n, 21)
    ref = smpl_forward(model, b, p, c, dtype=torch.float32)
    out = {"workload": "SMPL forward batch 64 (BASELINE.json configs[1]), CUDA-graph replay of one smplb200_forward, L2-warm",
           "modes": "fp32 = FMA blendshapes + FMA skinning (the config's regime); auto = split-fp16 tcgen05 blendshapes + FMA "
                    "skinning (stated <= 1e-5 m, measured ~2e-6); f16 = the fused tcgen05 kernel (<= 5e-5 m)"}
    for name, kw in (("fp32", dict(precision="fp32", lbs="auto")), ("auto", dict(precision="auto", lbs="auto")),
                     ("f16", dict(precision="f16"))):
        lay = SMPL(model, **kw).to(dev)
        g = GraphedSMPL(lay, n, dev)
        g.betas.copy_(torch.from_numpy(b)); g.pose.copy_(torch.from_numpy(p)); g.cam.copy_(torch.from_numpy(c))
        for _ in range(5):
            g.replay()
        dt = min(time_loop(g.replay, 200, torch) / 200 for _ in range(3))
        v, j, k = g.replay()
        torch.cuda.synchronize()
        atol = {"fp32": 1e-6, "auto": 1e-5, "f16": 5e-5}[name]
        ok = (torch.allclose(v.cpu(), ref[0], rtol=1e-5, atol=atol) and torch.allclose(j.cpu(), ref[1], rtol=1e-5, atol=1e-6)
              and torch.allclose(k.cpu(), ref[2], rtol=1e-5, atol=2e-6))
        assert ok, f"configs[1] parity failed ({name})"
        tfl = FLOPS_BODY * n / dt * 1e-12
        out[name] = {"us_per_call": dt * 1e6, "bodies_per_s": n / dt, "tflops": tfl,
                     "vertices_max_abs_err_m": (v.cpu() - ref[0]).abs().max().item(), "parity": "asserted in this run"}
        if name == "fp32" and run_peaks:
            out[name]["frac_of_fp32_fma_peak"] = tfl / run_peaks["fp32_fma_tflops"]
    return out


def bench_config5(torch, SMPL, synthetic, model, dev):
    """BASELINE.json configs[4]: random-init DLA-34 (seed 317) at 512x512, batch 32 -> decode (K=32) -> SMPL."""
    from human_3d_reconstruction_b200 import DCN, decode_gather
    from oracle.decode_ref import check_equivalent, decode_gather as decode_cpu
    from oracle.dla34_ref import HEADS_HMR, calibrate_batchnorm, dla_net
    from oracle.smpl_ref import smpl_forward_chunked
    B, K = 32, 32
    g = torch.Generator().manual_seed(11)
    x = torch.randn(B, 3, 512, 512, generator=g).to(dev)
    small = SMPL(model).to(dev)
    out = {"workload": "reference dla_net(heads hm1/wh2/reg2/pose72/shape10/cam3) restated in oracle/dla34_ref.py (pinned to the "
                       "unmodified reference), random init seed 317, eval, 512x512, batch 32 -> sigmoid(hm) -> decode_gather K=32 "
                       "-> SMPL (auto) for 1024 people; BatchNorm statistics calibrated on the batch for the timed / bit-exact "
                       "variant (the raw eval-mode random init collapses the heat map into exact ties)",
           "backbone": "PyTorch/cuDNN (test infrastructure: producer of the head maps, not the product)"}

    def pipeline(net):
        o = net(x)[0]
        hm = torch.sigmoid(o["hm"])
        dec = decode_gather(hm, [o["pose"], o["shape"], o["cam"]], K)
        po, be, ca = (t.reshape(B * K, -1) for t in dec[5])
        return o, hm, dec, small(be, po, ca)

    def timed(fn, it=5):
        for _ in range(2):
            fn()
        return time_loop(fn, it, torch) / it

    net = dla_net(dict(HEADS_HMR), seed=317).eval().to(dev)
    # raw random-init network (eval-mode BatchNorm with default statistics): the heat map collapses to its
    # bias and ties by the thousand -- the decode must still be a valid reference result (up to tie order)
    o, hm, dec, _ = pipeline(net)
    dec_cpu = tuple(t.cpu() for t in dec[:5]) + ([t.cpu() for t in dec[5]],)
    ok_raw, ties_raw = check_equivalent(dec_cpu, hm.cpu(), [o["pose"].cpu(), o["shape"].cpu(), o["cam"].cpu()], K)
    assert ok_raw, "configs[4]: decode on the raw random-init head maps is not a valid reference result"
    # BatchNorm statistics calibrated on the batch (weights stay the seeded random init): distinct peaks
    calibrate_batchnorm(net, x)
    o, hm, dec, (v, j, k) = pipeline(net)
    # parity, asserted in the run: decode bit-exact vs the reference functions' port on these very head maps,
    # meshes vs the CPU oracle
    refd = decode_cpu(hm.cpu(), [o["pose"].cpu(), o["shape"].cpu(), o["cam"].cpu()], K)
    same_scores = torch.equal(dec[0].cpu(), refd[0])
    same_rest = all(torch.equal(a.cpu(), b_) for a, b_ in zip(dec[1:5], refd[1:5])) and \
        all(torch.equal(a.cpu(), b_) for a, b_ in zip(dec[5], refd[5]))
    assert same_scores and same_rest, "configs[4]: decode differs from the reference port"
    po, be, ca = (t.reshape(B * K, -1) for t in dec[5])
    rv, rj, rk = smpl_forward_chunked(model, be.cpu().numpy(), po.cpu().numpy(), ca.cpu().numpy(), chunk=256)
    ev, ej = (v.cpu() - rv).abs().max().item(), (j.cpu() - rj).abs().max().item()
    assert torch.allclose(v.cpu(), rv, rtol=1e-5, atol=1e-5) and torch.allclose(j.cpu(), rj, rtol=1e-5, atol=1e-6), \
        "configs[4]: meshes differ from the oracle"
    t_net = timed(lambda: net(x))
    t_dec = timed(lambda: decode_gather(hm, [o["pose"], o["shape"], o["cam"]], K), 50)
    t_smpl = timed(lambda: small(be, po, ca), 50)
    t_all = timed(lambda: pipeline(net))
    out["not_use_dcn"] = {"backbone_ms": t_net * 1e3, "decode_gather_us": t_dec * 1e6, "smpl_1024_bodies_us": t_smpl * 1e6,
                          "end_to_end_ms": t_all * 1e3, "people_per_s": B * K / t_all, "images_per_s": B / t_all,
                          "parity": {"decode_scores_bit_exact": bool(same_scores), "decode_inds_and_vectors_bit_exact": bool(same_rest),
                                     "raw_random_init_valid_up_to_tie_order": bool(ok_raw), "raw_random_init_images_with_ties": ties_raw,
                                     "vertices_max_abs_err_m": ev, "joints_max_abs_err_m": ej, "asserted": True}}
    del net

    # upstream's USE_DCN=True: the neck's 16 DeformConvs through the product's DCN module (k_dcn_fwd)
    def deform(ci, co):
        return DCN(ci, co, kernel_size=(3, 3), stride=1, padding=1, dilation=1, deformable_groups=1)

    netd = dla_net(dict(HEADS_HMR), seed=317, deform=deform).eval()
    gg = torch.Generator().manual_seed(317)
    for m in netd.modules():
        if isinstance(m, DCN):      # upstream zero-initialises the offset conv: make the layers actually deform
            m.conv_offset_mask.weight.data.normal_(0.0, 0.6 / (m.in_channels * 9) ** 0.5, generator=gg)
            m.conv_offset_mask.bias.data.normal_(0.0, 0.5, generator=gg)
    netd = netd.to(dev)
    t_netd = timed(lambda: netd(x))
    t_alld = timed(lambda: pipeline(netd))
    out["use_dcn"] = {"backbone_ms": t_netd * 1e3, "end_to_end_ms": t_alld * 1e3, "people_per_s": B * K / t_alld,
                      "note": "per-layer parity of the 16 DCN layers vs oracle/dcn_ref is in tests/test_gpu_config5.py"}
    return out


def bench_next_rows(torch, SMPL, capi, synthetic, model, dev, n, tb, tp, tc):
    """the backward pass and the DCNv2 layers (decode is in configs_extra.dla34_decode_smpl)."""
    next_rows = {}
    from oracle.smpl_ref import smpl_forward as oracle_forward
    nt = 128
    lay_t = SMPL(model, precision="auto", lbs="auto").to(dev)
    arrs = synthetic.make_inputs(nt, 9)
    bt, pt, ct = (torch.from_numpy(x).to(dev).requires_grad_() for x in arrs)

    def loss_of(outs, with_verts):
        v, j, k = outs
        l = k.abs().mean() + j.pow(2).mean()
        return l + v.pow(2).mean() if with_verts else l

    def train_step(with_verts):
        bt.grad = pt.grad = ct.grad = None
        with torch.enable_grad():
            loss_of(lay_t(bt, pt, ct), with_verts).backward()

    bw = {}
    for name, wv in (("loss_on_joints_kp2d", False), ("loss_on_vertices_joints_kp2d", True)):
        for _ in range(3):
            train_step(wv)
        dt_t = time_loop(lambda: train_step(wv), 20, torch) / 20
        cb, cp, cc = (torch.from_numpy(x).requires_grad_() for x in arrs)
        t0 = time.perf_counter()
        with torch.enable_grad():
            loss_of(oracle_forward(model, cb, cp, cc), wv).backward()
        dt_c = time.perf_counter() - t0
        bw[name] = {"gpu_fwd_bwd_us": dt_t * 1e6, "cpu_autograd_port_us": dt_c * 1e6}
    # the backward call alone at the headline batch (vertex path, device-resident gradients)
    hb = lay_t.handle(dev)
    lib_ = capi.lib()
    wsb = int(lib_.smplb200_backward_workspace_bytes(hb.ptr, n, lay_t.flags, 1))
    wsb_t = torch.empty(wsb, dtype=torch.uint8, device=dev)
    gv_, gj_, gk_ = (torch.randn(n, d0, d1, device=dev) for d0, d1 in ((6890, 3), (24, 3), (24, 2)))
    gb_, gp_, gc_ = torch.empty_like(tb), torch.empty_like(tp), torch.empty_like(tc)
    jf_ = torch.empty(n, 24, 3, device=dev)
    sp_ = torch.cuda.current_stream(dev).cuda_stream

    def bwd_call():
        capi.check(lib_.smplb200_backward(
            hb.ptr, tb.data_ptr(), tp.data_ptr(), tc.data_ptr(), n, jf_.data_ptr(), gv_.data_ptr(),
            gj_.data_ptr(), gk_.data_ptr(), gb_.data_ptr(), gp_.data_ptr(), gc_.data_ptr(),
            None, 0, wsb_t.data_ptr(), wsb, lay_t.flags, sp_), "smplb200_backward")

    for _ in range(3):
        bwd_call()
    dt_b = time_loop(bwd_call, 10, torch) / 10
    next_rows["backward"] = {
        "workload": f"trainer-shaped loss at {nt} bodies through the autograd node; smplb200_backward alone at {n} bodies",
        f"train_step_{nt}_bodies": bw,
        f"backward_call_{n}_bodies_us": dt_b * 1e6, "backward_bodies_per_s": n / dt_b,
        "launches_per_backward": int(lib_.smplb200_backward_launch_count(hb.ptr, n, lay_t.flags, 1, 0)),
        "cpu_kind": "port (torch autograd of oracle/smpl_ref.py, fp32, one call)"}
    del wsb_t, gv_
    # ---- §8(f) row 4: the reference's one native op, DCNv2 forward, on the DLA-34 layer shapes ----
    from human_3d_reconstruction_b200 import dcn_v2_conv
    from oracle.dcn_ref import dcn_v2_forward as dcn_cpu
    dcn_layers = [(1, 512, 256, 16), (1, 256, 256, 32), (2, 256, 128, 32), (2, 128, 128, 64),
                  (4, 128, 64, 64), (5, 64, 64, 128), (1, 256, 64, 32)]   # (count, Ci, Co, H=W)
    gd = torch.Generator().manual_seed(317)
    dcn_total, dcn_rows = 0.0, {}
    for cnt, Ci_, Co_, Hd in dcn_layers:
        xd = torch.randn(32, Ci_, Hd, Hd, generator=gd).to(dev)
        wd = (torch.randn(Co_, Ci_, 3, 3, generator=gd) / (Ci_ * 9) ** 0.5).to(dev)
        bd_ = torch.randn(Co_, generator=gd).to(dev)
        od = (torch.randn(32, 18, Hd, Hd, generator=gd) * 2.0).to(dev)
        md = torch.rand(32, 9, Hd, Hd, generator=gd).to(dev)
        for _ in range(2):
            dcn_v2_conv(xd, od, md, wd, bd_)
        dt_d = time_loop(lambda: dcn_v2_conv(xd, od, md, wd, bd_), 5, torch) / 5
        dcn_rows[f"{Ci_}->{Co_}@{Hd}x{Hd}"] = dt_d * 1e6
        dcn_total += cnt * dt_d
        del xd, wd, od, md
    xc = torch.randn(1, 64, 128, 128, generator=gd)
    t0 = time.perf_counter()
    dcn_cpu(xc, torch.randn(64, 64, 3, 3, generator=gd) / 24.0, torch.zeros(64),
            torch.randn(1, 18, 128, 128, generator=gd) * 2.0, torch.rand(1, 9, 128, 128, generator=gd))
    dt_dc = time.perf_counter() - t0
    next_rows["dcn_v2_forward"] = {
        "workload": "the 16 DeformConv layers of the reference DLA-34 (7 distinct shapes), batch 32, 512x512 input, "
                    "random offsets (sigma 2 px) and masks",
        "layer_us": dcn_rows, "network_16_layers_ms": dcn_total * 1e3,
        "cpu_reference_port_ms_per_image_one_64to64_128x128_layer": dt_dc * 1e3,
        "cpu_kind": "port (oracle/dcn_ref.py, pinned against torchvision CPU deform_conv2d and the reference KAT)"}
    return next_rows


if __name__ == "__main__":
    sys.exit(main())
