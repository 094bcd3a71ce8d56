#!/usr/bin/env python3
"""
bench.py - depth frames/s of the DPT hot path (DPTModel.forward) on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference|reference-gpu] [--model vitl|vitb|vits]
                  [--batch B] [--size S] [--dtype bf16|fp16] [--scaling strong|weak]

Workload (BASELINE.json metric "depth frames/sec at 518^2 ViT-L bf16, 1/2/4/8xB200", configs[2]): Depth-Anything-V2
ViT-L, global batch 32, 504x504 (the reference's effective "518" setting - SURVEY.md section 0.1: 518 has an odd patch grid
and the reference's forward raises), bf16, synthetic seeded weights and inputs. One "step" = one forward of the
global batch. N>1: the batch is sharded over ranks (strong scaling, per-rank batch 32/N), every rank runs the same
kernels on its frames and the depth maps are all-gathered once per step over NCCL.

Printed JSON line (rank 0): value = frames/s with inputs resident in HBM (CUDA events, max over ranks);
e2e = same metric through the host-buffer C-ABI call (dpt_forward_host: pinned host -> H2D -> forward -> D2H);
roofline = the dominant kernel (tcgen05 GEMM) from per-launch CUDA events recorded inside the library during extra
profiled steps; cpu_baseline = the reference's fp32 CPU path on this box's host cores.

Reference arms (no product code on their path): `--impl reference` = the UNMODIFIED reference package from oracle/_ref
(placed there by oracle/build_ref.py; the oracle port stands in only when that copy is absent) through its own
make_dpt_from_state_dict() / DPTModel.forward() on the host CPU in fp32, W warm-up and K timed steps like the native
arm, each step a bounded sample (one frame) of the batch; `--impl reference-gpu` = the same reference model moved to
the B200 with stock PyTorch eager (`model.to("cuda", bf16, channels_last)`, SURVEY.md section 8d "the real bar").
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "depth_frames_per_sec"
UNIT = "frames/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference", "reference-gpu"])
    ap.add_argument("--ref-frames-per-step", type=int, default=1, help="reference arm: frames per timed step (bounded sample)")
    ap.add_argument("--model", default="vitl",
                    choices=["vitl", "vitb", "vits", "vitg", "tiny", "beit_large_384", "beit_base_384", "beit_tiny",
                             "swinv2_large_384", "swinv2_base_384", "swinv2_tiny_256", "swinv2_micro"])
    ap.add_argument("--batch", type=int, default=32, help="global batch (frames per step)")
    ap.add_argument("--size", type=int, default=504)
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp16"])
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"])
    ap.add_argument("--collective", default="c-abi", choices=["c-abi", "torch"],
                    help="N > 1: the depth all-gather through dpt_allgather_depth (own ncclComm_t) or torch.distributed")
    ap.add_argument("--cpu-baseline-frames", type=int, default=16, help="bounded CPU sample: frames of B=1 (about 10 s)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--profile-steps", type=int, default=2)
    ap.add_argument("--dump-profile", default="", help="write the per-launch table to this path")
    return ap.parse_args()


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            pk = json.load(f)
        return {"tflops_sustained": pk.get("bf16_tflops_sustained", 1413.6), "tflops_burst": pk.get("bf16_tflops", 1685.6),
                "hbm_gbs": pk.get("hbm_gbs", 6538.0), "source": "measured (MEASURED_PEAKS.json)"}
    return {"tflops_sustained": 1400.0, "tflops_burst": 1590.0, "hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}


def algorithmic_gflop_per_frame(model, size):
    """BASELINE.md section 3 (torch flop counter on the reference modules, 2*MAC, matmul/conv only)."""
    table = {("vits", 504): 107.32, ("vitb", 504): 356.69, ("vitl", 504): 1224.94, ("vitl", 532): 1385.8,
             ("beit_large_384", 384): 516.41, ("swinv2_large_384", 384): 343.73}
    return table.get((model, size))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)"""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.gpu_index)],
                stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        try:
            self.proc.terminate()
            self.proc.wait(timeout=5)
        except Exception:
            pass
        try:
            sm, mx, reasons = [], [], set()
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1]))
                    mx.append(float(f[2]))
                except ValueError:
                    continue
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            if sm:
                out = {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}
            os.unlink(self.path)
        except Exception:
            pass
        return out


def _oracle_model(O, model_name):
    """(synthetic upstream-format checkpoint, oracle forward) for a model name"""
    if model_name.startswith("beit"):
        return O.make_synthetic_state_dict_beit(model_name, seed=11), O.forward_beit
    if model_name.startswith("swinv2"):
        return O.make_synthetic_state_dict_swinv2(model_name, seed=11), O.forward_swinv2
    sd = O.make_synthetic_state_dict(model_name, seed=11)
    return (O.giantify(sd, seed=11) if model_name == "vitg" else sd), O.forward


def _checkpoint_file_name(model_name):
    # the reference sniffs Depth-Anything v1/v2 from the FILE NAME (make_dpt.py:98-104)
    if model_name.startswith(("beit", "swinv2")):
        return f"dpt_{model_name}_synthetic.pt"
    return f"depth_anything_v2_{model_name}_synthetic.pth"


def host_threads():
    """every host core this process may run on (torchrun sets OMP_NUM_THREADS=1 per rank: undo that for the CPU legs)"""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def workload_text(model_name, batch, size, dtype):
    fam = "MiDaS v3.1" if model_name.startswith(("beit", "swinv2")) else "Depth-Anything-V2"
    return (f"{fam} {model_name} (synthetic seeded weights), global batch {batch}, 3x{size}x{size}"
            + (" (reference's effective 518 setting)" if size == 504 else "") + f", {dtype}")


def workload_config(model_name, b_global, b_local, world, size, dtype):
    """the `config` object of the JSON line - identical for the native and the reference arms (it names the workload)"""
    return {"workload": workload_text(model_name, b_global, size, dtype), "global_batch": b_global,
            "per_gpu_batch": b_local, "parallelism": f"dp{world} batch-shard + all-gather",
            "l2": "256 MiB buffer rewritten between iterations (L2 flush); per-step working set >> 126 MB L2"}


def load_reference_model(model_name):
    """(model, kind): the unmodified reference from oracle/_ref through its own factory, else the oracle port"""
    import torch

    from oracle import dpt_oracle as O
    from oracle.build_ref import import_reference, ref_available

    sd, fwd = _oracle_model(O, model_name)
    if not ref_available():
        return (lambda x: fwd(sd, x)), "port", None
    make_dpt = import_reference()
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, _checkpoint_file_name(model_name))
        torch.save(sd, path)
        del sd
        _, model = make_dpt.make_dpt_from_state_dict(path)
    return model, "reference", model


def time_cpu_reference(model_name, size, frames, threads=None):
    """the reference's fp32 CPU path, B=1 per step: returns (frames/s, cores, sample text, kind)"""
    import torch

    from oracle import dpt_oracle as O

    torch.set_num_threads(threads or host_threads())
    cores = torch.get_num_threads()
    fwd, kind, _ = load_reference_model(model_name)
    img = O.make_input(1, size, size, seed=2)
    with torch.inference_mode():
        fwd(img)  # warm-up
        ts = []
        for _ in range(frames):
            t0 = time.perf_counter()
            fwd(img)
            ts.append(time.perf_counter() - t0)
    med = statistics.median(ts)
    what = "unmodified reference (oracle/_ref) DPTModel.forward" if kind == "reference" else "oracle port of the reference path"
    sample = (f"{what}: {frames} frames of B=1 {model_name} {size}x{size} fp32 after 1 warm-up, median; "
              f"torch threads={cores} of {os.cpu_count()} cpus")
    return 1.0 / med, cores, sample, kind


def run_reference(args, rank, world):
    """reference arm on the host CPU: the reference's own implementation of the path (oracle/_ref), all host threads,
    the same warm-up / step counts as the native arm, each step a bounded sample of the workload."""
    if rank != 0:
        return
    import torch

    from oracle import dpt_oracle as O

    torch.set_num_threads(host_threads())
    cores = torch.get_num_threads()
    fwd, kind, _ = load_reference_model(args.model)
    n = max(1, args.ref_frames_per_step)
    img = O.make_input(n, args.size, args.size, seed=2)
    warmup, steps = max(args.warmup, 0), max(args.steps, 1)
    with torch.inference_mode():
        for _ in range(warmup):
            fwd(img)
        t0 = time.perf_counter()
        for _ in range(steps):
            fwd(img)
        dt = time.perf_counter() - t0
    fps = n * steps / dt
    b_local = args.batch // world if args.scaling == "strong" else args.batch
    b_global = args.batch if args.scaling == "strong" else args.batch * world
    what = "unmodified reference (oracle/_ref) make_dpt_from_state_dict -> DPTModel.forward" if kind == "reference" \
        else "oracle port of the reference path (oracle/_ref absent)"
    sample = (f"{what}; each step = {n} frame(s), a bounded sample of the batch-{b_global} workload, {args.model} "
              f"{args.size}x{args.size}, fp32 on the host CPU, {warmup} warm-up + {steps} timed steps; "
              f"torch threads={cores} of {os.cpu_count()} cpus")
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": 1000.0 * dt / steps, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.model, b_global, b_local, world, args.size, args.dtype),
        "reference_run": {"device": "host cpu", "compute_dtype": "f32", "frames_per_step": n, "kind": kind},
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


def run_reference_gpu(args, rank, world):
    """the on-box bar: the unmodified reference on the same B200 through stock PyTorch eager,
    model.to("cuda", bf16 / fp16, channels_last) as run_image.py:158 does, same batch / size / dtype as the native arm."""
    if rank != 0:
        return
    import torch

    from oracle import dpt_oracle as O

    assert torch.cuda.is_available()
    fwd, kind, model = load_reference_model(args.model)
    if model is None:
        emit({"impl": "reference-gpu", "unavailable": "oracle/_ref is absent (run oracle/build_ref.py where /root/reference exists)"})
        return
    dtype = torch.bfloat16 if args.dtype == "bf16" else torch.float16
    dev = torch.device("cuda", 0)
    model.to(device=dev, dtype=dtype, memory_format=torch.channels_last)
    B, S = args.batch, args.size
    g = torch.Generator().manual_seed(1234)
    img = torch.randn(B, 3, S, S, generator=g).to(dtype).to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    sampler = ClockSampler(0)
    with torch.inference_mode():
        for _ in range(max(args.warmup, 3)):
            flush.zero_()
            model(img)
        torch.cuda.synchronize()
        sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            flush.zero_()
            out = model(img)
        e1.record()
        torch.cuda.synchronize()
        clocks = sampler.stop()
    ms = e0.elapsed_time(e1)
    line = {
        "impl": "reference-gpu", "metric": METRIC, "value": B * args.steps / (ms / 1000.0), "unit": UNIT, "n_gpus": 1,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": workload_config(args.model, B, B, 1, S, args.dtype), "clocks": clocks,
        "reference_run": {"device": "cuda:0 (stock PyTorch eager: cuBLAS / cuDNN / fused SDPA)", "compute_dtype": args.dtype,
                          "memory_format": "channels_last", "kind": kind, "output_shape": list(out.shape)},
    }
    emit(line)


_REAL_STDOUT = None


def quiet_stdout():
    """Libraries (NCCL's version banner, for one) print to fd 1: point it at stderr for the run and keep the original for
    the ONE JSON line."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line: dict):
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    args = parse_args()
    quiet_stdout()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.impl == "reference-gpu":
        run_reference_gpu(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    from muggled_dpt_b200 import make_dpt_from_state_dict
    from muggled_dpt_b200.distributed import all_gather_depth, shard_range
    from oracle import dpt_oracle as O  # synthetic checkpoint generator + cpu_baseline only

    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # keep whatever NCCL_DEBUG the caller / driver exported (its rank + topology lines verify the multi-GPU run);
        # NCCL writes its log to stdout by default: send it to stderr so stdout stays the one JSON line
        os.environ.setdefault("NCCL_DEBUG", os.environ.get("DPT_NCCL_DEBUG", "WARN"))
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)

    dtype = torch.bfloat16 if args.dtype == "bf16" else torch.float16
    if args.scaling == "strong":
        assert args.batch % world == 0, "global batch must divide over ranks"
        lo, hi = shard_range(args.batch, rank, world)
        b_local = hi - lo
        b_global = args.batch
    else:
        b_local = args.batch
        b_global = args.batch * world
    S = args.size

    # ---- model (synthetic seeded checkpoint in the upstream format, loaded through the reference-shaped factory)
    sd, _ = _oracle_model(O, args.model)
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, _checkpoint_file_name(args.model))
        torch.save(sd, path)
        del sd
        cfg, model = make_dpt_from_state_dict(path)
    model.to(device=dev, dtype=dtype, memory_format=torch.channels_last)

    g = torch.Generator().manual_seed(1234 + rank)
    host_img = torch.randn(b_local, 3, S, S, generator=g).to(dtype).pin_memory()
    host_out = torch.empty(b_local, S, S, dtype=dtype).pin_memory()
    img = host_img.to(dev)
    out = torch.empty(b_local, S, S, dtype=dtype, device=dev)
    gathered = torch.empty(b_global, S, S, dtype=dtype, device=dev) if world > 1 else None
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    # the path's one collective: through the library's own C entry point (a communicator created on the NCCL this
    # process carries), torch.distributed as the fallback
    native_gather, collective = None, "none"
    if world > 1:
        collective = "torch.distributed all_gather_into_tensor"
        if args.collective == "c-abi" and args.scaling == "strong":
            try:
                from muggled_dpt_b200.distributed import NativeDepthAllGather

                native_gather = NativeDepthAllGather()
                collective = "dpt_allgather_depth (C ABI, own ncclComm_t)"
            except Exception as e:  # noqa: BLE001 - any failure here must not cost the run
                native_gather = None
                collective += f" (dpt_allgather_depth unavailable: {type(e).__name__}: {e})"
        ok = torch.tensor([1 if native_gather is not None else 0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0 and native_gather is not None:  # all ranks or none
            native_gather.close()
            native_gather = None
            collective = "torch.distributed all_gather_into_tensor"

    def gather(local):
        if native_gather is not None:
            native_gather(local, gathered)
        else:
            all_gather_depth(local, b_global, out=gathered)

    def step():
        flush.zero_()  # L2 flush between iterations (B200_PROFILING.md timing hygiene)
        model.forward_into(img, out)
        if world > 1:
            gather(out)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        barrier()
        return ms

    with torch.inference_mode():
        for _ in range(max(args.warmup, 3)):
            step()
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        ms_total = timed(step, args.steps)
        clocks = sampler.stop() if rank == 0 else None
        launches_per_step = model.last_launch_count() + 1 + (1 if world > 1 else 0)
        ms_per_step = ms_total / args.steps
        fps = b_global * args.steps / (ms_total / 1000.0)

        # ---- e2e: host buffers through the C ABI. Every step copies its input from pinned host memory and its result
        #      back, inside the timed region; the calls are double-buffered (dpt_forward_host_async on two device buffer
        #      pairs), so the H2D of step i+1 and the D2H of step i-1 overlap the forward of step i.
        e2e = None
        if not args.no_e2e:
            host_outs = [host_out, torch.empty_like(host_out).pin_memory()]
            pending = [None, None]
            counter = [0]

            def step_host():
                k = counter[0] & 1
                counter[0] += 1
                if pending[k] is not None:
                    pending[k].synchronize()  # this slot's previous result is on the host before its buffers are reused
                pending[k] = model.forward_host_async(host_img, host_outs[k], slot=k)
                if world > 1:
                    gather(model._io_buffers(b_local, S, S, k)[1])

            def drain():
                for ev in pending:
                    if ev is not None:
                        ev.synchronize()

            for _ in range(4):
                step_host()
            drain()
            ms_e2e = timed(lambda: step_host(), args.steps)  # (timed() ends with a device-wide synchronize)
            drain()
            e2e = {"value": b_global * args.steps / (ms_e2e / 1000.0), "unit": UNIT,
                   "h2d_bytes_per_step": host_img.numel() * host_img.element_size() * world,
                   "d2h_bytes_per_step": host_out.numel() * host_out.element_size() * world,
                   "ms_per_step": ms_e2e / args.steps,
                   "path": "DPTModel.forward_host_async -> dpt_forward_host_async (pinned host buffers, double-buffered: the "
                           "copies of step i+1 / i-1 overlap the forward of step i)"}

        # ---- roofline leg: per-launch CUDA events inside the library on extra steps
        roofline, breakdown = None, None
        if rank == 0 and args.profile_steps > 0:
            model.enable_profiling(True)
            agg = {}
            for _ in range(args.profile_steps):
                flush.zero_()
                model.forward_into(img, out)
                torch.cuda.synchronize()
                for label, ms, fl, by in model.read_profile():
                    fam = label.split(":")[0]
                    a = agg.setdefault(fam, [0.0, 0.0, 0.0, 0])
                    a[0] += ms; a[1] += fl; a[2] += by; a[3] += 1
                last_profile = model.read_profile()
            model.enable_profiling(False)
            tot_ms = sum(a[0] for a in agg.values())
            breakdown = {fam: {"ms_per_step": a[0] / args.profile_steps, "share": a[0] / tot_ms,
                               "tflops": (a[1] / (a[0] / 1000.0) / 1e12) if a[0] > 0 else 0.0,
                               "gbs": (a[2] / (a[0] / 1000.0) / 1e9) if a[0] > 0 else 0.0,
                               "launches_per_step": a[3] // args.profile_steps} for fam, a in agg.items()}
            peaks = load_peaks()
            dom = max(agg.items(), key=lambda kv: kv[1][0])[0]
            a = agg[dom]
            traffic, traffic_src = None, None
            try:
                with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                    tj = json.load(f).get(dom)
                if tj and args.model == "vitl" and b_local == 32 and S == 504:
                    traffic, traffic_src = tj["traffic_bytes_per_launch"], tj["source"]
            except Exception:
                pass
            if a[1] > 0:
                ach = a[1] / (a[0] / 1000.0) / 1e12
                roofline = {"kernel": dom, "bound": "tensor", "achieved": ach, "peak": peaks["tflops_sustained"],
                            "unit": "TFLOP/s", "frac": ach / peaks["tflops_sustained"], "traffic": traffic,
                            "traffic_source": traffic_src, "algorithmic_bytes_per_launch": a[2] / a[3],
                            "peak_source": peaks["source"] + " sustained bf16 (kernel timed inside a long step)",
                            "avg_launch_ms": a[0] / a[3], "flop_per_launch": a[1] / a[3],
                            "share_of_step": a[0] / tot_ms}
            else:
                ach = a[2] / (a[0] / 1000.0) / 1e9
                roofline = {"kernel": dom, "bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                            "frac": ach / peaks["hbm_gbs"], "traffic": None, "peak_source": peaks["source"],
                            "avg_launch_ms": a[0] / a[3], "share_of_step": a[0] / tot_ms}
            if args.dump_profile:
                with open(args.dump_profile, "w") as f:
                    f.write("label,ms,gflop,mbytes,tflops,gbs\n")
                    for label, ms, fl, by in last_profile:
                        f.write(f"{label},{ms:.4f},{fl / 1e9:.3f},{by / 1e6:.3f},"
                                f"{(fl / (ms / 1e3) / 1e12) if ms > 0 else 0:.1f},{(by / (ms / 1e3) / 1e9) if ms > 0 else 0:.0f}\n")

    # ---- CPU baseline (rank 0, N = 1 only): the oracle port on this box's host cores
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, cores, sample, kind = time_cpu_reference(args.model, S, args.cpu_baseline_frames)
        cpu_baseline = {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample}

    if rank == 0:
        gf = algorithmic_gflop_per_frame(args.model, S)
        peaks = load_peaks()
        line = {
            "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": workload_config(args.model, b_global, b_local, world, S, args.dtype),
            "collective": collective,
            "value_path": "DPTModel.forward_into on caller-owned device buffers (model(x) adds one input copy_ and one "
                          "output clone, ~65 MB of device copies at B=32)",
            "clocks": clocks,
            "e2e": e2e,
            "gpu_launches": launches_per_step * args.steps,
            "gpu_launches_per_step": launches_per_step,
            "roofline": roofline,
            "cpu_baseline": cpu_baseline,
            "kernel_breakdown": breakdown,
        }
        if gf:
            tf = gf * 1e9 * fps / 1e12
            line["model_tflops"] = {"algorithmic_gflop_per_frame": gf, "achieved_tflops_whole_job": tf,
                                    "frac_of_sustained_peak_per_gpu": tf / world / peaks["tflops_sustained"]}
        emit(line)

    if world > 1:
        if native_gather is not None:
            native_gather.close()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
