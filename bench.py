#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path (BASELINE.json): OCP-iterations/s, ANYmal trot N=40, batch 1024/GPU.

  python bench.py --gpus N --steps K --warmup W            our arm (one rank per GPU under torchrun for N > 1)
  python bench.py --impl reference --gpus N --steps K ...  the reference's CPU algorithm (oracle port, all host threads)

A "step" is one pass of the hot path over one batch of synthetic, HBM-resident KKT records:
backward Riccati sweep + forward Riccati sweep for every OCP of the batch (the parity-checked core of
OCPSolver::updateSolution, /root/reference/src/solver/ocp_solver.cpp:118-123).  Weak scaling: every GPU owns
`--batch` OCPs (instances are independent); with N > 1 the Newton step is all-gathered once per step over NCCL.
Prints ONE JSON line (rank 0).  PyTorch is plumbing only (streams, events, pinned memory, torch.distributed).
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

N_HORIZON = 40
BYTES_PER_STAGE_BWD = (3468 + 1776) * 8      # backward reads 3468 + writes 1776 doubles per standard stage
BYTES_PER_STAGE_FWD = (3324 + 84) * 8        # forward reads 3324 + writes 84 doubles
FLOP_PER_STAGE_BWD = 285.7e3
FLOP_PER_STAGE_FWD = 6.5e3


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=1024, help="OCP instances per GPU (BASELINE config 3: 1024)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def build_problem(batch, seed):
    from helpers import trot_schedule
    from robotoc_b200 import ANYMAL, Layout
    from robotoc_b200.synth import make_kkt
    dims = ANYMAL
    L = Layout(dims)
    td, ev, ctrl = trot_schedule(N_HORIZON)
    kkt, dx0 = make_kkt(dims, L, ctrl, batch=batch, seed=seed)
    return dims, L, ctrl, kkt, dx0


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """Samples SM clocks / throttle reasons with nvidia-smi while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.06)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def _oracle_runner(dims, L, ctrl, kkt, dx0):
    """Returns (lib, cores, run) where run() executes ONE timed pass of the oracle over the batch and returns seconds.
    The reference mutates its KKT in place, so the input is restored (untimed) before every pass; outputs are
    preallocated.  Only the C call is timed."""
    import ctypes as ct
    import oracle_lib
    lib = oracle_lib.load()
    native = os.path.join(oracle_lib.ORACLE_DIR, "liboracle_native.so")
    try:  # -march=native build of the same sources, made on this machine
        subprocess.run(["make", "-s", "-C", oracle_lib.ORACLE_DIR, "native"], check=True, stdout=subprocess.DEVNULL,
                       stderr=subprocess.DEVNULL)
        nl = ct.CDLL(native)
        nl.orc_riccati_batch.argtypes = lib.orc_riccati_batch.argtypes
        nl.orc_max_threads.restype = ct.c_int
        lib = nl
    except Exception:
        pass
    cores = int(lib.orc_max_threads())
    b, n_grid = kkt.shape[0], kkt.shape[1]
    kk = np.empty_like(kkt)
    ric = np.zeros((b, n_grid, L.r_stride))
    d = np.zeros((b, n_grid, L.d_stride))
    cd = dims.c()

    def run():
        np.copyto(kk, kkt)
        t0 = time.perf_counter()
        info = lib.orc_riccati_batch(ct.byref(cd), ctrl, n_grid, 0.1, b, oracle_lib.ptr(kk), oracle_lib.ptr(ric),
                                     oracle_lib.ptr(dx0), oracle_lib.ptr(d), 0)
        el = time.perf_counter() - t0
        assert info == 0
        return el

    return cores, run


def cpu_leg(dims, L, ctrl, kkt, dx0, min_seconds):
    """Times the oracle (CPU restatement of the reference algorithm) with all host threads; OCP-iterations/s."""
    cores, run = _oracle_runner(dims, L, ctrl, kkt, dx0)
    run()  # warm-up (page faults, thread pool)
    n, el = 0, 0.0
    while el < min_seconds:
        el += run()
        n += 1
    b = kkt.shape[0]
    return {"value": b * n / el, "unit": "OCP-iterations/s", "cores": cores, "kind": "port",
            "sample": f"{n} passes over {b} OCPs (riccati backward+forward; OpenMP over OCP instances, all host threads), "
                      f"{el:.1f} s of CPU-timed work"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    dims, L, ctrl, kkt, dx0 = build_problem(args.batch, 20260927)
    cores, run = _oracle_runner(dims, L, ctrl, kkt, dx0)
    for _ in range(max(args.warmup, 1)):
        run()
    el = 0.0
    for _ in range(args.steps):
        el += run()
    val = args.batch * args.steps / el
    line = {
        "impl": "reference", "metric": "SQP-iterations/s (ANYmal N=40, batch=1024) at 1/2/4/8 B200 vs CPU ref",
        "value": val, "unit": "OCP-iterations/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * el / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"anymal_trot_N40 batch={args.batch} riccati backward+forward", "n_grid": len(ctrl),
                   "note": "CPU restatement of the reference algorithm (oracle port; Eigen/Pinocchio absent so the reference "
                           "itself cannot be built), OpenMP over OCP instances on all host threads"},
        "cpu_baseline": {"value": val, "unit": "OCP-iterations/s", "cores": cores, "kind": "port",
                         "sample": f"{args.steps} passes over {args.batch} OCPs"},
        "e2e": {"value": val, "unit": "OCP-iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
        return
    import torch
    import torch.distributed as dist
    from robotoc_b200 import RiccatiRecursion
    from robotoc_b200.riccati import DIR, KKT

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: robotoc_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dims, L, ctrl, kkt, dx0 = build_problem(args.batch, 20260927 + rank)
    n_grid = len(ctrl)
    rr = RiccatiRecursion(dims, n_grid, args.batch, device=local)
    rr.setTimeDiscretization(ctrl)
    stream = torch.cuda.current_stream()
    sp = ctypes.c_void_p(stream.cuda_stream)

    # the Newton step lives in a torch tensor so NCCL can gather it in place
    d_local = torch.zeros((args.batch, n_grid, L.d_stride), dtype=torch.float64, device="cuda")
    rr.bind_buffer(DIR, ctypes.c_void_p(d_local.data_ptr()))
    d_all = torch.empty((world * args.batch, n_grid, L.d_stride), dtype=torch.float64, device="cuda") if world > 1 else None

    # ---- device-resident arm: upload once, time K steps
    rr.backwardRiccatiRecursion(kkt, stream=sp)   # uploads kkt
    rr.forwardRiccatiRecursion(dx0, stream=sp)    # uploads dx0
    torch.cuda.synchronize()

    def step(ev=None):
        if ev is not None:
            ev[0].record(stream)
        rr.backwardRiccatiRecursion(stream=sp)
        if ev is not None:
            ev[1].record(stream)
        rr.forwardRiccatiRecursion(stream=sp)
        if ev is not None:
            ev[2].record(stream)
        if world > 1:
            dist.all_gather_into_tensor(d_all.view(-1), d_local.view(-1))

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    l0 = rr.launch_count()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    t_beg, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    clk = ClockSampler(local)
    if rank == 0:
        clk.start()
    t_beg.record(stream)
    for k in range(args.steps):
        step(evs[k])
    t_end.record(stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    clocks = clk.stop() if rank == 0 else None
    launches = rr.launch_count() - l0
    ms = t_beg.elapsed_time(t_end)
    bwd_ms = float(np.mean([e[0].elapsed_time(e[1]) for e in evs]))
    fwd_ms = float(np.mean([e[1].elapsed_time(e[2]) for e in evs]))
    if world > 1:
        tt = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item())
    assert int(rr.info().max()) == 0, "Cholesky failure flagged on device"

    # ---- end-to-end arm: host (pinned) buffers through the one-call C-ABI entry point, H2D + D2H inside the timed region
    kkt_pin = torch.from_numpy(kkt).pin_memory()
    dx0_pin = torch.from_numpy(dx0).pin_memory()
    dir_pin = torch.empty((args.batch, n_grid, L.d_stride), dtype=torch.float64).pin_memory()
    lib = rr._lib

    def e2e_step():
        rc = lib.rbt_riccati_solve_host(rr._h, ctypes.c_void_p(kkt_pin.data_ptr()), ctypes.c_void_p(dx0_pin.data_ptr()),
                                        None, ctypes.c_void_p(dir_pin.data_ptr()), sp)
        assert rc == 0, rr._err()

    e2e_step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e_beg, e_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e_beg.record(stream)
    for _ in range(args.e2e_steps):
        e2e_step()
    e_end.record(stream)
    torch.cuda.synchronize()
    e2e_ms = e_beg.elapsed_time(e_end)
    if world > 1:
        tt = torch.tensor([e2e_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_ms = float(tt.item())
    h2d = int(lib.rbt_upload_bytes(rr._h, KKT)) + dx0.nbytes
    d2h = dir_pin.numel() * 8
    # the D2H result equals the device-resident result
    assert np.array_equal(dir_pin.numpy(), d_local.cpu().numpy()), "e2e path disagrees with the device-resident path"

    if rank == 0:
        peak, peak_src = peaks()
        units = args.batch * N_HORIZON                       # standard stages per launch
        ach = BYTES_PER_STAGE_BWD * units / (bwd_ms * 1e-3) / 1e9
        line = {
            "metric": "SQP-iterations/s (ANYmal N=40, batch=1024) at 1/2/4/8 B200 vs CPU ref",
            "value": world * args.batch * args.steps / (ms * 1e-3), "unit": "OCP-iterations/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"anymal_trot_N40 batch={args.batch}/GPU: riccati backward + forward"
                                   + (" + NCCL all-gather of the step" if world > 1 else ""),
                       "n_grid": n_grid, "dims": "nv18 nu12 nx36", "parallelism": f"batch-sharded x{world}",
                       "l2": f"per-step inputs {kkt.nbytes / 1e9:.2f} GB + outputs {rr.buf_doubles(1) * 8 / 1e9:.2f} GB >> 126 MB L2 "
                             "(no flush needed)"},
            "clocks": clocks, "gpu_launches": int(launches),
            "kernels_ms": {"riccati_backward": bwd_ms, "riccati_forward": fwd_ms},
            "e2e": {"value": world * args.batch * args.e2e_steps / (e2e_ms * 1e-3), "unit": "OCP-iterations/s",
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": args.e2e_steps,
                    "api": "rbt_riccati_solve_host (pinned host KKT in, Newton direction out)"},
            "roofline": {"bound": "hbm", "kernel": "riccati_backward_kernel<18,12,12>", "achieved": ach, "peak": peak,
                         "unit": "GB/s", "frac": ach / peak, "traffic": None, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": BYTES_PER_STAGE_BWD * units,
                         "fp64_tflops": FLOP_PER_STAGE_BWD * units / (bwd_ms * 1e-3) / 1e12,
                         "fp64_peak_tflops": 37.1,
                         "forward": {"achieved": BYTES_PER_STAGE_FWD * units / (fwd_ms * 1e-3) / 1e9, "unit": "GB/s"}},
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_leg(dims, L, ctrl, kkt, dx0, min_seconds=10.0)
        print(json.dumps(line), flush=True)
    rr.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
