#!/usr/bin/env python
"""bench.py — effective FP64 TFLOP/s of the contraction engine on BASELINE.json's headline
workload (configs[1]: dense DGEMM N=32768, tile=1024, FP64, SUMMA on a ProcGrid of N GPUs).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--n 32768] [--tile 1024]
  N > 1:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
              --master-port P bench.py --gpus N --steps K --warmup W

One "step" = one full contraction c("m,n") = a("m,k") * b("k,n") (2*N^3 flop), evaluated by the
product path: ProcGrid -> cyclic distribution -> tadev_summa_f64 (NCCL panel broadcasts overlapped
with grouped DMMA GEMM launches). Prints ONE JSON line (rank 0). `value` is device-timed (CUDA
events on the launching stream, max over ranks) with operands resident in HBM; `e2e` repeats the
measurement with HOST (pinned) operands and result, host<->device copies inside the timed region.
The oracle (oracle/) is used only for the cpu_baseline / --impl reference legs and a spot check.
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "effective FP64 TFLOP/s (device-timed, max over ranks)"


# ---------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        self.idx = gpu_index

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                       "-i", str(self.idx)], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        self.p.wait()
        self.f.flush()
        rows = [ln.strip().split(", ") for ln in open(self.f.name) if ln.strip()]
        os.unlink(self.f.name)
        sm, smax, power, reasons = [], 0, [], set()
        for r in rows:
            try:
                sm.append(float(r[1])); smax = max(smax, float(r[2])); power.append(float(r[3]))
            except (ValueError, IndexError):
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if len(r) > col and r[col].strip().lower() == "active":
                    reasons.add(name)
        load = [s for s, p in zip(sm, power) if p > 300] or sm
        return {"sm_mhz": statistics.median(load) if load else None, "sm_max_mhz": smax or None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------
def cpu_reference_leg(n, tile, budget_s=12.0):
    """The reference's CPU path restated (oracle/cpu_oracle.c): SUMMA on a 1x1 grid, one
    single-threaded vendor DGEMM per tile pair (tiledarray.cpp:112), one task thread per host core.
    Timed on a bounded S x S block of result tiles of the same workload (full K)."""
    from oracle import cpu as ocpu
    ocpu.load()
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    kt = n // tile
    # calibrate one pair on one core
    A = np.full((tile, tile), 0.5)
    B = np.full((tile, tile), 0.25)
    t0 = time.perf_counter()
    ocpu.gemm(0, 0, tile, tile, tile, 1.0, A, B, 0.0, np.empty((tile, tile)))
    t_pair = time.perf_counter() - t0
    pairs_budget = budget_s * cores * 0.7 / max(t_pair, 1e-6)
    S = int(max(1, min(kt, 12, np.floor(np.sqrt(pairs_budget / kt)))))
    rng = np.random.default_rng(0)
    base = rng.uniform(-1, 1, (tile, tile))
    a_tiles = {(i, k): base + (i * kt + k) for i in range(S) for k in range(kt)}  # distinct memory per tile
    b_tiles = {(k, j): base - (k * S + j) for k in range(kt) for j in range(S)}
    _, secs, npairs = ocpu.cpu_contract(a_tiles, b_tiles, [tile] * S, [tile] * S, [tile] * kt, 0, 0, 1.0, None, cores)
    flops = 2.0 * (S * tile) ** 2 * n
    return {"value": flops / secs / 1e12, "unit": "TFLOP/s", "cores": cores, "kind": "port",
            "sample": f"{S}x{S} result tiles of the N={n} tile={tile} contraction, full K ({npairs} tile pairs, "
                      f"{flops:.3e} flop, {secs:.2f} s); one single-threaded OpenBLAS DGEMM per pair on {cores} task threads"}, secs


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps_v = []
    base = None
    for _ in range(max(1, args.warmup if args.warmup < 2 else 1)):
        cpu_reference_leg(args.n, args.tile, budget_s=3.0)
    t_total = 0.0
    for _ in range(args.steps):
        base, secs = cpu_reference_leg(args.n, args.tile, budget_s=args.cpu_budget)
        steps_v.append(base["value"])
        t_total += secs
    v = float(np.mean(steps_v))
    base["value"] = v
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "TFLOP/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"dense DGEMM N={args.n} tile={args.tile} FP64 (BASELINE configs[1]); CPU leg timed on a bounded sample",
                   "parallelism": f"1 process x {base['cores']} task threads (MADWorld-style), BLAS pinned to 1 thread"},
        "cpu_baseline": base,
        "e2e": {"value": v, "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ---------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="tadev", choices=["tadev", "reference"])
    ap.add_argument("--n", type=int, default=32768)
    ap.add_argument("--tile", type=int, default=1024)
    ap.add_argument("--cpu-budget", type=float, default=12.0)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    rank = int(os.environ.get("RANK", "0"))
    size = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert size == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={size}: launch with torch.distributed.run for N>1"

    from tiledarray_b200 import Device, _lib
    from tiledarray_b200.tiledarray import ContEngine, DistArray, TiledRange, TiledRange1, World, summa_arrays

    dist = None
    if size > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = Device(local)
    world = World(device=dev, rank=rank, size=size)
    n, tile = args.n, args.tile
    t1 = TiledRange1.make_uniform(n, tile)
    tr = TiledRange([t1, t1])
    nt = t1.ntiles
    g = world.proc_grid(nt, nt, n, n)
    world.init_comm(g.proc_rows, g.proc_cols)
    a, b = summa_arrays(world, tr, tr)
    a.fill_random(3)
    b.fill_random(4)
    c = DistArray(world, tr)
    flops = 2.0 * float(n) ** 3

    def barrier():
        dev.sync()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def one_step():
        c["m,n"] = a["m,k"] * b["k,n"]
        return ContEngine.last_stats

    # FP64 roofline denominator measured live on this GPU (not in MEASURED_PEAKS.json)
    peak_tf = max(dev.probe_fp64_peak(0, 40000)[0] for _ in range(2))

    for _ in range(args.warmup):
        one_step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = dev.launch_count()
    kernel_ms, step_stats = [], None
    with dev.timer() as tm:  # CUDA events on the stream the driver launches on
        for _ in range(args.steps):
            step_stats = one_step()
            kernel_ms.append(step_stats.device_ms)
    launches = dev.launch_count() - launches0
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    t_ms = tm.ms
    if dist is not None:
        tt = torch.tensor([t_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_ms = tt.item()
    value = flops * args.steps / (t_ms * 1e-3) / 1e12

    # spot parity check of one result tile against the oracle (outside the timed region)
    parity = None
    if rank == 0:
        from oracle import ta_oracle as O
        from tests import util_rng
        o = sorted(c.tiles)[0]
        i, j = tr.tile_index(o)
        ref = np.zeros((tile, tile))
        for k in range(nt):
            ref += util_rng.tile_fill(i * nt + k, tile * tile, 3).reshape(tile, tile) @ util_rng.tile_fill(k * nt + j, tile * tile, 4).reshape(tile, tile)
        parity = O.rel_frobenius(c.find(o), ref)

    # ---- e2e: the same contraction through the public API with HOST-resident (pinned) operands and
    # result: c_h["m,n"] = a_h["m,k"] * b_h["k,n"]. Every step moves all operand bytes host->device
    # and the whole result device->host inside the timed region (the driver streams operand panels
    # window by window and returns the result in row blocks so the copies overlap the GEMMs).
    e2e = None
    if not args.no_e2e:
        lib = dev.lib
        a_h, b_h = summa_arrays(world, tr, tr, memory="host")
        a_h._allocate()
        b_h._allocate()
        assert a_h._arena.nbytes == a._arena.nbytes and b_h._arena.nbytes == b._arena.nbytes
        _lib.check(lib.tadev_memcpy_d2h(dev.ctx, a_h._arena.ptr, a._arena.ptr, a._arena.nbytes, dev.stream))
        _lib.check(lib.tadev_memcpy_d2h(dev.ctx, b_h._arena.ptr, b._arena.ptr, b._arena.nbytes, dev.stream))
        dev.sync()
        c_ref_tile = c.find(sorted(c.tiles)[0]) if rank == 0 else None
        for x in (a, b, c):
            x.release()  # the device-resident copies are not used by the e2e leg
        c_h = DistArray(world, tr, memory="host")
        h2d = d2h = 0

        def e2e_step():
            c_h["m,n"] = a_h["m,k"] * b_h["k,n"]
            return ContEngine.last_stats

        e2e_step()
        barrier()
        with dev.timer() as te:
            for _ in range(args.steps):
                st_e = e2e_step()
                h2d, d2h = st_e.h2d_bytes, st_e.d2h_bytes
        barrier()
        te_ms = te.ms
        if rank == 0:  # the streamed path must reproduce the device-resident result (window partial sums
            #            are added in a different association, so equal to rounding, not bit for bit)
            t_e = c_h.find(sorted(c_h.tiles)[0])
            assert np.linalg.norm(t_e - c_ref_tile) <= 1e-13 * np.linalg.norm(c_ref_tile), "e2e result differs from the device-resident run"
        if dist is not None:
            tt = torch.tensor([te_ms, float(h2d), float(d2h)], dtype=torch.float64, device="cuda")
            mx = tt.clone()
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            dist.all_reduce(tt, op=dist.ReduceOp.SUM)
            te_ms, h2d, d2h = mx[0].item(), int(tt[1].item()), int(tt[2].item())
        e2e = {"value": flops * args.steps / (te_ms * 1e-3) / 1e12, "unit": "TFLOP/s", "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "ms_per_step": te_ms / args.steps,
               "api": "DistArray(memory='host') operands + result; panels streamed, result returned in row blocks"}
        for x in (a_h, b_h, c_h):
            x.release()

    cpu = None
    if rank == 0 and size == 1 and not args.no_cpu:
        cpu, _ = cpu_reference_leg(n, tile, args.cpu_budget)

    if rank == 0:
        # dominant kernel: the grouped DMMA GEMM. algorithmic flop per launch = this rank's share
        # of 2*N^3 / launches per step; duration = CUDA-event time of the driver's launches
        nl = max(1, step_stats.nlaunches)
        k_ms = float(np.mean(kernel_ms)) / nl
        achieved = step_stats.flops / nl / (k_ms * 1e-3) / 1e12
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "gemm_traffic.json")
        if os.path.exists(tpath) and size == 1:
            traffic = json.load(open(tpath)).get(f"n{n}_t{tile}")
        out = {
            "metric": METRIC, "value": value, "unit": "TFLOP/s", "n_gpus": size, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": t_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"dense DGEMM N={n} tile={tile} FP64, c(m,n)=a(m,k)*b(k,n) (BASELINE configs[1])",
                       "flop_per_step": flops, "parallelism": f"SUMMA {g.proc_rows}x{g.proc_cols} process grid, 1 process/GPU",
                       "l2": f"operands {2 * 8 * n * n / 1e9:.1f} GB per step >> 126 MB L2 (no flush needed)",
                       "spot_parity_rel_frobenius": parity},
            "clocks": clocks, "gpu_launches": int(launches), "e2e": e2e,
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf,
                         "traffic": traffic, "kernel": "gemm_grouped_f64_ws_kernel<N,N>",
                         "peak_source": "DMMA.8x8x4 register-resident issue-rate probe measured in this run "
                                        "(tadev_probe_fp64_peak; MEASURED_PEAKS.json has no FP64 entry)",
                         "launches_per_step": int(nl), "ms_per_launch": k_ms},
            "cpu_baseline": cpu,
        }
        print(json.dumps(out))
    if dist is not None:
        dist.barrier()
    dev.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
