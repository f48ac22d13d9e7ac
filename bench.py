#!/usr/bin/env python
"""bench.py — effective FP64 TFLOP/s of the contraction engine on BASELINE.json's headline
workload (configs[1]: dense DGEMM N=32768, tile=1024, FP64, SUMMA on a ProcGrid of N GPUs).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config C1|C2|C3|C3m|C4|C5] [--n ..] [--tile ..]
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
def measure_gemm_traffic(args, max_launches=20, timeout_s=180):
    """DRAM bytes per launch of the GEMM kernel, measured IN THIS RUN: a child `ncu` profiles the same command line
    (one step, no warm-up, no e2e / CPU legs) for dram__bytes_read.sum + dram__bytes_write.sum of up to
    `max_launches` launches of gemm_grouped_f64_ws_kernel and the mean per launch is reported (counters, not timings:
    nothing timed under the profiler is used). Returns (bytes_per_launch, note) or (None, why)."""
    import csv
    import io
    import shutil
    if any(k.startswith(("NV_NSIGHT", "CUDA_INJECTION", "NV_COMPUTE_PROFILER", "NSYS_")) for k in os.environ):
        return None, "this process is itself running under a profiler: no nested ncu"
    ncu = shutil.which("ncu") or "/usr/local/cuda/bin/ncu"
    if not os.path.exists(ncu):
        return None, "ncu not found"
    cmd = [ncu, "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum", "--clock-control", "none", "--print-units", "base",
           "-k", "regex:gemm_grouped_f64_ws_kernel", "-c", str(max_launches), "--csv", sys.executable, os.path.abspath(__file__),
           "--config", args.config, "--steps", "1", "--warmup", "0", "--no-e2e", "--no-cpu", "--no-traffic", "--parity-tiles", "0"]
    if args.n:
        cmd += ["--n", str(args.n)]
    if args.tile:
        cmd += ["--tile", str(args.tile)]
    try:
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout_s)
    except (subprocess.TimeoutExpired, OSError) as ex:
        return None, f"ncu child failed: {ex}"
    rows = [ln for ln in out.stdout.splitlines() if ln.startswith('"')]
    if len(rows) < 2:
        return None, "ncu produced no counters (profiling not permitted on this box?): " + (out.stderr or out.stdout)[-200:].replace("\n", " ")
    per_launch = {}
    rd = csv.DictReader(io.StringIO("\n".join(rows)))
    for r in rd:
        try:
            per_launch[r["ID"]] = per_launch.get(r["ID"], 0.0) + float(r["Metric Value"].replace(",", ""))
        except (KeyError, ValueError):
            continue
    if not per_launch:
        return None, "could not parse the ncu output"
    vals = list(per_launch.values())
    return float(np.mean(vals)), f"mean of {len(vals)} launch(es) profiled by a child ncu in this run (dram__bytes_read.sum + dram__bytes_write.sum)"


# ---------------------------------------------------------------------------------------------
def _cpu_plan(config, n, tile):
    """(label, m_ext, n_ext, k_ext, opA, opB, a_nz, b_nz, c_nz) of the tile-level contraction the reference's CPU path
    executes for `config`: fused tile extents, BLAS op flags as the reference's permutation optimizer picks them (no
    operand exchange: permopt.h:254-376), and the non-zero patterns for the block-sparse config."""
    from bench_workloads import example_tiling
    from tiledarray_b200.tiledarray import TiledRange1
    if config in ("C1", "C2"):
        nn, tt = (n or (4096 if config == "C1" else 32768)), (tile or (256 if config == "C1" else 1024))
        e = TiledRange1.make_uniform(nn, tt).extents
        return f"dense N={nn} tile={tt}", e, e, e, 0, 0, None, None, None
    if config in ("C3", "C3m"):
        nn, tt = n or 65536, tile or 512
        nt = nn // tt

        def pat(seed):
            rng = np.random.default_rng(seed)
            p = np.zeros(nt * nt, dtype=bool)
            p[rng.permutation(nt * nt)[: int(0.10 * nt * nt)]] = True
            return p.reshape(nt, nt)
        A, B = pat(5), pat(6)
        Cz = (A.astype(np.int64) @ B.astype(np.int64)) > 0
        if config == "C3m":
            Cz &= A
        e = [tt] * nt
        return f"block-sparse N={nn} tile={tt} 10%", e, e, e, 0, 0, A, B, Cz
    if config in ("C4", "C4h", "C4r"):
        v = {"C4": 800, "C4h": 400, "C4r": 256}[config]
        o1, v1 = example_tiling(100, 64).extents, example_tiling(v, 64).extents
        oo = [x * y for x in o1 for y in o1]
        vv = [x * y for x in v1 for y in v1]
        # as written: left T(c,d,i,j) -> opA = T (stored [cd][ij]), right V(a,b,c,d) -> opB = T (stored [ab][cd]); the
        # result permutation (i,j,a,b)->(a,b,i,j) of every tile is not part of the timed sample (GEMM only)
        return f"CCSD PPL o=100 v={v} tile=64 (GEMM part, ops T,T as the reference plans it)", oo, vv, vv, 1, 1, None, None, None
    if config in ("C5", "C5r"):
        small = 128 if config == "C5" else 64
        s1, b1 = TiledRange1.make_uniform(small, 16).extents, TiledRange1.make_uniform(512, 64).extents
        sb = [x * y for x in s1 for y in b1]
        return (f"permuted 4-index i=j=k={small} a=b=c=512 (GEMM part after the explicit tile permutations)", sb, sb,
                [x * y for x in b1 for y in s1], 0, 0, None, None, None)
    raise ValueError(config)


def cpu_reference_leg(config, n=None, tile=None, budget_s=12.0):
    """The reference's CPU path restated (oracle/cpu_oracle.c): SUMMA on a 1x1 grid, one single-threaded vendor DGEMM
    per tile pair (tiledarray.cpp:112), one task thread per host core. Timed on a bounded S x S block of result tiles
    of the same workload at full K; the rate of that sample is reported (an extrapolation, not a full run)."""
    from oracle import cpu as ocpu
    ocpu.load()
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    label, m_ext, n_ext, k_ext, opA, opB, a_nz, b_nz, c_nz = _cpu_plan(config, n, tile)
    kt = len(k_ext)
    mt, ntl, kk = int(np.median(m_ext)), int(np.median(n_ext)), int(np.median(k_ext))
    # calibrate one typical pair on one core
    A = np.full((mt, kk) if opA == 0 else (kk, mt), 0.5)
    B = np.full((kk, ntl) if opB == 0 else (ntl, kk), 0.25)
    t0 = time.perf_counter()
    ocpu.gemm(opA, opB, mt, ntl, kk, 1.0, A, B, 0.0, np.empty((mt, ntl)))
    t_pair = max(time.perf_counter() - t0, 1e-6)
    dens = 1.0 if a_nz is None else float(a_nz.mean())
    pairs_budget = budget_s * cores * 0.7 / t_pair
    cap = 12 if a_nz is None else 48
    S = int(max(1, min(len(m_ext), len(n_ext), cap, np.floor(np.sqrt(pairs_budget / max(kt * dens * dens, 1e-9))))))
    if S * S < cores:  # few, huge tile pairs (config 4): keep every task thread busy and shorten the K chain instead
        S = int(min(len(m_ext), len(n_ext), np.ceil(np.sqrt(cores))))
    k_ext = list(k_ext)
    k_note = "full K"

    def est_s(kn):
        return S * S * kn * dens * dens * t_pair / min(S * S, cores) / 0.7

    def mem_b(kl):
        return 8.0 * S * sum(kl) * (mt + ntl) * max(dens, 0.02)

    while len(k_ext) > 1 and (mem_b(k_ext) > 6e9 or est_s(len(k_ext)) > 1.5 * budget_s):
        k_ext.pop()
        k_note = f"first {len(k_ext)} of {kt} K tiles"
    while S > 1 and mem_b(k_ext) > 6e9:
        S -= 1
    kt = len(k_ext)
    rng = np.random.default_rng(0)
    me, ne = list(m_ext[:S]), list(n_ext[:S])
    a_tiles, b_tiles = {}, {}
    for i in range(S):
        for k in range(kt):
            if a_nz is None or a_nz[i, k]:  # noqa: E501
                a_tiles[(i, k)] = rng.uniform(-1, 1, (me[i], k_ext[k]) if opA == 0 else (k_ext[k], me[i]))
    for k in range(kt):
        for j in range(S):
            if b_nz is None or b_nz[k, j]:
                b_tiles[(k, j)] = rng.uniform(-1, 1, (k_ext[k], ne[j]) if opB == 0 else (ne[j], k_ext[k]))
    c_zero = None if c_nz is None else ~c_nz[:S, :S]
    _, secs, npairs = ocpu.cpu_contract(a_tiles, b_tiles, me, ne, list(k_ext), opA, opB, 1.0, c_zero, cores)
    flops = 0.0
    for (i, k) in a_tiles:
        for j in range(S):
            if (k, j) in b_tiles and (c_zero is None or not c_zero[i, j]):
                flops += 2.0 * me[i] * ne[j] * k_ext[k]
    return {"value": flops / secs / 1e12, "unit": "TFLOP/s", "cores": cores, "kind": "port",
            "sample": f"{S}x{S} result tiles of [{label}], {k_note} ({npairs} tile pairs, {flops:.3e} flop, {secs:.2f} s); "
                      f"one single-threaded OpenBLAS DGEMM per pair on {cores} task threads; rate of the sample, not a full run"}, secs


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps_v = []
    base = None
    for _ in range(1 if args.warmup > 0 else 0):
        cpu_reference_leg(args.config, args.n, args.tile, budget_s=3.0)
    t_total = 0.0
    for _ in range(args.steps):
        base, secs = cpu_reference_leg(args.config, args.n, args.tile, budget_s=args.cpu_budget)
        steps_v.append(base["value"])
        t_total += secs
    v = float(np.mean(steps_v))
    base["value"] = v
    label = _cpu_plan(args.config, args.n, args.tile)[0]
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "TFLOP/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{label} FP64 ({args.config}); CPU leg timed on a bounded sample, rate reported",
                   "parallelism": f"1 process x {base['cores']} task threads (MADWorld-style), BLAS pinned to 1 thread"},
        "cpu_baseline": base,
        "e2e": {"value": v, "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ---------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="tadev", choices=["tadev", "reference"])
    ap.add_argument("--config", default="C2", choices=["C1", "C2", "C3", "C3m", "C4", "C4h", "C4r", "C5", "C5r"],
                    help="BASELINE.json config (default C2 = configs[1], the headline); C4r/C5r are reduced variants")
    ap.add_argument("--n", type=int, default=None)
    ap.add_argument("--tile", type=int, default=None)
    ap.add_argument("--cpu-budget", type=float, default=12.0)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-traffic", action="store_true", help="skip the child-ncu DRAM traffic measurement of the GEMM kernel")
    ap.add_argument("--spl", type=int, default=0, help="SUMMA steps per GEMM launch (0 = auto)")
    ap.add_argument("--parity-tiles", type=int, default=16, help="result tiles EVERY rank verifies (sampled elements)")
    args = ap.parse_args()
    if args.impl == "reference":
        args.steps = 3 if args.steps is None else args.steps
        args.warmup = 1 if args.warmup is None else args.warmup
        return run_reference(args)

    rank = int(os.environ.get("RANK", "0"))
    size = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert size == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={size}: launch with torch.distributed.run for N>1"

    import bench_workloads as W
    from tests import sampled_parity
    from tiledarray_b200 import Device, _lib
    from tiledarray_b200.tiledarray import ContEngine, DistArray, World, contraction_arrays

    dist = None
    if size > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = Device(local)
    world = World(device=dev, rank=rank, size=size)
    ContEngine.steps_per_launch = args.spl
    wl = W.build(world, args.config, args.n, args.tile)
    steps = wl.steps_hint if args.steps is None else args.steps
    warmup = wl.warmup_hint if args.warmup is None else args.warmup
    flops = wl.flops

    def barrier():
        dev.sync()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def allmax(vals):
        if dist is None:
            return list(vals)
        t = torch.tensor(list(vals), dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.tolist()

    def allsum(vals):
        if dist is None:
            return list(vals)
        t = torch.tensor(list(vals), dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t.tolist()

    # FP64 roofline denominator measured live on this GPU (not in MEASURED_PEAKS.json): burst = best short probe,
    # sustained = a 10x longer probe (clocks settled)
    peak_tf = max(dev.probe_fp64_peak(0, 40000)[0] for _ in range(2))
    peak_sustained = dev.probe_fp64_peak(0, 400000)[0]

    # the sampler (nvidia-smi -lms 200) is started BEFORE the warm-up steps: its start-up (NVML initialisation takes
    # driver locks for tens of ms and stalled a CUDA call of the first timed step when it was started after them) then
    # falls into the warm-up; it keeps sampling every 200 ms throughout the timed region
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(warmup):
        wl.step()
    barrier()
    launches0 = dev.launch_count()
    kernel_ms, gemm_ms, step_stats = [], [], None
    with dev.timer() as tm:  # CUDA events on the stream the driver launches on
        for _ in range(steps):
            step_stats = wl.step()
            kernel_ms.append(step_stats.device_ms)
            gemm_ms.append(step_stats.gemm_ms)
    launches = dev.launch_count() - launches0
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    t_ms = allmax([tm.ms])[0]
    value = flops * steps / (t_ms * 1e-3) / 1e12

    # ---- parity: EVERY rank verifies sampled elements of `parity_tiles` of its own result tiles (spread over its
    # row blocks, first and last tile included) against a host recomputation from regenerated operand elements
    par = sampled_parity.check_local_tiles(wl.c, wl.a, wl.b, wl.target, wl.lidx, wl.ridx,
                                           (None, None) if wl.name == "C3m" else wl.seeds, ntiles=args.parity_tiles)
    worst = allmax([par["worst_rel"]])[0]
    ptiles, pelems, ctiles, executed, npairs = allsum([par["tiles"], par["elements"], len(wl.c.tiles), step_stats.flops, step_stats.npairs])
    parity = {"worst_rel_frobenius": worst, "tiles_checked": int(ptiles), "elements_checked": int(pelems), "ranks": size,
              "result_tiles": int(ctiles), "tolerance": 1e-12, "ok": bool(worst < 1e-12)}
    if "pairs" in wl.extra:
        parity["tile_pairs"] = int(npairs)
        parity["tile_pairs_expected"] = wl.extra["pairs"]
        parity["ok"] = parity["ok"] and int(npairs) == wl.extra["pairs"]

    # bytes every launch must move at least once: this rank's operand tiles + its result tiles
    compulsory = sum(x._arena.nbytes for x in (wl.a, wl.b, wl.c) if getattr(x, "_arena", None) is not None)

    # ---- the permute kernel alone (config 5: north_star asks for its HBM fraction): one batched launch over this
    # rank's tiles of the left operand, device-timed
    permute_roofline = None
    if wl.name in ("C5", "C5r") and wl.a.tiles:
        ords = sorted(wl.a.tiles)[:512]
        ext = wl.a.trange.tile_extent(wl.a.trange.tile_index(ords[0]))
        nb = int(np.prod(ext)) * 8
        tmp = dev.alloc(nb * len(ords))
        src = np.array([wl.a.tiles[o].ptr for o in ords], dtype=np.uint64)
        dst = np.array([tmp.ptr + i * nb for i in range(len(ords))], dtype=np.uint64)
        perm = [0, 3, 1, 2]  # (i,k,a,c) -> (i,a,c,k) in image form: out.extent[perm[d]] = extent[d]
        dev.permute_batched_ptrs(ext, perm, 8, src, dst)
        dev.sync()
        with dev.timer() as tp:
            for _ in range(3):
                dev.permute_batched_ptrs(ext, perm, 8, src, dst)
        gbs = 3 * 2.0 * nb * len(ords) / (tp.ms * 1e-3) / 1e9
        tmp.free()
        hbm_peak = None
        mp = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(mp):
            hbm_peak = json.load(open(mp)).get("hbm_gbs")
        hbm_peak = hbm_peak or 6448.1
        permute_roofline = {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak,
                            "kernel": "transpose_fast_kernel (A tiles (16,16,64,64) -> (i,a,c,k))", "tiles": len(ords),
                            "bytes_per_launch": 2.0 * nb * len(ords),
                            "peak_source": "MEASURED_PEAKS.json hbm_gbs (driver-written) or the recipe's 6448.1 fallback"}

    # ---- e2e: the same contraction through the public API with HOST-resident (pinned) operands and result. Every step
    # moves all operand bytes host->device and the whole result device->host inside the timed region (the driver
    # streams operand panels window by window and returns the result in row blocks so the copies overlap the GEMMs).
    e2e = None
    if not args.no_e2e and wl.e2e_ok:
        lib = dev.lib
        a, b, c = wl.a, wl.b, wl.c
        a_h, b_h, _ = contraction_arrays(world, wl.target, wl.lidx, a.trange, wl.ridx, b.trange,
                                         None if a.shape.is_dense() else a.shape, None if b.shape.is_dense() else b.shape,
                                         memory=("host", "host"))
        a_h._allocate()
        b_h._allocate()
        assert a_h._arena.nbytes == a._arena.nbytes and b_h._arena.nbytes == b._arena.nbytes
        _lib.check(lib.tadev_memcpy_d2h(dev.ctx, a_h._arena.ptr, a._arena.ptr, a._arena.nbytes, dev.stream))
        _lib.check(lib.tadev_memcpy_d2h(dev.ctx, b_h._arena.ptr, b._arena.ptr, b._arena.nbytes, dev.stream))
        dev.sync()
        ref_ord = sorted(c.tiles)[0] if c.tiles else None
        c_ref_tile = c.find(ref_ord) if ref_ord is not None else None
        for x in (a, b, c):
            x.release()  # the device-resident copies are not used by the e2e leg
        c_h = DistArray(world, c.trange, memory="host")
        h2d = d2h = 0

        def e2e_step():
            prod = a_h[wl.lidx] * b_h[wl.ridx]
            if wl.mask is not None:
                prod = prod.set_shape(wl.mask)
            c_h[wl.target] = prod
            return ContEngine.last_stats

        for _ in range(min(warmup, 1) or 1):
            e2e_step()
        barrier()
        with dev.timer() as te:
            for _ in range(steps):
                st_e = e2e_step()
                h2d, d2h = st_e.h2d_bytes, st_e.d2h_bytes
        barrier()
        if c_ref_tile is not None:  # the streamed path must reproduce the device-resident result (window partial sums
            #                         are added in a different association, so equal to rounding, not bit for bit)
            t_e = c_h.find(ref_ord)
            assert np.linalg.norm(t_e - c_ref_tile) <= 1e-13 * np.linalg.norm(c_ref_tile), "e2e result differs from the device-resident run"
        te_ms = allmax([te.ms])[0]
        h2d, d2h = allsum([h2d, d2h])
        e2e = {"value": flops * steps / (te_ms * 1e-3) / 1e12, "unit": "TFLOP/s", "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "ms_per_step": te_ms / steps,
               "api": "DistArray(memory='host') operands + result; panels streamed, result returned in row blocks"}
        for x in (a_h, b_h, c_h):
            x.release()
    elif not wl.e2e_ok:
        e2e = None

    cpu = None
    if rank == 0 and size == 1 and not args.no_cpu:
        cpu, _ = cpu_reference_leg(args.config, args.n, args.tile, args.cpu_budget)

    if rank == 0:
        # dominant kernel: the grouped DMMA GEMM. algorithmic flop per launch = this rank's executed flop / launches;
        # duration = sum of the per-launch CUDA-event durations the driver records around every GEMM launch
        nl = max(1, step_stats.nlaunches)
        k_ms = float(np.mean(gemm_ms)) / nl if np.mean(gemm_ms) > 0 else float(np.mean(kernel_ms)) / nl
        achieved = step_stats.flops / nl / (k_ms * 1e-3) / 1e12
        traffic, traffic_note = None, "not measured (multi-GPU run or --no-traffic)"
        if size == 1 and not args.no_traffic:
            traffic, traffic_note = measure_gemm_traffic(args)
            if traffic is None:  # profiling not possible here: fall back to the committed ncu capture of the same config
                tpath = os.path.join(ROOT, "profiles", "gemm_traffic.json")
                if os.path.exists(tpath):
                    traffic = json.load(open(tpath)).get(wl.name)
                    traffic_note += "; value from profiles/gemm_traffic.json (committed ncu capture)"
        cfg = {"workload": wl.label, "flop_per_step": flops,
               "parallelism": f"SUMMA {wl.grid[0]}x{wl.grid[1]} process grid, 1 process/GPU",
               "l2": "operands per step >> 126 MB L2 (no flush needed)" if wl.name != "C1" else
                     "operands 268 MB per step > 126 MB L2 (no flush)",
               "parity_all_ranks": parity, "note": wl.note}
        if wl.apparent_flops:
            cfg["apparent_tflops_2N3"] = wl.apparent_flops * steps / (t_ms * 1e-3) / 1e12
        out = {
            "metric": METRIC, "value": value, "unit": "TFLOP/s", "n_gpus": size, "steps": steps, "warmup": warmup,
            "ms_per_step": t_ms / steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": cfg,
            "clocks": clocks, "gpu_launches": int(launches), "e2e": e2e,
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf,
                         "peak_sustained": peak_sustained, "frac_of_sustained": achieved / peak_sustained,
                         "traffic": traffic, "traffic_source": traffic_note,
                         "compulsory_bytes_per_launch": compulsory / nl if compulsory else None, "kernel": "gemm_grouped_f64_ws_kernel",
                         "peak_source": "DMMA.8x8x4 register-resident issue-rate probe measured in this run "
                                        "(tadev_probe_fp64_peak: burst = best of two 40k-iteration probes, sustained = one "
                                        "400k-iteration probe; MEASURED_PEAKS.json has no FP64 entry)",
                         "launches_per_step": int(nl), "ms_per_launch": k_ms,
                         "ms_per_launch_source": "CUDA events around each GEMM launch (rank 0), mean over the timed steps",
                         "gemm_share_of_step": float(np.mean(gemm_ms)) / float(np.mean(kernel_ms)) if np.mean(gemm_ms) > 0 else None,
                         "list_ms_per_step": step_stats.list_ms},
            "cpu_baseline": cpu,
        }
        if permute_roofline:
            out["roofline_permute"] = permute_roofline
        if not wl.e2e_ok:
            out["e2e_note"] = "not measured for this config: the operands do not fit a pinned host allocation this run may safely make"
        print(json.dumps(out))
    if dist is not None:
        dist.barrier()
    dev.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
