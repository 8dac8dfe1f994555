#!/usr/bin/env python3
"""bench.py -- headline benchmark of the blaze hot path on B200.

Metric (BASELINE.json): BLS12-381 MSM scalar-mults/sec at 2^26, HBM-resident points (configs[1]).
A "step" is one MSM over one batch of synthetic scalars against the resident point set.

  value   whole-job scalar-mults/s, device time (CUDA events on the library's launch stream), scalars
          already resident in HBM when the timed region starts
  e2e     the same metric through the reference-facing call order
          (initialize -> start_process -> set_data(host scalars) -> wait_result -> result) with pinned HOST
          buffers: the H2D copy of the step's scalars and the D2H read of the result are inside the
          timed region
  roofline  the dominant kernel (k_accumulate): algorithmic bytes / CUDA-event duration vs measured HBM peak
  cpu_baseline  the oracle's arkworks-0.3-style Pippenger ("port") on the box's host cores, bounded sample

Multi-GPU (torchrun, one rank per GPU): the MSM is point-sharded (rank g owns points/scalars
[g N/G, (g+1) N/G)), no data-path collective; the only exchange is an all-gather of the G 144-byte
partial results, summed by bz_msm_combine_results.  scaling = "strong" (total work fixed at 2^26).

Secondary sections of the same JSON line (N = 1 only unless noted): `ntt` (2^27 NTT ms, device-resident; also at
N > 1 as the four-step across the ranks; `ntt.e2e` = through the client calls with pinned host buffers), `dma_mode`
(BASELINE.json configs[2]: BN254 2^24 with points AND scalars streamed from host memory every call).

`--impl reference` times the reference's CPU definition of the path (the oracle port: the reference
itself is Rust + an FPGA bitstream and cannot run here) on the host cores, same metric and config.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "BLS12-381 MSM scalar-mults/sec at 2^26"
UNIT = "scalar-mults/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--log-n", type=int, default=26, help="log2 of the MSM size (default: the headline 2^26)")
    ap.add_argument("--cpu-sample-log-n", type=int, default=18)
    ap.add_argument("--no-verify", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--curve", default="BLS381", choices=["BLS381", "BLS377", "BN254"],
                    help="headline metric is BLS381; the other curves run the same workload (e.g. configs[4])")
    ap.add_argument("--no-ntt", action="store_true", help="skip the secondary metric (2^27 NTT ms)")
    ap.add_argument("--ntt-log-n", type=int, default=27)
    ap.add_argument("--no-dma", action="store_true", help="skip the DMA-mode measurement (BN254 2^24, configs[2])")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe).  ONE sampler for
    the whole job (rank 0 polls the N GPUs of the run every 500 ms): a poller per rank at 5 Hz measurably slows
    the CUDA driver calls of all ranks at N = 8."""

    def __init__(self, indices):
        self.indices = list(indices)
        self.proc = None
        self.lines = []

    def start(self):
        q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", ",".join(str(i) for i in self.indices), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "500"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.3)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "gpus_sampled": len(self.indices)}


def cpu_port_rate(log_n, threads=0, seed=900):
    """Oracle Pippenger (arkworks-0.3-style port) on a 2^log_n sample of the same workload."""
    from oracle import capi
    from oracle.py import curves
    from util import chain_points, random_scalars
    c = curves.BLS12_381
    n = 1 << log_n
    pts, p0, q = chain_points(c, n, seed=seed)
    sc = random_scalars(c, n, seed=seed + 1)
    th = threads or capi.hw_threads()
    t = time.perf_counter()
    capi.msm_pippenger("BLS12_381", pts, sc, n, th)
    dt = time.perf_counter() - t
    return n / dt, th, dt


class _DevView:
    """torch view of raw device memory owned by the library (via __cuda_array_interface__)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def ntt_section(args, bz, torch, dist, dc, rank, world, local):
    """Secondary metric of BASELINE.json: 2^27 NTT over BLS12-381 Fr, ms (device-resident data).
    N = 1: NTTClient (3 Stockham passes).  N > 1: four-step across the ranks, exchange fused into the
    last column pass (peer stores over NVLink), one host barrier between the two steps."""
    log_n = args.ntt_log_n
    n = 1 << log_n
    reps = max(3, args.steps)
    gen = torch.Generator(device="cuda")
    gen.manual_seed(1234 + rank)
    if world == 1:
        t = bz.NTTClient.new_ex(dc, 2, log_n, False)
        t.initialize()
        view = torch.as_tensor(_DevView(t.slot_device_ptr(0), n * 32), device="cuda")
        view.copy_(torch.randint(0, 256, (n * 32,), dtype=torch.uint8, device="cuda", generator=gen))
        view.view(n, 32)[:, 31] &= 0x3f          # canonical elements (< 2^254 < r)
        torch.cuda.synchronize()
        ms = []
        for i in range(reps + 2):
            t.start_process(0)
            t.wait_result()
            if i >= 2:
                ms.append(t.phase_times()["total"])
        passes = t.phase_times()["passes"]
        # end to end through the client calls with pinned HOST buffers (4 GiB in + 4 GiB out per transform at 2^27):
        # serial = set_data -> start_process -> wait_result -> result; pipelined = the reference's double-buffer cycle
        # (integration_ntt.rs:103-136) from one host thread: the transform of slot 1-h runs while slot h is read out
        # and refilled
        e2e = None
        try:
            hin = torch.empty(n * 32, dtype=torch.uint8).pin_memory()
            hout = torch.empty(n * 32, dtype=torch.uint8).pin_memory()
            hin.copy_(view)
            bi, bo = (hin.data_ptr(), n * 32), (hout.data_ptr(), n * 32)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(2):
                t.set_data(bz.NTTInput(0, bi))
                t.start_process(0)
                t.wait_result()
                t.result(0, out=bo)
            serial_ms = 1e3 * (time.perf_counter() - t0) / 2
            k = 4
            t.set_data(bz.NTTInput(0, bi))
            t.start_process(0)
            t.set_data(bz.NTTInput(1, bi))
            t.wait_result()
            h = 0
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(k):
                t.start_process(1 - h)
                t.result(h, out=bo)
                t.set_data(bz.NTTInput(h, bi))
                t.wait_result()
                h = 1 - h
            pipe_ms = 1e3 * (time.perf_counter() - t0) / k
            e2e = {"serial_ms": serial_ms, "pipelined_ms": pipe_ms, "h2d_bytes": n * 32, "d2h_bytes": n * 32,
                   "note": "PCIe bound: 2 x %.1f GiB per transform; one host thread, so H2D and D2H do not overlap each other" % (n * 32 / 2**30)}
            del hin, hout
        except Exception as ex:
            e2e = {"error": repr(ex)}
        t.close()
        ms_val = sum(ms) / len(ms)
        layout = "natural order in / natural order out, in place in slot 0"
    else:
        def exchange(h):
            out = [None] * world
            dist.all_gather_object(out, h)
            return out
        t = bz.DistributedNTT(dc, log_n, rank, world, exchange=exchange, barrier=dist.barrier)
        a_ptr, o_ptr, per = t.buffers()
        view = torch.as_tensor(_DevView(a_ptr, per * 32), device="cuda")
        view.copy_(torch.randint(0, 256, (per * 32,), dtype=torch.uint8, device="cuda", generator=gen))
        view.view(per, 32)[:, 31] &= 0x3f
        torch.cuda.synchronize()
        walls = []
        for i in range(reps + 2):
            dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            t.run()
            dist.barrier()
            if i >= 2:
                walls.append(time.perf_counter() - t0)
        tm = t.times()
        pl = t.plan()
        v = torch.tensor([sum(walls) / len(walls), tm["step1_ms"], tm["step3_ms"]], dtype=torch.float64, device="cuda")
        dist.all_reduce(v, op=dist.ReduceOp.MAX)
        ms_val = float(v[0]) * 1e3
        passes = pl["column_passes"] + pl["row_passes"]
        layout = ("N = 2^%d x 2^%d; " % (pl["log_n1"], pl["log_n2"]) + "rank g holds column slab in[j1*N2 + g*C + c] in, X[(h*T+t) + N1*k2] out (strided slabs); "
                  "step1 %.2f ms + step3 %.2f ms device time, rest = 2 host barriers" % (float(v[1]), float(v[2])))
        t.close()
    hbm_peak, _ = peaks()
    npass = passes if passes else 4
    gbs = npass * 2 * n * 32 / world / (ms_val / 1e3) / 1e9
    return {"metric": "2^%d NTT over BLS12-381 Fr, ms" % log_n, "ms": ms_val, "n_gpus": world, "passes": npass,
            "e2e": e2e if world == 1 else None,
            "layout": layout, "semantics": "arkworks Radix2EvaluationDomain::fft (natural in/out), forward",
            "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s per GPU", "frac": gbs / hbm_peak,
                         "algorithmic_bytes": npass * 2 * n * 32}}


def dma_section(args, bz, torch, dc):
    """BASELINE.json configs[2]: BN254 MSM 2^24 in DMA mode -- points AND scalars come from (pinned) host memory
    with every call (msm_api.rs:175-202), so the H2D copies and the canonical->Montgomery table build are
    inside every timed step.  Checked against the oracle's closed form."""
    import numpy as np
    from oracle import capi
    from oracle.py import curves
    from util import random_scalars, seed_points
    c = curves.BN254
    log_n = 24
    n = 1 << log_n
    p0, q = seed_points(c, 91)
    gen = bz.MSMClient.new(bz.MSMInit(bz.PointMemoryType.HBM, False, bz.Curve.BN254), dc)
    try:
        gen.generate_chain_points(p0 + q, 0, n, 0x100000000, 0)
        pts_pinned = torch.empty(n * c.point_size, dtype=torch.uint8).pin_memory()
        pts_pinned.numpy()[:] = np.frombuffer(gen.get_data_from_hbm(n * c.point_size, 0x100000000, 0), dtype=np.uint8)
    finally:
        gen.close()
    sc_np = random_scalars(c, n, seed=92)
    sc_pinned = torch.empty(n * 32, dtype=torch.uint8).pin_memory()
    sc_pinned.numpy()[:] = sc_np
    m = bz.MSMClient.new(bz.MSMInit(bz.PointMemoryType.DMA, False, bz.Curve.BN254), dc)
    try:
        params = bz.MSMParams(n, None)
        walls, dev = [], []
        res = None
        for i in range(2 + max(3, args.steps)):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            m.initialize(params)
            m.start_process()
            m.set_data(bz.MSMInput((pts_pinned.data_ptr(), n * c.point_size), (sc_pinned.data_ptr(), n * 32), params))
            m.wait_result()
            res = m.result().result
            if i >= 2:
                walls.append(time.perf_counter() - t0)
                dev.append(m.phase_times()["total"])
        ok = bool(res == capi.chain_expected("BN254", p0, q, sc_np, n))
        ms = 1e3 * sum(walls) / len(walls)
        return {"workload": "BN254 MSM 2^24, DMA mode (configs[2]): points + scalars streamed from pinned host memory every call",
                "ms_per_call": ms, "scalar_mults_per_s": n / (ms / 1e3), "device_pipeline_ms": sum(dev) / len(dev),
                "h2d_bytes_per_call": n * (c.point_size + 32), "verified_bit_exact_vs_oracle_closed_form": ok,
                "plan": m.plan_info()}
    finally:
        m.close()


def run_reference(args, rank, world):
    """CPU arm: the reference's own definition of the path (oracle port), bounded sample per step."""
    if rank != 0:
        return
    from oracle import capi
    capi.build()
    ln = args.cpu_sample_log_n
    for _ in range(max(1, min(args.warmup, 1))):
        cpu_port_rate(min(ln, 14))
    rates, times = [], []
    # bounded sample per step: size it so that K steps take a couple of minutes at most on this box
    r0, th, dt0 = cpu_port_rate(16)
    budget = 120.0 / max(1, args.steps + 1)
    while ln < 24 and (1 << (ln + 1)) / r0 < budget:
        ln += 1
    for s in range(args.steps):
        r, th, dt = cpu_port_rate(ln, seed=900 + s)
        rates.append(r)
        times.append(dt)
    total_n = args.steps * (1 << ln)
    value = total_n / sum(times)
    sample = "2^%d-point prefix of the 2^%d workload per step (chain points P0+iQ, uniform scalars)" % (ln, args.log_n)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u32 limbs (381-bit Fq)",
        "data": "synthetic",
        "config": {"workload": "BLS12-381 MSM 2^%d, arkworks-0.3-style Pippenger restated in C++ (oracle port), "
                               "host CPU" % args.log_n, "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": th, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import numpy as np
    import torch
    import blaze_b200 as bz
    from blaze_b200._lib import lib
    from oracle.py import curves
    from util import random_scalars, seed_points

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    cname = {"BLS381": "BLS12_381", "BLS377": "BLS12_377", "BN254": "BN254"}[args.curve]
    c = curves.CURVES[cname]
    N = 1 << args.log_n
    per = N // world
    first = rank * per
    dc = bz.DriverClient(str(local), bz.DriverConfig.driver_client_cfg(bz.CardType.B200))
    m = bz.MSMClient.new(bz.MSMInit(bz.PointMemoryType.HBM, False, getattr(bz.Curve, args.curve)), dc)
    p0, q = seed_points(c, 2026)
    HBM_ADDR = 0
    # resident points: P_i = P0 + i*Q for this rank's index range, generated on the device (untimed)
    m.generate_chain_points(p0 + q, first, per, HBM_ADDR, 0)
    params = bz.MSMParams(per, (HBM_ADDR, 0))
    if world > 1:
        m.set_raw_result(True)   # shards stay projective; the combine normalises once

    # scalars: pinned host copy (for e2e) + device copy (for value)
    sc_np = random_scalars(c, N, seed=4242)[first * 32:(first + per) * 32]
    sc_pinned = torch.empty(per * 32, dtype=torch.uint8).pin_memory()
    sc_pinned.numpy()[:] = sc_np
    sc_dev = sc_pinned.cuda(non_blocking=False)
    torch.cuda.synchronize()

    def step_resident():
        m.initialize(params)
        m.start_process()
        m.set_scalars_device(sc_dev.data_ptr(), params)
        m.wait_result()
        return m.result().result

    def step_e2e():
        m.initialize(params)
        m.start_process()
        m.set_data(bz.MSMInput(None, (sc_pinned.data_ptr(), per * 32), params))
        m.wait_result()
        return m.result().result

    def combine(partial):
        if world == 1:
            return partial
        t = torch.frombuffer(bytearray(partial), dtype=torch.uint8).cuda()
        gathered = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(gathered, t)
        recs = b"".join(bytes(g.cpu().numpy()) for g in gathered)
        return m.combine_results(recs, world)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (also builds the Montgomery table and the workspace)
    res = None
    warm_s = []
    for _ in range(max(args.warmup, 3)):
        t0 = time.perf_counter()
        res = combine(step_resident())
        warm_s.append(round(time.perf_counter() - t0, 3))   # step 2 contains the one-off build of the merged table

    # ---- verification at full size: closed form of the chain workload (bit-exact)
    verified = None
    if not args.no_verify and rank == 0:
        from oracle import capi
        capi.build()
        full_sc = random_scalars(c, N, seed=4242)
        exp = capi.chain_expected(cname, p0, q, full_sc, N)
        verified = bool(res == exp)
        if not verified:
            raise SystemExit("bench: GPU result differs from the oracle closed form -- number is INVALID")

    # ---- timed region 1: device-resident (value)
    sampler = ClockSampler(range(world)) if rank == 0 else None
    launches0 = lib().bz_kernel_launch_count()
    barrier()
    if sampler:
        sampler.start()
    t0 = time.perf_counter()
    dev_ms, acc_ms, sort_ms, red_ms = [], [], [], []
    for _ in range(args.steps):
        combine(step_resident())
        pt = m.phase_times()
        dev_ms.append(pt["total"])
        acc_ms.append(pt["accumulate"])
        sort_ms.append(pt["sort"])
        red_ms.append(pt["reduce"])
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop() if sampler else None
    launches = lib().bz_kernel_launch_count() - launches0

    # ---- timed region 2: end to end with host buffers, strictly serial calls
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        combine(step_e2e())
    barrier()
    wall_e2e_serial = time.perf_counter() - t0

    # ---- timed region 3: end to end with host buffers, two tasks in flight (the reference's task queue:
    # start_process/set_data of task k+1 are issued before wait_result/result of task k, so the H2D
    # copy of the next step's scalars overlaps the kernels of the current one); every step still copies
    # its 2 GiB of scalars from pinned host memory and reads its result back inside the timed region
    def enqueue():
        m.initialize(params)
        m.start_process()
        m.set_data(bz.MSMInput(None, (sc_pinned.data_ptr(), per * 32), params))

    barrier()
    t0 = time.perf_counter()
    enqueue()
    for k in range(args.steps):
        if k + 1 < args.steps:
            enqueue()
        m.wait_result()
        combine(m.result().result)
    barrier()
    wall_e2e = time.perf_counter() - t0

    # max over ranks (device time per step, wall times)
    vals = torch.tensor([sum(dev_ms) / len(dev_ms), wall, wall_e2e, sum(acc_ms) / len(acc_ms), wall_e2e_serial],
                        dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
    dev_step_ms, wall, wall_e2e, acc_step_ms, wall_e2e_serial = [float(x) for x in vals.cpu()]
    # whole-job step time: wall clock of the K steps bracketed by barrier + synchronize (max over ranks);
    # the CUDA-event time of the device pipeline alone is reported beside it in config.device_ms_per_step
    ms_per_step = 1e3 * wall / args.steps
    value = N / (ms_per_step / 1e3)
    e2e_value = N / (wall_e2e / args.steps)

    ntt = None
    if not args.no_ntt:
        m_plan = m.plan_info()
        try:
            ntt = ntt_section(args, bz, torch, dist, dc, rank, world, local)
        except Exception as e:     # the headline line must still be printed
            ntt = {"error": repr(e)}

    dma = None
    if world == 1 and not args.no_dma:
        try:
            dma = dma_section(args, bz, torch, dc)
        except Exception as e:
            dma = {"error": repr(e)}

    if rank == 0:
        plan = m.plan_info()
        W, cbits = plan["windows"], plan["c"]
        hbm_peak, peak_src = peaks()
        # algorithmic bytes of the accumulate sweep: (point_size + 4 B index) per (scalar, window)
        alg_bytes = (c.point_size + 4) * per * W
        achieved = alg_bytes / (acc_step_ms / 1e3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "r1_traffic.json")
        if os.path.exists(tp):
            try:
                tj = json.load(open(tp))
                # the ncu capture is of the 2^26 / 13-window launch; other shard sizes scale with the algorithmic bytes
                traffic = tj.get("k_accumulate_dram_bytes_per_launch") * alg_bytes / tj.get("algorithmic_bytes_per_launch")
            except Exception:
                traffic = None
        line = {
            "metric": METRIC.replace("BLS12-381", cname.replace("_", "-")).replace("2^26", "2^%d" % args.log_n),
            "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None,
            "dtype": "u32 limbs (%d-bit Fq Montgomery, %d-bit Fr)" % (c.q.bit_length(), c.r.bit_length()),
            "data": "synthetic",
            "config": {
                "workload": "%s MSM 2^%d, HBM-resident points (configs[1]); points P0+iQ generated on device, "
                            "uniform random canonical scalars" % (cname.replace("_", "-"), args.log_n),
                "precompute_factor": 1, "window_bits": cbits, "windows": W, "segment": plan["segment"],
                "bucket_sets": plan["bucket_sets"],
                "resident_table": ("window-merged: 2^(c*w)*P_i for the %d digit windows derived once from the resident "
                                   "points (%d MiB of HBM, built in warm-up, not timed), all windows share one bucket set"
                                   % (W, plan["merged_table_mib"])) if plan["merged_table"] else "points only (Montgomery form)",
                "parallelism": "point-sharded x%d" % world,
                "l2": "inputs (2 GiB scalars + 6 GiB points per 2^26) exceed the 126 MB L2; no flush needed",
                "verified_bit_exact_vs_oracle_closed_form": verified,
                "phase_ms": {"sort": sum(sort_ms) / len(sort_ms), "accumulate": sum(acc_ms) / len(acc_ms),
                             "reduce": sum(red_ms) / len(red_ms)},
                "device_ms_per_step": dev_step_ms,
                "warmup_step_s": warm_s,
            },
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": per * 32 * world,
                    "d2h_bytes_per_step": c.result_point_size * world, "ms_per_step": 1e3 * wall_e2e / args.steps,
                    "mode": "two tasks in flight through the client's task queue (H2D of step k+1 overlaps step k)",
                    "serial_value": N / (wall_e2e_serial / args.steps),
                    "serial_ms_per_step": 1e3 * wall_e2e_serial / args.steps},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "k_accumulate<%s>" % cname, "achieved": achieved, "peak": hbm_peak,
                         "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": traffic,
                         "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes,
                         "note": "the sweep is integer-multiplier bound, not HBM bound: ncu shows the fmaheavy pipe "
                                 "(IMAD.WIDE.U32, 4 issue cycles each) ~82% busy and DRAM ~5% "
                                 "(profiles/r1_ncu_k_accumulate_merged_2p22.txt)"},
        }
        if ntt is not None:
            line["ntt"] = ntt
        if dma is not None:
            line["dma_mode"] = dma
        if world == 1 and not args.no_cpu_baseline:
            from oracle import capi
            capi.build()
            # bounded sample: aim at 10-30 s of CPU work on this box (probe at 2^16, then size it)
            r0, th, dt0 = cpu_port_rate(16)
            ln = args.cpu_sample_log_n
            while ln < 24 and (1 << (ln + 1)) / r0 < 20.0:
                ln += 1
            args.cpu_sample_log_n = ln
            r, th, dt = cpu_port_rate(ln)
            line["cpu_baseline"] = {"value": r, "unit": UNIT, "cores": th, "kind": "port",
                                    "sample": "one 2^%d-point prefix of the workload, %.1f s (oracle: arkworks-0.3-style "
                                              "Pippenger, C++)" % (args.cpu_sample_log_n, dt)}
        print(json.dumps(line), flush=True)
    m.close()
    dc.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
