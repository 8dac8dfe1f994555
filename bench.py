#!/usr/bin/env python3
"""bench.py -- headline benchmark of the blaze hot path on B200.

Metric (BASELINE.json): BLS12-381 MSM scalar-mults/sec at 2^26, HBM-resident points (configs[1]).
A "step" is one MSM over one batch of synthetic scalars against the resident point set.

  value   whole-job scalar-mults/s: K steps bracketed by barrier + synchronize (max over ranks), scalars already
          resident in HBM when the timed region starts, two tasks in flight through the client's task queue
          (config.serial_value: every result awaited before the next task is queued); the per-phase device times
          come from CUDA events the library records on its launch streams
  e2e     the same metric through the reference-facing call order
          (initialize -> start_process -> set_data(host scalars) -> wait_result -> result) with pinned HOST
          buffers: the H2D copy of the step's scalars and the D2H read of the result are inside the
          timed region
  roofline  the dominant kernel (k_accumulate): algorithmic bytes / CUDA-event duration vs measured HBM peak
  cpu_baseline  the oracle's arkworks-0.3-style Pippenger ("port") on the box's host cores, bounded sample

Multi-GPU (torchrun, one rank per GPU): every rank opens a ranked DriverClient (bz_dclient_comm_init; torch.distributed
only hands the NCCL id around).  The MSM is point-sharded (rank g owns points/scalars [g N/G, (g+1) N/G)); the library
all-gathers the projective partial records with NCCL and sums them on the device (its tail stream), so result() is the full sum on every
rank.  scaling = "strong" (total work fixed at 2^26).

Other sections of the same JSON line, each VERIFIED inside the run: `config5` (N > 1: BLS12-377 2^26 across the GPUs,
BASELINE.json configs[4]), `ntt` (2^27 NTT ms, device-resident; N > 1: four-step across the ranks; outputs checked against
their definition at spot positions; `ntt.e2e` = through the client calls with pinned host buffers), `dma_mode` (configs[2]:
BN254 2^24 with points AND scalars streamed from host memory every call), `precompute_x8` (the reference's x8 precomputed wire
format in HBM mode), `poseidon` (TreeC tree of height 7), `config.plain_table_value` (no window-merged table).

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
    ap.add_argument("--cpu-sample-log-n", type=int, default=None, help="CPU arm sample size (default 2^22 everywhere)")
    ap.add_argument("--no-verify", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--curve", default="BLS381", choices=["BLS381", "BLS377", "BN254"],
                    help="headline metric is BLS381; the other curves run the same workload (e.g. configs[4])")
    ap.add_argument("--no-ntt", action="store_true", help="skip the secondary metric (2^27 NTT ms)")
    ap.add_argument("--ntt-log-n", type=int, default=27)
    ap.add_argument("--no-dma", action="store_true", help="skip the DMA-mode measurement (BN254 2^24, configs[2])")
    ap.add_argument("--no-config5", action="store_true", help="N > 1: skip the BLS12-377 run (configs[4])")
    ap.add_argument("--no-precompute", action="store_true", help="skip the x8 precomputed-bases measurement")
    ap.add_argument("--precompute-log-n", type=int, default=24)
    ap.add_argument("--no-poseidon", action="store_true")
    ap.add_argument("--poseidon-height", type=int, default=7)
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


# ---- synthetic field elements by index (same stream on the GPU with torch and on the host with numpy), so that a
# rank can fill its strided slab on the device and rank 0 can rebuild the whole vector on the host for the check
_SM = (0x9E3779B97F4A7C15, 0xBF58476D1CE4E5B9, 0x94D049BB133111EB)


def _i64(v):
    return v - (1 << 64) if v >= (1 << 63) else v


def elems_torch(torch, idx, seed):
    """idx: int64 tensor of element indices -> uint8 tensor [len, 32]: canonical elements (< 2^254 < r)."""
    words = []
    for k in range(4):
        z = idx * 4 + (k + seed * 0x1000003)
        z = z + _i64(_SM[0])
        z = (z ^ ((z >> 30) & ((1 << 34) - 1))) * _i64(_SM[1])
        z = (z ^ ((z >> 27) & ((1 << 37) - 1))) * _i64(_SM[2])
        z = z ^ ((z >> 31) & ((1 << 33) - 1))
        if k == 3:
            z = z & ((1 << 62) - 1)
        words.append(z)
    return torch.stack(words, dim=1).contiguous().view(torch.uint8).view(-1, 32)


def elems_numpy(n, seed):
    import numpy as np
    out = np.empty((n, 4), dtype=np.uint64)
    idx = np.arange(n, dtype=np.uint64)
    with np.errstate(over="ignore"):
        for k in range(4):
            z = idx * np.uint64(4) + np.uint64((k + seed * 0x1000003) & 0xFFFFFFFFFFFFFFFF)
            z = z + np.uint64(_SM[0])
            z = (z ^ (z >> np.uint64(30))) * np.uint64(_SM[1])
            z = (z ^ (z >> np.uint64(27))) * np.uint64(_SM[2])
            z = z ^ (z >> np.uint64(31))
            if k == 3:
                z = z & np.uint64((1 << 62) - 1)
            out[:, k] = z
    return out.view(np.uint8).reshape(-1)


def ntt_section(args, bz, torch, dist, dc, rank, world, local):
    """Secondary metric of BASELINE.json: 2^27 NTT over BLS12-381 Fr, ms (device-resident data), VERIFIED: outputs at
    spot positions are compared with their definition out[k] = sum_j in[j] w^(jk) (oracle Horner, O(n) per point).
    N = 1: NTTClient (3 Stockham passes).  N > 1: four-step across the ranks, exchange fused into the last column pass
    (peer stores over NVLink); handles and both barriers go through the ranked DriverClient's NCCL communicator."""
    import numpy as np
    log_n = args.ntt_log_n
    n = 1 << log_n
    reps = max(3, args.steps)
    seed = 31
    verified = None
    if world == 1:
        t = bz.NTTClient.new_ex(dc, 2, log_n, False)
        t.initialize()
        view = torch.as_tensor(_DevView(t.slot_device_ptr(0), n * 32), device="cuda").view(n, 32)
        step = 1 << 24
        for lo in range(0, n, step):
            view[lo:lo + step] = elems_torch(torch, torch.arange(lo, min(n, lo + step), dtype=torch.int64, device="cuda"), seed)
        torch.cuda.synchronize()
        hin = torch.empty(n * 32, dtype=torch.uint8).pin_memory()
        hout = torch.empty(n * 32, dtype=torch.uint8).pin_memory()
        hin.copy_(view.view(-1))
        bi, bo = (hin.data_ptr(), n * 32), (hout.data_ptr(), n * 32)
        # ---- verification of the first transform (input = the synthetic vector)
        t.start_process(0)
        t.wait_result()
        t.result(0, out=bo)
        if not args.no_verify:
            from oracle import capi
            capi.build()
            ks = sorted(set(k % n for k in (0, 1, n - 1, n // 2, 511, 512, (1 << 18) - 1, 1 << 18, 0x2AAAAAA, 0x5555555,
                                            123456789, (1 << 26) + (1 << 9) + 1)))
            host_in = hin.numpy()
            assert bytes(host_in[:64]) == bytes(elems_numpy(2, seed))          # the host generator is the same stream
            exp = capi.ntt_eval("BLS12_381", host_in, log_n, ks)
            ho = hout.numpy()
            got = [int.from_bytes(bytes(ho[32 * k:32 * k + 32]), "little") for k in ks]
            verified = bool(got == exp)
            if not verified:
                raise SystemExit("bench: NTT output differs from the oracle's definition -- number is INVALID")
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
            # the same two slots driven by THREE host threads (the clients are Send + Sync like the reference's, dclient.rs:28-46):
            # a feeder calls set_data, the main thread start_process / wait_result, a drainer result.  set_data / result /
            # wait_result block without the client lock, so the H2D of transform i+1 runs while the D2H of transform i-1
            # is still in flight (PCIe is full duplex); a slot is refilled only after its previous result has been read.
            hout2 = torch.empty(n * 32, dtype=torch.uint8).pin_memory()
            outs = ((hout.data_ptr(), n * 32), (hout2.data_ptr(), n * 32))
            kt = 6
            in_ready = [threading.Semaphore(0) for _ in range(kt)]
            cmp_done = [threading.Semaphore(0) for _ in range(kt)]
            out_done = [threading.Semaphore(0) for _ in range(kt)]
            errs = []

            WAIT = 120   # seconds: a failed thread must not leave the others waiting for ever

            def take(sem):
                if errs or not sem.acquire(timeout=WAIT):
                    raise RuntimeError("threaded NTT cycle: a peer thread failed or timed out")

            def feeder():
                try:
                    for i in range(kt):
                        if i >= 2:
                            take(out_done[i - 2])
                        t.set_data(bz.NTTInput(i % 2, bi))
                        in_ready[i].release()
                except Exception as ex:      # pragma: no cover
                    errs.append(ex)

            def drainer():
                try:
                    for i in range(kt):
                        take(cmp_done[i])
                        t.result(i % 2, out=outs[i % 2])
                        out_done[i].release()
                except Exception as ex:      # pragma: no cover
                    errs.append(ex)
            torch.cuda.synchronize()
            tf, td = threading.Thread(target=feeder, daemon=True), threading.Thread(target=drainer, daemon=True)
            t0 = time.perf_counter()
            tf.start()
            td.start()
            try:
                for i in range(kt):
                    take(in_ready[i])
                    t.start_process(i % 2)
                    t.wait_result()
                    cmp_done[i].release()
            except Exception as ex:      # pragma: no cover
                errs.append(ex)
            for sem in in_ready + cmp_done + out_done:   # whatever happened: nobody stays blocked
                sem.release()
            tf.join(WAIT)
            td.join(WAIT)
            torch.cuda.synchronize()
            thr_ms = 1e3 * (time.perf_counter() - t0) / kt
            thr_ok = (not errs) and bool(torch.equal(hout, hout2))       # same input in both slots: same transform out
            if verified is not None and thr_ok:
                ho = hout2.numpy()
                thr_ok = [int.from_bytes(bytes(ho[32 * kk:32 * kk + 32]), "little") for kk in ks] == exp
            e2e = {"serial_ms": serial_ms, "pipelined_ms": pipe_ms, "threaded_ms": thr_ms, "threaded_outputs_ok": bool(thr_ok),
                   "h2d_bytes": n * 32, "d2h_bytes": n * 32,
                   "note": "PCIe bound: 2 x %.1f GiB per transform.  serial / pipelined: ONE host thread, so the blocking set_data and "
                           "result calls keep H2D and D2H from overlapping; threaded: a feeder and a drainer thread beside the main "
                           "one (%d transforms incl. pipeline fill), both PCIe directions busy" % (n * 32 / 2**30, kt)}
            del hin, hout, hout2
        except Exception as ex:
            e2e = {"error": repr(ex)}
        t.close()
        ms_val = sum(ms) / len(ms)
        layout = "natural order in / natural order out, in place in slot 0"
    else:
        t = bz.DistributedNTT(dc, log_n, rank, world)      # ranked client: handle exchange + barriers over its communicator
        a_ptr, o_ptr, per = t.buffers()
        pl = t.plan()
        N1, N2 = 1 << pl["log_n1"], 1 << pl["log_n2"]
        C, T = N2 // world, N1 // world
        slab = torch.as_tensor(_DevView(a_ptr, per * 32), device="cuda").view(per, 32)

        def fill():
            rows = max(1, (1 << 22) // C)
            for j1 in range(0, N1, rows):
                r = torch.arange(j1, min(N1, j1 + rows), dtype=torch.int64, device="cuda")
                idx = (r[:, None] * N2 + rank * C + torch.arange(C, dtype=torch.int64, device="cuda")[None, :]).reshape(-1)
                slab[j1 * C:j1 * C + idx.numel()] = elems_torch(torch, idx, seed)
        fill()
        torch.cuda.synchronize()
        t.run()
        # ---- verification: two outputs of every rank's block O[k2][t] = X[(rank T + t) + N1 k2]
        if not args.no_verify:
            oview = torch.as_tensor(_DevView(o_ptr, per * 32), device="cuda").view(per, 32)
            picks = [(0, 0), (N2 - 1 - rank, T - 1), ((7919 * (rank + 1)) % N2, (104729 * (rank + 3)) % T)]
            mine = torch.zeros(len(picks), 40, dtype=torch.uint8, device="cuda")
            for i, (k2, tt) in enumerate(picks):
                kglob = rank * T + tt + N1 * k2
                mine[i, :32] = oview[k2 * T + tt]
                mine[i, 32:] = torch.tensor(list(int(kglob).to_bytes(8, "little")), dtype=torch.uint8, device="cuda")
            allv = [torch.zeros_like(mine) for _ in range(world)]
            dist.all_gather(allv, mine)
            ok = torch.ones(1, dtype=torch.int32, device="cuda")
            if rank == 0:
                from oracle import capi
                capi.build()
                host_in = elems_numpy(n, seed)
                recs = torch.cat(allv).cpu().numpy()
                ks = [int.from_bytes(bytes(r[32:40]), "little") for r in recs]
                exp = capi.ntt_eval("BLS12_381", host_in, log_n, ks)
                got = [int.from_bytes(bytes(r[:32]), "little") for r in recs]
                ok[0] = 1 if got == exp else 0
                del host_in
            dist.broadcast(ok, src=0)
            verified = bool(int(ok[0]))
            if not verified:
                raise SystemExit("bench: distributed NTT output differs from the oracle's definition -- number is INVALID")
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
        v = torch.tensor([sum(walls) / len(walls), tm["step1_ms"], tm["step3_ms"]], dtype=torch.float64, device="cuda")
        dist.all_reduce(v, op=dist.ReduceOp.MAX)
        ms_val = float(v[0]) * 1e3
        passes = pl["column_passes"] + pl["row_passes"]
        e2e = None
        layout = ("N = 2^%d x 2^%d; " % (pl["log_n1"], pl["log_n2"]) + "rank g holds column slab in[j1*N2 + g*C + c] in, X[(h*T+t) + N1*k2] out (strided slabs); "
                  "step1 %.2f ms + step3 %.2f ms device time; barriers = 4-byte NCCL all-reduces on the stream" % (float(v[1]), float(v[2])))
        t.close()
    hbm_peak, _ = peaks()
    npass = passes if passes else 4
    gbs = npass * 2 * n * 32 / world / (ms_val / 1e3) / 1e9
    return {"metric": "2^%d NTT over BLS12-381 Fr, ms" % log_n, "ms": ms_val, "n_gpus": world, "passes": npass,
            "verified": verified,
            "verification": "outputs at spot positions vs their definition sum_j in[j] w^(jk) (oracle Horner over the whole input)",
            "e2e": e2e,
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
        exp = capi.chain_expected("BN254", p0, q, sc_np, n)
        ok = bool(res == exp)
        ms = 1e3 * sum(walls) / len(walls)

        # the same calls with two tasks in flight (task queue): the copies of call k+1 -- 1.5 GiB over PCIe, points AND
        # scalars again -- run under the kernels of call k; every call still streams its inputs and reads its result
        def enqueue():
            m.initialize(params)
            m.start_process()
            m.set_data(bz.MSMInput((pts_pinned.data_ptr(), n * c.point_size), (sc_pinned.data_ptr(), n * 32), params))
        k = max(3, args.steps)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        enqueue()
        for i in range(k):
            if i + 1 < k:
                enqueue()
            m.wait_result()
            ok = ok and bool(m.result().result == exp)
        torch.cuda.synchronize()
        ms_pipe = 1e3 * (time.perf_counter() - t0) / k
        return {"workload": "BN254 MSM 2^24, DMA mode (configs[2]): points + scalars streamed from pinned host memory every call",
                "ms_per_call": ms, "scalar_mults_per_s": n / (ms / 1e3), "device_pipeline_ms": sum(dev) / len(dev),
                "pipelined_ms_per_call": ms_pipe, "pipelined_scalar_mults_per_s": n / (ms_pipe / 1e3),
                "pipelined_mode": "two calls in flight through the task queue: the H2D copies of call k+1 overlap the kernels of call k",
                "h2d_bytes_per_call": n * (c.point_size + 32), "verified_bit_exact_vs_oracle_closed_form": ok,
                "plan": m.plan_info()}
    finally:
        m.close()


def precompute_section(args, bz, torch, dc):
    """The reference's own "precomputed points" wire format at scale (integration_msm_hbm.rs:13-119,
    tests/msm/mod.rs:360-380): is_precompute = true, HBM mode, N = 2^24 bases x 8 records 2^(32 i) P = 12 GiB of wire
    bytes resident in the card's address space; the core pairs the eight 32-bit limbs of every scalar with them, i.e. an
    MSM over 2^27 points with 32-bit scalars.  The x8 records are DERIVED on the device from a generated base set
    (bz_msm_expand_precompute: the host-side precompute of the reference's test helper takes hours at this size) and
    checked against the oracle on sampled records; the MSM result is checked against the closed form."""
    import numpy as np
    from oracle import capi
    from oracle.py import curves
    from util import random_scalars, seed_points
    c = curves.BLS12_381
    log_n = args.precompute_log_n
    n = 1 << log_n
    p0, q = seed_points(c, 55)
    rec = c.point_size * 8
    base_addr, x8_addr = 0x10_0000_0000, 0x20_0000_0000
    gen = bz.MSMClient.new(bz.MSMInit(bz.PointMemoryType.HBM, False, bz.Curve.BLS381), dc)
    try:
        gen.generate_chain_points(p0 + q, 0, n, base_addr, 0)
        gen.expand_precompute(base_addr, n, x8_addr)
        # sampled records against the oracle's naive 2^(32 i) P
        ok_rec = True
        for k in (0, 1, n // 3, n - 1):
            got = gen.get_data_from_hbm(rec, x8_addr, k * rec)
            p = got[:c.point_size]
            for i in range(8):
                ok_rec &= got[i * c.point_size:(i + 1) * c.point_size] == capi.point_mul("BLS12_381", p, pow(2, 32 * i, c.r))
    finally:
        gen.close()
    sc_np = random_scalars(c, n, seed=56)
    sc_pinned = torch.empty(n * 32, dtype=torch.uint8).pin_memory()
    sc_pinned.numpy()[:] = sc_np
    m = bz.MSMClient.new(bz.MSMInit(bz.PointMemoryType.HBM, True, bz.Curve.BLS381), dc)
    try:
        params = bz.MSMParams(n, (x8_addr, 0))
        walls, dev = [], []
        res = None
        for i in range(2 + max(3, args.steps)):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            m.initialize(params)
            m.start_process()
            m.set_data(bz.MSMInput(None, (sc_pinned.data_ptr(), n * 32), params))
            m.wait_result()
            res = m.result().result
            if i >= 2:
                walls.append(time.perf_counter() - t0)
                dev.append(m.phase_times()["total"])
        exp = capi.chain_expected("BLS12_381", p0, q, sc_np, n)
        ok = bool(res == exp)
        ms = 1e3 * sum(walls) / len(walls)
        # the same calls with two tasks in flight (the copy of the next scalars and the tail of a task under its neighbour)
        k = max(3, args.steps)

        def enqueue():
            m.initialize(params)
            m.start_process()
            m.set_data(bz.MSMInput(None, (sc_pinned.data_ptr(), n * 32), params))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        enqueue()
        for i in range(k):
            if i + 1 < k:
                enqueue()
            m.wait_result()
            ok = ok and bool(m.result().result == exp)
        torch.cuda.synchronize()
        ms_pipe = 1e3 * (time.perf_counter() - t0) / k
        return {"workload": "BLS12-381 MSM 2^%d with is_precompute = true in HBM mode: %d GiB of x8 records 2^(32 i) P resident, "
                            "scalars from pinned host memory every call" % (log_n, (n * rec) >> 30),
                "ms_per_call": ms, "scalar_mults_per_s": n / (ms / 1e3), "device_pipeline_ms": sum(dev) / len(dev),
                "pipelined_ms_per_call": ms_pipe, "pipelined_scalar_mults_per_s": n / (ms_pipe / 1e3),
                "sampled_records_match_oracle": bool(ok_rec), "verified_bit_exact_vs_oracle_closed_form": ok,
                "plan": m.plan_info()}
    finally:
        m.close()


def poseidon_section(args, bz, torch, dc):
    """PoseidonClient throughput: a TreeC tree of height h (base layer: arity-11 column hashes, upper layers arity 8,
    integration_poseidon.rs:109-116) fed in bulk, all records drained; root + sampled nodes checked against the oracle."""
    import numpy as np
    from oracle.py import poseidon as P
    h = args.poseidon_height
    nbase = 8 ** (h - 1)
    total = (8 ** h - 1) // 7
    rng = np.random.default_rng(77)
    raw = rng.integers(0, 256, size=(nbase * 11, 32), dtype=np.uint8)
    raw[:, 31] &= 0x3f
    data = raw.reshape(-1)
    import ctypes
    from blaze_b200._lib import lib
    pinned_in = torch.empty(data.nbytes, dtype=torch.uint8).pin_memory()
    pinned_in.numpy()[:] = data
    out_buf = torch.empty((total + 8) * 64, dtype=torch.uint8).pin_memory()
    pc = bz.PoseidonClient.new(bz.Hash.Poseidon, dc)
    try:
        best = None
        res = None
        for it in range(3):
            pc.initialize(bz.PoseidonInitializeParameters(h, bz.TreeMode.TreeC, ""))
            torch.cuda.synchronize()
            got = ctypes.c_size_t()
            t0 = time.perf_counter()
            # the C-ABI calls a Rust / C++ caller makes: one bulk set_data, then result(expected) into the caller's buffer
            rc1 = lib().bz_poseidon_set_data(pc._h, ctypes.c_void_p(pinned_in.data_ptr()), data.nbytes)
            rc2 = lib().bz_poseidon_result(pc._h, total, ctypes.c_void_p(out_buf.data_ptr()), total + 8, ctypes.byref(got))
            dt = time.perf_counter() - t0
            assert rc1 == 0 and rc2 == 0 and got.value == total
            best = dt if best is None else min(best, dt)
        res = bz.PoseidonResult.parse_poseidon_hash_results(bytes(out_buf.numpy()[:total * 64]))   # outside the timed region
        ok = len(res) == total
        # the first base node, and the root recomputed from the returned layer below it
        elems = [int.from_bytes(bytes(data[32 * i:32 * i + 32]), "little") for i in range(11)]
        by = {(r.layer_id, r.hash_id): int.from_bytes(r.hash_byte, "little") for r in res}
        ok &= by[(0, 0)] == P.hash_elems(elems)
        ok &= by[(h - 1, 0)] == P.hash_elems([by[(h - 2, i)] for i in range(8)])
        dev_ms = pc.device_ms()
        hbm_peak, _ = peaks()
        alg = nbase * 12 * 32 + (total - nbase) * 9 * 32
        return {"workload": "Poseidon (x^5, BLS12-381 Fr, R_F = 8, R_P = 57) TreeC tree of height %d: %d arity-11 + %d arity-8 hashes" % (h, nbase, total - nbase),
                "hashes": total, "e2e_s": best, "hashes_per_s_e2e": total / best,
                "device_ms": dev_ms, "hashes_per_s_device": total / (dev_ms / 1e3) if dev_ms else None,
                "h2d_bytes": int(data.nbytes), "d2h_bytes": total * 64, "verified_vs_oracle": bool(ok),
                "roofline": {"bound": "hbm", "achieved": alg / (dev_ms / 1e3) / 1e9 if dev_ms else None, "peak": hbm_peak, "unit": "GB/s",
                             "frac": (alg / (dev_ms / 1e3) / 1e9 / hbm_peak) if dev_ms else None, "algorithmic_bytes": alg,
                             "note": "multiplier bound (~%d field products per hash), not HBM bound" % 700}}
    finally:
        pc.close()


CPU_SAMPLE_LOG_N = 22   # ONE bounded sample size for the CPU arm: --impl reference at every N and cpu_baseline at N = 1


def cpu_sample_note(log_n_sample, log_n):
    return ("2^%d-point prefix of the 2^%d workload per step (chain points P0+iQ, scalars uniform in [0, r)); arkworks' window rule "
            "c = ceil(log2 n)*69/100 + 2 gives c = %d on the sample and %d at 2^%d (about 10-15%% more adds per scalar on the sample), "
            "and its window-parallel schedule keeps at most ceil(255/c) = %d threads busy"
            % (log_n_sample, log_n, log_n_sample * 69 // 100 + 2, log_n * 69 // 100 + 2, log_n,
               (255 + log_n_sample * 69 // 100 + 1) // (log_n_sample * 69 // 100 + 2)))


def run_reference(args, rank, world):
    """CPU arm: the reference's own definition of the path (oracle port), bounded sample per step."""
    if rank != 0:
        return
    from oracle import capi
    capi.build()
    ln = min(CPU_SAMPLE_LOG_N, args.log_n) if args.cpu_sample_log_n is None else args.cpu_sample_log_n
    for _ in range(max(1, min(args.warmup, 1))):
        cpu_port_rate(min(ln, 14))
    rates, times = [], []
    th = 0
    for s in range(args.steps):
        r, th, dt = cpu_port_rate(ln, seed=900 + s)
        rates.append(r)
        times.append(dt)
    total_n = args.steps * (1 << ln)
    value = total_n / sum(times)
    sample = cpu_sample_note(ln, args.log_n)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u64 limbs (381-bit Fq Montgomery)",
        "data": "synthetic",
        "config": {"workload": "BLS12-381 MSM 2^%d, arkworks-0.3-style Pippenger restated in C++ (oracle port), "
                               "host CPU" % args.log_n, "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": th, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def bind_to_gpu_numa_node(local):
    """Pin this rank (and the pinned buffers it is about to allocate) to the CPUs next to its GPU: with eight ranks on
    one node the H2D copies otherwise all cross from whatever socket the processes happened to start on."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        pynvml.nvmlDeviceSetCpuAffinity(h)
        return sorted(os.sched_getaffinity(0))[:2] + ["..."] if len(os.sched_getaffinity(0)) > 2 else sorted(os.sched_getaffinity(0))
    except Exception as e:      # best effort
        return "unavailable: %r" % (e,)


def msm_workload(args, bz, torch, dist, dc, rank, world, cname, curve_enum, seed, steps, e2e=True, plain=False):
    """The headline workload on one curve: 2^log_n MSM, HBM-resident points P0 + iQ (generated on the device, untimed),
    uniform scalars.  N > 1: `dc` is a ranked client -- every rank owns points / scalars [g N/G, (g+1) N/G), the
    library all-gathers the projective partial records (NCCL) and sums them on its stream; result() is the full sum on
    every rank.  Returns a dict of measurements (max over ranks)."""
    import numpy as np
    from oracle.py import curves
    from util import random_scalars, seed_points
    from blaze_b200._lib import lib
    c = curves.CURVES[cname]
    N = 1 << args.log_n
    per = N // world
    first = rank * per
    m = bz.MSMClient.new(bz.MSMInit(bz.PointMemoryType.HBM, False, curve_enum), dc)
    out = {}
    try:
        p0, q = seed_points(c, seed)
        HBM_ADDR = 0
        m.generate_chain_points(p0 + q, first, per, HBM_ADDR, 0)
        params = bz.MSMParams(per, (HBM_ADDR, 0))
        full_sc = random_scalars(c, N, seed=seed + 1)
        sc_np = full_sc[first * 32:(first + per) * 32]
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

        def enqueue_host():
            m.initialize(params)
            m.start_process()
            m.set_data(bz.MSMInput(None, (sc_pinned.data_ptr(), per * 32), params))

        def barrier():
            if dist is not None:
                dist.barrier()
            torch.cuda.synchronize()

        def timed(fn, k):
            barrier()
            t0 = time.perf_counter()
            fn(k)
            barrier()
            return time.perf_counter() - t0

        def loop_resident(k):
            for _ in range(k):
                step_resident()

        # ---- plain table (no window-merged precompute): what a point set used once or twice gets
        if plain:
            m.set_precompute(0)
            step_resident()
            w = timed(loop_resident, 3)
            out["plain_table_ms"] = 1e3 * w / 3
            out["plain_plan"] = m.plan_info()
            m.set_precompute(1)
        # ---- warm-up (also builds the workspace; the second use of the resident points derives the merged table)
        res = None
        warm_s = []
        for _ in range(max(args.warmup, 3)):
            t0 = time.perf_counter()
            res = step_resident()
            warm_s.append(round(time.perf_counter() - t0, 3))
        out["warmup_step_s"] = warm_s
        out["table_build_s"] = m.table_build_ms() / 1e3
        # ---- verification at full size: closed form of the chain workload (bit-exact), on every rank
        verified = None
        if not args.no_verify:
            from oracle import capi
            capi.build()
            exp = capi.chain_expected(cname, p0, q, full_sc, N) if rank == 0 else None
            if dist is not None:
                box = [exp]
                dist.broadcast_object_list(box, src=0)
                exp = box[0]
            verified = bool(res == exp)
            flag = torch.tensor([1 if verified else 0], device="cuda")
            if dist is not None:
                dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            verified = bool(int(flag))
            if not verified:
                raise SystemExit("bench: GPU result differs from the oracle closed form -- number is INVALID")
        out["verified"] = verified
        del full_sc
        # ---- timed region 1: device-resident (value)
        dev_ms, acc_ms, sort_ms, red_ms = [], [], [], []

        def enqueue_resident():
            m.initialize(params)
            m.start_process()
            m.set_scalars_device(sc_dev.data_ptr(), params)

        def loop_value(k):
            # two tasks in flight through the client's task queue (like the e2e loop below): the latency-bound end of
            # task k (upper reduction levels, window combine, exchange, result copy -- the library's tail stream) and
            # the host's enqueue work overlap the windowing / sort / accumulation of task k+1
            enqueue_resident()
            for i in range(k):
                if i + 1 < k:
                    enqueue_resident()
                m.wait_result()
                m.result()
                pt = m.phase_times()
                dev_ms.append(pt["total"])
                acc_ms.append(pt["accumulate"])
                sort_ms.append(pt["sort"])
                red_ms.append(pt["reduce"])
        out["sampler"] = ClockSampler(range(world)) if rank == 0 else None
        launches0 = lib().bz_kernel_launch_count()
        barrier()
        if out["sampler"]:
            out["sampler"].start()
        wall = timed(loop_value, steps)
        out["clocks"] = out.pop("sampler").stop() if rank == 0 else None
        out["launches"] = lib().bz_kernel_launch_count() - launches0
        # ---- the same, strictly serial (wait for every result before the next task is queued): reported beside `value`
        out["wall_serial"] = timed(loop_resident, steps)
        wall_e2e_serial = wall_e2e = None
        if e2e:
            # ---- timed region 2: end to end with host buffers, strictly serial calls
            def loop_serial(k):
                for _ in range(k):
                    enqueue_host()
                    m.wait_result()
                    m.result()
            wall_e2e_serial = timed(loop_serial, steps)

            # ---- timed region 3: end to end with host buffers, two tasks in flight (the reference's task queue:
            # start_process/set_data of task k+1 are issued before wait_result/result of task k, so the H2D
            # copy of the next step's scalars overlaps the kernels of the current one); every step still copies
            # its scalars from pinned host memory and reads its result back inside the timed region
            def loop_pipe(k):
                enqueue_host()
                for i in range(k):
                    if i + 1 < k:
                        enqueue_host()
                    m.wait_result()
                    m.result()
            wall_e2e = timed(loop_pipe, steps)
        vals = torch.tensor([sum(dev_ms) / len(dev_ms), wall, wall_e2e or 0.0, sum(acc_ms) / len(acc_ms), wall_e2e_serial or 0.0,
                             sum(sort_ms) / len(sort_ms), sum(red_ms) / len(red_ms), out.get("plain_table_ms", 0.0), out["table_build_s"], out["wall_serial"]],
                            dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(vals, op=dist.ReduceOp.MAX)
        (out["dev_step_ms"], out["wall"], out["wall_e2e"], out["acc_ms"], out["wall_e2e_serial"], out["sort_ms"], out["red_ms"],
         out["plain_table_ms"], out["table_build_s"], out["wall_serial"]) = [float(x) for x in vals.cpu()]
        out["plan"] = m.plan_info()
        out["N"], out["per"], out["point_size"], out["result_point_size"] = N, per, c.point_size, c.result_point_size
        out["q_bits"], out["r_bits"] = c.q.bit_length(), c.r.bit_length()
    finally:
        m.close()
    return out


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import blaze_b200 as bz

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    all_cpus = os.sched_getaffinity(0)
    affinity = bind_to_gpu_numa_node(local) if world > 1 else "not bound (single GPU)"
    torch.cuda.set_device(local)
    dist = None
    dc = bz.DriverClient(str(local), bz.DriverConfig.driver_client_cfg(bz.CardType.B200))
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        # one process per GPU: the library's own communicator carries the data-path exchange (the MSM's final sum, the
        # NTT's handles and barriers); torch.distributed only hands the NCCL id around and takes the max of the timings
        uid = [bz.DriverClient.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        dc.comm_init(rank, world, uid[0])

    cname = {"BLS381": "BLS12_381", "BLS377": "BLS12_377", "BN254": "BN254"}[args.curve]
    r = msm_workload(args, bz, torch, dist, dc, rank, world, cname, getattr(bz.Curve, args.curve), 2026, args.steps,
                     e2e=True, plain=True)
    N, per = r["N"], r["per"]
    ms_per_step = 1e3 * r["wall"] / args.steps
    value = N / (ms_per_step / 1e3)
    e2e_value = N / (r["wall_e2e"] / args.steps)

    config5 = None
    if world > 1 and not args.no_config5 and args.curve == "BLS381":
        # BASELINE.json configs[4]: BLS12-377 MSM 2^26 across the GPUs of the box (point-sharded)
        try:
            r5 = msm_workload(args, bz, torch, dist, dc, rank, world, "BLS12_377", bz.Curve.BLS377, 3026, max(3, args.steps // 2),
                              e2e=False, plain=False)
            ms5 = 1e3 * r5["wall"] / max(3, args.steps // 2)
            config5 = {"workload": "BLS12-377 MSM 2^%d, HBM-resident points, point-sharded x%d (configs[4])" % (args.log_n, world),
                       "value": r5["N"] / (ms5 / 1e3), "unit": UNIT, "ms_per_step": ms5, "device_ms_per_step": r5["dev_step_ms"],
                       "serial_ms_per_step": 1e3 * r5["wall_serial"] / max(3, args.steps // 2),
                       "verified_bit_exact_vs_oracle_closed_form": r5["verified"], "plan": r5["plan"],
                       "phase_ms": {"sort": r5["sort_ms"], "accumulate": r5["acc_ms"], "reduce": r5["red_ms"]}}
        except SystemExit:
            raise
        except Exception as e:
            config5 = {"error": repr(e)}

    ntt = None
    if not args.no_ntt:
        try:
            ntt = ntt_section(args, bz, torch, dist, dc, rank, world, local)
        except SystemExit:
            raise
        except Exception as e:     # the headline line must still be printed
            ntt = {"error": repr(e)}

    dma = pre8 = pos = None
    if world == 1 and not args.no_dma:
        try:
            dma = dma_section(args, bz, torch, dc)
        except Exception as e:
            dma = {"error": repr(e)}
    if world == 1 and not args.no_precompute:
        try:
            pre8 = precompute_section(args, bz, torch, dc)
        except Exception as e:
            pre8 = {"error": repr(e)}
    if world == 1 and not args.no_poseidon:
        try:
            pos = poseidon_section(args, bz, torch, dc)
        except Exception as e:
            pos = {"error": repr(e)}

    if rank == 0:
        plan = r["plan"]
        W, cbits = plan["windows"], plan["c"]
        hbm_peak, peak_src = peaks()
        # algorithmic bytes of the accumulate sweep: (point_size + 4 B index) per (scalar, window)
        alg_bytes = (r["point_size"] + 4) * per * W
        achieved = alg_bytes / (r["acc_ms"] / 1e3) / 1e9
        traffic, traffic_src = None, None
        for name in ("r2_traffic.json", "r1_traffic.json"):
            tp = os.path.join(ROOT, "profiles", name)
            if os.path.exists(tp):
                try:
                    tj = json.load(open(tp))
                    # an ncu --set full capture of the same kernel (not taken in this run: ncu replays would void the
                    # timing); other shard sizes scale with the algorithmic bytes
                    traffic = tj["k_accumulate_dram_bytes_per_launch"] * alg_bytes / tj["algorithmic_bytes_per_launch"]
                    traffic_src = "profiles/" + name + " (ncu dram__bytes_read.sum + dram__bytes_write.sum of this kernel, scaled by algorithmic bytes)"
                    break
                except Exception:
                    traffic = None
        # the HBM-bound part of the sweep -- scalar windowing + bucket sort (k_digits, two partition levels, final level):
        # algorithmic bytes as implemented: digits 32 + 4 W per scalar; partition level 1 reads the 4-byte digit
        # entry twice (histogram, scatter) and writes an 8-byte pair; level 2 reads the pair twice and writes it; the final
        # level reads it twice and writes the 4-byte reference
        entries = per * W
        sort_bytes = per * (32 + 4 * W) + entries * ((4 + 4 + 8) + (8 + 8 + 8) + (8 + 8 + 4))
        sort_gbs = sort_bytes / (r["sort_ms"] / 1e3) / 1e9
        line = {
            "metric": METRIC.replace("BLS12-381", cname.replace("_", "-")).replace("2^26", "2^%d" % args.log_n),
            "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None,
            "dtype": "u32 limbs (%d-bit Fq Montgomery, %d-bit Fr)" % (r["q_bits"], r["r_bits"]),
            "data": "synthetic",
            "config": {
                "workload": "%s MSM 2^%d, HBM-resident points (configs[1]); points P0+iQ generated on device, "
                            "canonical scalars uniform in [0, r)" % (cname.replace("_", "-"), args.log_n),
                "precompute_factor": 1, "window_bits": cbits, "windows": W, "segment": plan["segment"],
                "bucket_sets": plan["bucket_sets"],
                "resident_table": ("window-merged: 2^(c*w)*P_i for the %d digit windows derived once from the resident "
                                   "points (%d MiB of HBM, built in warm-up, not timed), all windows share one bucket set"
                                   % (W, plan["merged_table_mib"])) if plan["merged_table"] else "points only (Montgomery form)",
                "table_build_s": r["table_build_s"],
                "plain_table_value": (N / (r["plain_table_ms"] / 1e3)) if r["plain_table_ms"] else None,
                "plain_table_ms_per_step": r["plain_table_ms"] or None,
                "plain_table_plan": {k: r["plain_plan"][k] for k in ("c", "windows", "bucket_sets")} if r.get("plain_plan") else None,
                "parallelism": ("point-sharded x%d, one process per GPU; partial results all-gathered (NCCL) and summed on "
                                "the device inside the library" % world) if world > 1 else "single GPU",
                "l2": "inputs (2 GiB scalars + 6 GiB points per 2^26) exceed the 126 MB L2; no flush needed",
                "verified_bit_exact_vs_oracle_closed_form": r["verified"],
                "phase_ms": {"sort": r["sort_ms"], "accumulate": r["acc_ms"], "reduce": r["red_ms"]},
                "device_ms_per_step": r["dev_step_ms"],
                "value_mode": "two tasks in flight through the client's task queue: the latency-bound end of step k (upper "
                              "reduction levels, window combine, exchange, result copy on the library's tail stream) overlaps "
                              "the windowing / sort / accumulation of step k+1; every step's result is read and its phase "
                              "times are recorded",
                "serial_value": N / (r["wall_serial"] / args.steps),
                "serial_ms_per_step": 1e3 * r["wall_serial"] / args.steps,
                "warmup_step_s": r["warmup_step_s"],
                "cpu_affinity": affinity,
            },
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": per * 32 * world,
                    "d2h_bytes_per_step": r["result_point_size"] * world, "ms_per_step": 1e3 * r["wall_e2e"] / args.steps,
                    "mode": "two tasks in flight through the client's task queue (H2D of step k+1 overlaps step k)",
                    "serial_value": N / (r["wall_e2e_serial"] / args.steps),
                    "serial_ms_per_step": 1e3 * r["wall_e2e_serial"] / args.steps},
            "gpu_launches": int(r["launches"]),
            "clocks": r["clocks"],
            "roofline": {"bound": "hbm", "kernel": "bucket accumulation (k_accumulate*<%s>)" % cname, "achieved": achieved, "peak": hbm_peak,
                         "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": traffic, "traffic_source": traffic_src,
                         "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes,
                         "note": "the sweep is integer-multiplier bound, not HBM bound (ncu: profiles/r2_ncu_accumulate*.txt); "
                                 "its HBM fraction is small by construction"},
        }
        line["roofline_bucket_sort"] = {
            "bound": "hbm", "kernels": "k_digits + k_part_hist/scan/scatter x2 + k_final (windowing and bucket sort of the sweep)",
            "achieved": sort_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": sort_gbs / hbm_peak, "algorithmic_bytes": sort_bytes,
            "ms": r["sort_ms"], "note": "the memory-bound phase of the sweep (5 % of the step); the EC-addition phase above is multiplier-bound"}
        if config5 is not None:
            line["config5"] = config5
        if ntt is not None:
            line["ntt"] = ntt
        if dma is not None:
            line["dma_mode"] = dma
        if pre8 is not None:
            line["precompute_x8"] = pre8
        if pos is not None:
            line["poseidon"] = pos
        if world == 1 and not args.no_cpu_baseline:
            from oracle import capi
            capi.build()
            os.sched_setaffinity(0, all_cpus)     # the CPU arm gets every host core the process started with
            ln = min(CPU_SAMPLE_LOG_N, args.log_n) if args.cpu_sample_log_n is None else args.cpu_sample_log_n
            cpu_port_rate(14)
            rr, th, dt = cpu_port_rate(ln)
            line["cpu_baseline"] = {"value": rr, "unit": UNIT, "cores": th, "kind": "port",
                                    "sample": "one step, %.1f s: " % dt + cpu_sample_note(ln, args.log_n)}
        print(json.dumps(line), flush=True)
    dc.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
