#!/usr/bin/env python
"""bench.py -- throughput of the batched Go1 MPC QP hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            (own arm; torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K --warmup W   (CPU arm, host cores)

A "step" is one pass of the Go1 MPC hot path over one batch of synthetic robots: per robot one
step-location/step-timing SQP tick (3 QP solves, n=4 p=1 m=24, plus write-back, LIPM roll-out
and step indices) and one body-inclination MPC tick (condensation -> Goldfarb-Idnani QP ->
clamp -> roll-out; 1 QP solve, n=2nh m=12nh) -- two kernels on two streams.  `value` counts QP
solves (one solve = one solve_quadprog call of the reference).
Workload at every N: BASELINE.json configs[1] -- Go1 MPC, batch 4096 randomised states and
velocity commands per GPU, horizon 10 (seed 0xB2000002 + rank); weak
scaling: every rank owns its own 4096 robots, no collective inside the step.

Timing: CUDA events on the stream the kernels are launched on, W untimed steps, then exactly K
timed steps bracketed by barrier + synchronize; max over ranks.  The steps rotate through
enough distinct input batches that the input footprint exceeds 2x L2 (126 MB), so no step
re-reads a cache-resident batch.  The `e2e` leg calls the pipelined host-buffer C-ABI entries (pinned
host memory, H2D + kernels + D2H inside the timed region); what the reference classes keep as members
(planner state; body step table and previous body results) stays on the device, what their tick methods
take as arguments moves every step (`--e2e-records full` uploads whole body records instead).  It also
reports the host-to-host latency of one lone batch (`e2e.latency_ms`).
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
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "batched MPC QP solves/sec"
UNIT = "solves/s"
L2_BYTES = 126 * 1024 * 1024


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=4096, help="instances per GPU per step")
    ap.add_argument("--nh", type=int, default=10, help="body-MPC horizon")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-records", choices=("resident", "full"), default="resident",
                    help="body records of the e2e leg: 'resident' keeps tx and the previous output record on the device "
                         "(go1mpc_body_mpc_step_batch_resident_host_async), 'full' uploads whole records every tick")
    ap.add_argument("--streams", type=int, default=20, help="CUDA streams the device-resident leg deals its steps over")
    ap.add_argument("--body-streams", type=int, default=8, help="of those, streams the body-MPC ticks are dealt over")
    ap.add_argument("--sweep", action="store_true", help="also print per-batch-size throughput (stderr)")
    return ap.parse_args()


def workload_config(a, n_gpus):
    return {"workload": f"cfg2: Go1 MPC, batch {a.batch} robots per GPU with randomised CoM / body-angle states and velocity "
                        f"commands; per robot and step: one step-location/step-timing SQP tick (3 QP solves, n=4 p=1 m=24) + "
                        f"one body-inclination MPC tick at horizon {a.nh} (1 QP solve, n={2 * a.nh} m={12 * a.nh})",
            "batch_per_gpu": a.batch, "global_batch": a.batch * n_gpus, "horizon": a.nh, "sqp_iterations": 3,
            "qp_shapes": [[2 * a.nh, 0, 12 * a.nh], [4, 1, 24]], "seed": "0xB2000002+rank",
            "parallelism": f"batch-sharded x{n_gpus}, no collective",
            "l2": "inputs rotate through distinct batches totalling > 2x L2 (no flush needed)",
            "streams": f"device-resident leg: body ticks round-robin over {max(1, a.body_streams)} CUDA streams, planner ticks over "
                       f"{max(max(1, a.body_streams) + 1, a.streams) - max(1, a.body_streams)} (independent robot batches in flight)"}


# ----------------------------------------------------------------------------- CPU arm
def cpu_runner(a, threads):
    """Returns (run_once, solves): run_once() runs one step (step-timing tick + body tick for a.batch
    robots) with `threads` host threads through the oracle port (-O3 -march=native build), returns seconds."""
    from tests import oracle_lib
    from quadrupedal_loco_b200 import synth
    orc = oracle_lib.Oracle(fast=True)
    nh, B = a.nh, a.batch
    d = synth.body_mpc_inputs(B, nh, seed=synth.SEED_CFG2)
    cfg = orc.body_cfg(nh)
    scfg = orc.step_cfg(3)
    tick, st0, sin = synth.step_timing_inputs(B, orc.step_default_state(scfg), seed=synth.SEED_CFG2)
    bounds = np.linspace(0, B, threads + 1).astype(int)
    nsolved = np.zeros(threads, np.int64)

    def work(t, lo, hi):
        theta = d["theta"][lo:hi].copy(); x = d["x_warm"][lo:hi].copy(); o14 = np.zeros((hi - lo, 14))
        orc.body_step_batch(cfg, d["tick"][lo:hi], np.ascontiguousarray(d["tx"][lo:hi]), theta,
                            np.ascontiguousarray(d["bstate"][lo:hi]), np.ascontiguousarray(d["refs"][lo:hi]), o14, x)
        st = st0[lo:hi].copy(); out = np.zeros((hi - lo, 38))
        orc.lib.orc_step_timing_batch(ctypes.byref(scfg), int(hi - lo), oracle_lib.PI(np.ascontiguousarray(tick[lo:hi])),
                                      oracle_lib.P(st), oracle_lib.P(np.ascontiguousarray(sin[lo:hi])), oracle_lib.P(out), None)

    import ctypes
    # solves per step: body 1 per robot + the SQP solves that actually run (the planner skips them near a step's end)
    _, dg = orc.step_tick_batch(scfg, tick[:512], st0[:512].copy(), sin[:512])
    solves = B + int(round(dg[:, 4].mean() * B))

    def run_once():
        t0 = time.perf_counter()
        if threads == 1:
            work(0, 0, B)
        else:
            ts = [threading.Thread(target=work, args=(i, bounds[i], bounds[i + 1])) for i in range(threads)]
            for t in ts:
                t.start()
            for t in ts:
                t.join()
        return time.perf_counter() - t0
    return run_once, solves


def cpu_baseline(a, budget_s=10.0):
    cores = os.cpu_count() or 1
    out = {}
    for label, thr in (("1thread", 1), ("all", cores)):
        run, solves = cpu_runner(a, thr)
        run()
        n, el = 0, 0.0
        while el < budget_s / 2 and n < 200:
            el += run(); n += 1
        out[label] = (solves * n / el, n)
    return {"value": out["all"][0], "unit": UNIT, "cores": cores, "kind": "port",
            "value_1thread": out["1thread"][0],
            "sample": f"{out['all'][1]} passes over the same {a.batch}-robot cfg2 batch with {cores} host threads "
                      f"({out['1thread'][1]} passes single-threaded); oracle/ C restatement of NLPClass::step_timing_opti_loop, "
                      f"PRMPCClass::body_theta_mpc and EiQuadProg at -O3 -march=native (the reference needs Eigen, absent here, "
                      f"and compiles the body MPC only at horizon 4)"}


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    run, solves = cpu_runner(a, cores)
    steps = min(a.steps, 100)
    for _ in range(min(a.warmup, 5)):
        run()
    times = [run() for _ in range(steps)]
    tot = sum(times)
    v = solves * steps / tot
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": steps,
            "warmup": min(a.warmup, 5), "ms_per_step": 1e3 * tot / steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(a, 1),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"each step = one pass over the {a.batch}-robot cfg2 batch split over {cores} host threads "
                                       f"(oracle/ C port at -O3 -march=native; host CPU only, GPU count does not apply)"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """Polls NVML (SM clock, power, clock-event reasons) every few ms from a thread while the
    timed region runs; only samples taken between start() and stop() are kept."""
    REASONS = {8: "hw_slowdown", 64: "hw_thermal_slowdown", 32: "sw_thermal_slowdown", 4: "sw_power_cap"}

    def __init__(self, index):
        self.index, self.samples, self.run, self.t, self.h = index, [], False, None, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.h = None

    def _poll(self):
        nv = self.nv
        while self.run:
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                pw = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                self.samples.append((sm, pw, rs))
            except Exception:
                pass
            time.sleep(0.004)

    def start(self):
        if self.h is None:
            return
        self.run = True
        self.t = threading.Thread(target=self._poll, daemon=True)
        self.t.start()

    def stop(self):
        if self.h is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "note": "NVML unavailable"}
        self.run = False
        self.t.join()
        sm = [x[0] for x in self.samples]
        reasons = set()
        for _, _, rs in self.samples:
            for bit, nm in self.REASONS.items():
                if rs & bit:
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(self.max_sm),
                "power_w_max": max((x[1] for x in self.samples), default=None), "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------------------- GPU arm
# DRAM bytes (read + write) of the kernels of one three-launch body-MPC call, per batch size, from ncu --set full
TRAFFIC_TRI = {4096: 32007936, 65536: 230868736}   # profiles/r01_tri_kernels.md (4096: 11.8 - 32.0 MB, depends on what L2 still holds)


def run_b200(a):
    import torch
    import quadrupedal_loco_b200 as q
    from quadrupedal_loco_b200 import synth

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        # stdout carries exactly one JSON line: keep NCCL's own banner / warnings on stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        os.environ["NCCL_DEBUG"] = os.environ.get("NCCL_DEBUG", "WARN") if os.environ.get("NCCL_DEBUG", "").upper() not in ("VERSION", "INFO", "TRACE") else "WARN"
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    nh, B, K, W = a.nh, a.batch, a.steps, max(a.warmup, 3)
    mpc = q.Go1Mpc(local)
    stream = torch.cuda.ExternalStream(mpc.stream, device=dev)
    side = torch.cuda.Stream(device=dev)          # the step-timing tick runs beside the body tick
    in_s, out_s, dg_s = q.body_in_stride(nh), q.body_out_stride(nh), q.body_diag_stride(nh)

    # distinct input batches: footprint > 2x L2
    per_batch = B * (in_s + q.STEP_STATE + q.STEP_IN) * 8
    nrot = min(512, max(4, int(np.ceil(2.2 * L2_BYTES / per_batch))))
    big = synth.body_mpc_inputs(B * nrot, nh, seed=synth.SEED_CFG2 + rank)
    rec = q.pack_body_inputs(nh, big["tick"], big["tx"], big["theta"], big["bstate"], big["x_warm"], big["refs"])
    rec_h = torch.from_numpy(rec).pin_memory()
    in_d = rec_h.to(dev).view(nrot, B, in_s)
    out_d = torch.zeros(nrot, B, out_s, dtype=torch.float64, device=dev)
    diag_d = torch.zeros(nrot, B, dg_s, dtype=torch.int32, device=dev)
    # step-timing side: SoA [field][B] per rotation slot; state is double-buffered (out-of-place) so
    # every pass sees the same inputs
    stick, sst, sinp = synth.step_timing_inputs(B * nrot, mpc.step_default_state(), seed=synth.SEED_CFG2 + rank)
    soa = lambda x, f: np.ascontiguousarray(x.reshape(nrot, B, f).transpose(0, 2, 1))
    st_h = torch.from_numpy(soa(sst, q.STEP_STATE)).pin_memory(); si_h = torch.from_numpy(soa(sinp, q.STEP_IN)).pin_memory()
    tk_h = torch.from_numpy(stick.reshape(nrot, B).copy()).pin_memory()
    st_d, si_d, tk_d = st_h.to(dev), si_h.to(dev), tk_h.to(dev)
    # the tick updates the planner state in place; to keep every pass on the same inputs each step first
    # refreshes its slot's working copy with a device-to-device copy (6.6 MB, inside the timed region)
    st_o = torch.zeros(nrot, B * q.STEP_STATE, dtype=torch.float64, device=dev)
    so_d = torch.zeros(nrot, q.STEP_OUT, B, dtype=torch.float64, device=dev)
    sd_d = torch.zeros(nrot, q.STEP_DIAG, B, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    ev_fork = torch.cuda.Event(); ev_join = torch.cuda.Event()
    # raw device addresses per rotation slot: the launch loop is then two ctypes calls per step
    lib, hh = mpc.lib, mpc.h
    P_in = [in_d[r].data_ptr() for r in range(nrot)]; P_out = [out_d[r].data_ptr() for r in range(nrot)]
    P_dg = [diag_d[r].data_ptr() for r in range(nrot)]
    P_tk = [tk_d[r].data_ptr() for r in range(nrot)]; P_st = [st_d[r].data_ptr() for r in range(nrot)]
    P_si = [si_d[r].data_ptr() for r in range(nrot)]; P_so = [so_d[r].data_ptr() for r in range(nrot)]
    P_sd = [sd_d[r].data_ptr() for r in range(nrot)]; P_sto = [st_o[r].data_ptr() for r in range(nrot)]
    side_ptr = side.cuda_stream

    def launch_body(i, st_ptr=None):
        r = i % nrot
        rc = lib.go1mpc_body_mpc_step_batch(hh, nh, B, P_in[r], P_out[r], P_dg[r], st_ptr)
        assert rc == 0, rc

    st_bytes = B * q.STEP_STATE * 8

    def launch_sqp(i, st_ptr=None):
        r = i % nrot
        rc = lib.go1mpc_copy_device_async(hh, P_sto[r], P_st[r], st_bytes, st_ptr)
        assert rc == 0, rc
        rc = lib.go1mpc_step_timing_step_batch(hh, 3, B, P_tk[r], P_sto[r], P_sto[r], P_si[r], P_so[r], P_sd[r], st_ptr)
        assert rc == 0, rc

    def step(i):
        """one robot batch through both ticks: the SQP tick on the side stream beside the body tick"""
        ev_fork.record(stream)
        side.wait_event(ev_fork)
        launch_sqp(i, side_ptr)
        launch_body(i)
        ev_join.record(side)
        stream.wait_event(ev_join)

    def barrier():
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    # warm-up (also fills every diag record once: algorithmic flops and solve counts per batch)
    with torch.cuda.stream(stream):
        for i in range(max(W, nrot)):
            step(i)
    torch.cuda.synchronize()
    diag_all = diag_d.cpu().numpy()
    assert (diag_all[:, :, 0] == 0).all(), "a warm-up body-MPC solve did not converge"
    flops_per_batch = diag_all[:, :, 9].astype(np.float64).sum(axis=1)          # [nrot]
    mean_iters = diag_all[:, :, 2:6].reshape(-1, 4).mean(axis=0)
    mean_l2a = diag_all[:, :, 8].mean()
    sdiag = sd_d.cpu().numpy()
    sqp_solves = sdiag[:, 4, :].astype(np.int64).sum(axis=1)                     # [nrot] QPs the SQP really solved
    sqp_status = sdiag[:, 5::11, :][:, :3, :]
    # algorithmic flops of the planner tick (SURVEY.md 8d closed form for the (4,1,24) QP: 0.15 k + 0.35 k per
    # outer pass; front-end 0.6 k per SQP iteration)
    sqp_outer = sdiag[:, 7::11, :][:, :3, :].astype(np.float64)
    sqp_flops_per_batch = (150.0 * (sqp_status >= 0) + 350.0 * sqp_outer * (sqp_status >= 0) + 600.0).sum(axis=(1, 2))
    solves_per_step = B + float(np.mean(sqp_solves))

    dfma_gflops = mpc.measure_dfma_peak(300)

    # ---- device-resident leg ----
    clocks = ClockSampler(local)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
    barrier()
    # the two ticks of a robot are independent and so are the steps (distinct robot batches): fork once,
    # deal the steps round-robin over NSTREAM streams (step i: SQP tick then body tick on stream i mod
    # NSTREAM), join once; the timed region is fork -> join.  One 4096-robot batch fills 1/64 of the
    # GPU's warp slots in the SQP kernel and 1.7 waves in the body kernel, so batches in flight on
    # several streams are what keeps the SMs busy at this batch size.
    # body ticks go round-robin over NB streams (one call is three dependent launches -- setup, solve, merge --
    # that each fill only part of the GPU at this batch size: calls in flight on several streams overlap them),
    # planner ticks over the remaining a.streams - NB (each launch is 128 warps for ~0.19 ms: several must be
    # in flight to hide that latency)
    NB = max(1, a.body_streams)
    NSTREAM = max(NB + 1, a.streams)
    lanes = [stream] + [torch.cuda.Stream(device=dev) for _ in range(NSTREAM - 1)]
    lane_ptr = [x.cuda_stream for x in lanes]
    joins = [torch.cuda.Event() for _ in range(NSTREAM)]
    mpc_b = q.Go1Mpc(local)          # the body feeder thread's handle
    hb = mpc_b.h
    # first use of a stream allocates the body path's per-stream workspace: keep that out of the timed region
    for k in range(NB):
        launch_body(k, lane_ptr[k])
        rc = lib.go1mpc_body_mpc_step_batch(hb, nh, B, P_in[k % nrot], P_out[k % nrot], P_dg[k % nrot], lane_ptr[k])
        assert rc == 0, rc
    for k in range(NB, NSTREAM):
        launch_sqp(k, lane_ptr[k])
    torch.cuda.synchronize()
    barrier()
    clocks.start()
    l0 = mpc.launch_count; l0b = mpc_b.launch_count
    # two host threads enqueue (ctypes releases the GIL inside the C calls): one deals the planner ticks, the other --
    # through a second handle, handles being thread-compatible, not thread-safe -- the body ticks (4 launches per
    # call).  One thread alone needs ~40 us per step for the 6 launches + 1 copy, close to what the GPU needs.
    def feed_planner():
        torch.cuda.set_device(local)         # CUDA's current device is per host thread
        for i in range(K):
            launch_sqp(i, lane_ptr[NB + i % (NSTREAM - NB)])

    def feed_body():
        torch.cuda.set_device(local)
        for i in range(K):
            r = i % nrot
            rc = lib.go1mpc_body_mpc_step_batch(hb, nh, B, P_in[r], P_out[r], P_dg[r], lane_ptr[i % NB])
            assert rc == 0, rc
    with torch.cuda.stream(stream):
        ev[0].record(stream)
        for x in lanes[1:]:
            x.wait_event(ev[0])
        t_host = time.perf_counter()
        ta = threading.Thread(target=feed_planner); tb = threading.Thread(target=feed_body)
        ta.start(); tb.start(); ta.join(); tb.join()
        t_host = time.perf_counter() - t_host
        for k in range(1, NSTREAM):
            joins[k].record(lanes[k])
            stream.wait_event(joins[k])
        ev[K].record(stream)
    torch.cuda.synchronize()
    barrier()
    launches = (mpc.launch_count - l0) + (mpc_b.launch_count - l0b)
    total_ms = ev[0].elapsed_time(ev[K])
    solves_timed = float(sum(B + sqp_solves[i % nrot] for i in range(K)))

    # single-batch latency of the whole step (both kernels, one batch in flight) and the duration of
    # each kernel alone: the same launches, a sync between them so the events bracket exactly one
    k_lat = min(K, 400)

    def timed(fn):
        out = []
        for i in range(k_lat):
            with torch.cuda.stream(stream):
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                e0.record(stream); fn(i); e1.record(stream)
            e1.synchronize()
            out.append(e0.elapsed_time(e1))
        return np.array(out)
    lat_ms = timed(step)
    body_ms = timed(launch_body)
    sqp_ms = timed(launch_sqp)

    # each tick alone under the timed region's stream layout: k_ov calls dealt over the same streams, one
    # fork, one join -- the average duration of a call when independent batches are in flight, which is what
    # the roofline of the body tick is computed from (a lone call leaves most of the GPU idle at this batch size)
    k_ov = min(K, 600)

    def overlapped(fn, first, count):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            for x in lanes[1:]:
                x.wait_event(e0)
            for i in range(k_ov):
                fn(i, lane_ptr[first + i % count])
            for k in range(1, NSTREAM):
                joins[k].record(lanes[k])
                stream.wait_event(joins[k])
            e1.record(stream)
        e1.synchronize()
        return e0.elapsed_time(e1) / k_ov
    body_ov_ms = overlapped(launch_body, 0, NB)
    sqp_ov_ms = overlapped(launch_sqp, NB, NSTREAM - NB)
    clk = clocks.stop()
    handed_over = mpc.body_handover_total() + mpc_b.body_handover_total(); guard_trips = mpc.body_guard_trips() + mpc_b.body_guard_trips()
    mpc_b.close()
    body_mode = os.environ.get("GO1MPC_BODY_MODE", "auto")

    t = torch.tensor([total_ms, solves_timed], dtype=torch.float64, device=dev)
    if dist:
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        total_ms_max, solves_all = float(tmax[0].item()), float(tsum[1].item())
    else:
        total_ms_max, solves_all = total_ms, solves_timed
    value = solves_all / (total_ms_max * 1e-3)

    # ---- e2e leg: host buffers through the pipelined *_host_async C-ABI entries ----
    # per step: H2D of that step's inputs (body records; planner tick + sensor inputs) from pinned host
    # memory, both kernels, D2H of the results (body out + diag; planner out38 + diag).  The planner
    # STATE is resident on the device -- that is what the ABI offers a real caller.
    e2e = None
    if not a.no_e2e:
        Ke = min(K, 400)
        in_np = rec_h.view(nrot, B, in_s).numpy()
        out_np = torch.zeros(nrot, B, out_s, dtype=torch.float64).pin_memory().numpy()
        diag_np = torch.zeros(nrot, B, dg_s, dtype=torch.int32).pin_memory().numpy()
        so_np = torch.zeros(nrot, q.STEP_OUT, B, dtype=torch.float64).pin_memory().numpy()
        sd_np = torch.zeros(nrot, q.STEP_DIAG, B, dtype=torch.int32).pin_memory().numpy()
        si_np, tk_np = si_h.numpy(), tk_h.numpy()
        H_in = [in_np[r].ctypes.data for r in range(nrot)]; H_out = [out_np[r].ctypes.data for r in range(nrot)]
        H_dg = [diag_np[r].ctypes.data for r in range(nrot)]; H_tk = [tk_np[r].ctypes.data for r in range(nrot)]
        H_si = [si_np[r].ctypes.data for r in range(nrot)]; H_so = [so_np[r].ctypes.data for r in range(nrot)]
        H_sd = [sd_np[r].ctypes.data for r in range(nrot)]
        resident = a.e2e_records == "resident"
        if resident:
            # what stays in HBM between ticks: tx [B][28] and the output records (warm start x, stale out14)
            tk_s = q.body_tick_in_stride(nh)
            tx_np, xw_np, tick_np = q.split_body_record(nh, in_np.reshape(nrot * B, in_s))
            R_tx = torch.from_numpy(tx_np).to(dev).view(nrot, B, 28)
            R_out = torch.zeros(nrot, B, out_s, dtype=torch.float64, device=dev)
            R_out[:, :, 18:18 + 2 * nh] = torch.from_numpy(xw_np).to(dev).view(nrot, B, 2 * nh)
            ti_h = torch.from_numpy(tick_np).view(nrot, B, tk_s).pin_memory()
            to_h = torch.zeros(nrot, B, q.BODY_TICK_OUT, dtype=torch.float64).pin_memory()
            ti_np, to_np = ti_h.numpy(), to_h.numpy()
            P_rtx = [R_tx[r].data_ptr() for r in range(nrot)]; P_rout = [R_out[r].data_ptr() for r in range(nrot)]
            H_ti = [ti_np[r].ctypes.data for r in range(nrot)]; H_to = [to_np[r].ctypes.data for r in range(nrot)]
            torch.cuda.synchronize()

        # one handle per host thread (handles are thread-compatible, not thread-safe): thread A feeds the
        # planner ticks, thread B the body ticks; ctypes releases the GIL inside the C calls, so the two
        # enqueue loops (about 40 us of driver calls per tick each) run side by side
        mpc2 = q.Go1Mpc(local)
        lib2, hh2 = mpc2.lib, mpc2.h

        def feed_sqp(n):
            for i in range(n):
                r = i % nrot
                rc = lib.go1mpc_step_timing_step_batch_host_async(hh, 3, B, H_tk[r], P_st[r], P_sto[r], H_si[r], H_so[r], H_sd[r])   # out-of-place: the inputs stay put
                assert rc == 0, rc
                if r == nrot - 1:
                    mpc.synchronize()        # a host slot is about to be reused
            mpc.synchronize()

        def feed_body(n):
            for i in range(n):
                r = i % nrot
                if resident:
                    rc = lib2.go1mpc_body_mpc_step_batch_resident_host_async(hh2, nh, B, P_rtx[r], P_rout[r], H_ti[r], H_to[r], H_dg[r])
                else:
                    rc = lib2.go1mpc_body_mpc_step_batch_host_async(hh2, nh, B, H_in[r], H_out[r], H_dg[r])
                assert rc == 0, rc
                if r == nrot - 1:
                    mpc2.synchronize()
            mpc2.synchronize()

        def run_e2e(n):
            ta = threading.Thread(target=feed_sqp, args=(n,)); tb = threading.Thread(target=feed_body, args=(n,))
            ta.start(); tb.start(); ta.join(); tb.join()
        run_e2e(8)
        barrier()
        l1 = mpc.launch_count + mpc2.launch_count
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        t_wall = time.perf_counter()
        run_e2e(Ke)
        t_wall = time.perf_counter() - t_wall
        e1.record(stream)
        e1.synchronize()
        barrier()
        e2e_launches = int(mpc.launch_count + mpc2.launch_count - l1)
        # host-to-host latency of ONE batch through the same entries: enqueue both ticks, wait for both handles
        # (wall clock on the calling thread: H2D + kernels + D2H + the driver calls; nothing else in flight)
        lat_e2e = []
        for i in range(8 + 100):
            r = i % nrot
            t0 = time.perf_counter()
            feed_sqp_one = lib.go1mpc_step_timing_step_batch_host_async(hh, 3, B, H_tk[r], P_st[r], P_sto[r], H_si[r], H_so[r], H_sd[r])
            if resident:
                rc = lib2.go1mpc_body_mpc_step_batch_resident_host_async(hh2, nh, B, P_rtx[r], P_rout[r], H_ti[r], H_to[r], H_dg[r])
            else:
                rc = lib2.go1mpc_body_mpc_step_batch_host_async(hh2, nh, B, H_in[r], H_out[r], H_dg[r])
            assert rc == 0 and feed_sqp_one == 0
            mpc.synchronize(); mpc2.synchronize()
            if i >= 8:
                lat_e2e.append(1e3 * (time.perf_counter() - t0))
        lat_t = torch.tensor([np.percentile(lat_e2e, 50), np.percentile(lat_e2e, 99), max(lat_e2e)], dtype=torch.float64, device=dev)
        if dist:
            dist.all_reduce(lat_t, op=dist.ReduceOp.MAX)
        lat_e2e_ms = {"p50": float(lat_t[0].item()), "p99": float(lat_t[1].item()), "max": float(lat_t[2].item()),
                      "what": "one %d-robot batch host to host through the same two entries (enqueue both ticks, synchronize both handles), "
                              "wall clock on the calling thread, 100 samples, max over ranks" % B}
        mpc2.close()
        se = float(sum(B + sqp_solves[i % nrot] for i in range(Ke)))
        te = torch.tensor([e0.elapsed_time(e1), se], dtype=torch.float64, device=dev)
        if dist:
            tm = te.clone(); dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            tsu = te.clone(); dist.all_reduce(tsu, op=dist.ReduceOp.SUM)
            te_ms, se_all = float(tm[0].item()), float(tsu[1].item())
        else:
            te_ms, se_all = float(te[0].item()), se
        assert (diag_np[:min(Ke, nrot), :, 0] == 0).all()
        assert np.array_equal(diag_np[0], diag_all[0]) and np.array_equal(sd_np[0], sdiag[0]), "e2e results differ from the device-resident leg"
        if resident:
            want = out_d[0].cpu().numpy()
            assert np.array_equal(to_np[0][:, :18], want[:, :18]) and np.array_equal(to_np[0][:, 18], want[:, 18 + 2 * nh]), \
                "e2e body results differ from the device-resident leg"
        body_up, body_down = (tk_s, q.BODY_TICK_OUT) if resident else (in_s, out_s)
        e2e = {"value": se_all / (te_ms * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": B * (body_up * 8 + q.STEP_IN * 8 + 4),
               "d2h_bytes_per_step": B * (body_down * 8 + dg_s * 4 + q.STEP_OUT * 8 + q.STEP_DIAG * 4),
               "body_records": a.e2e_records, "latency_ms": lat_e2e_ms,
               "steps": Ke, "ms_per_step": te_ms / Ke, "wall_ms_per_step": 1e3 * t_wall / Ke, "launches": e2e_launches,
               "api": "go1mpc_step_timing_step_batch_host_async + " + ("go1mpc_body_mpc_step_batch_resident_host_async" if resident else "go1mpc_body_mpc_step_batch_host_async") + " (pinned host buffers; "
                      "H2D, kernel, D2H per call on 8 internal lanes; planner state" + (", body step table and previous body output record" if resident else "") + " resident on the device; two handles fed by two "
                      "host threads; go1mpc_synchronize before a host slot is reused and at the end; timed with CUDA events "
                      "recorded before the first enqueue and after the last synchronize)"}

    if rank == 0:
        kern_ms = float(body_ov_ms)
        kern_alone_ms = float(np.mean(body_ms))
        fl = float(np.mean(flops_per_batch[np.arange(k_lat) % nrot]))
        achieved_tf = fl / (kern_ms * 1e-3) / 1e12
        peak_tf = dfma_gflops / 1e3
        io_bytes = B * ((in_s + out_s) * 8 + dg_s * 4)
        try:
            hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]; hbm_src = "measured"
        except Exception:
            hbm_peak = 6650.0; hbm_src = "fallback"
        tri = body_mode in ("auto", "tri") and nh in (4, 10) and (body_mode == "tri" or B >= 2048)
        body_kernel = ("body-inclination MPC tick = tri_setup_kernel + tri_solve_kernel + tri_merge_kernel (+ the list-mode "
                       "body_fast_kernel launch, empty on this workload)") if tri else "body_fast_kernel (body-inclination MPC tick)"
        # DRAM bytes of one call's kernels, ncu --set full at this batch size (profiles/r01_summary.md)
        traffic = (TRAFFIC_TRI.get(B) if nh == 10 else None) if tri else ({4096: 4855552}.get(B) if nh == 10 else None)
        sqp_fl = float(np.mean(sqp_flops_per_batch))
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": max(W, nrot),
            "ms_per_step": total_ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(a, world),
            "solves_per_step_per_gpu": solves_per_step, "robot_ticks_per_s": world * B * K / (total_ms_max * 1e-3),
            "latency_ms": {"p50": float(np.percentile(lat_ms, 50)), "p99": float(np.percentile(lat_ms, 99)),
                           "max": float(lat_ms.max()),
                           "what": f"one {B}-robot batch through both ticks (planner launch beside the body launches, 2 streams), "
                                   f"CUDA events, {k_lat} samples"},
            "kernels_ms": {"body_tick_alone": kern_alone_ms, "body_tick_overlapped": kern_ms,
                           "step_timing_kernel_alone": float(np.mean(sqp_ms)), "step_timing_kernel_overlapped": float(sqp_ov_ms),
                           "step_both_alone": float(np.mean(lat_ms))},
            "host_enqueue_ms_per_step": 1e3 * t_host / K,
            "body_path": {"mode": body_mode, "three_launch": bool(tri), "handed_to_combined_kernel": int(handed_over),
                          "guard_trips": int(guard_trips)},
            "roofline": {"bound": "fp64", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                         "frac": achieved_tf / peak_tf if peak_tf else None,
                         "traffic": traffic,
                         "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum summed over the kernels of one body-MPC call, "
                                           "ncu --set full (profiles/r01_summary.md); algorithmic I/O is bytes_per_launch below, the rest "
                                           "is the setup -> solve -> merge hand-over (J per instance, state / result records per half)",
                         "kernel": body_kernel, "kernel_ms": kern_ms, "kernel_ms_alone": kern_alone_ms,
                         "kernel_ms_def": f"average duration of a call with independent {B}-robot batches in flight on {NB} streams "
                                          f"({k_ov} calls between one fork and one join, CUDA events on the launching stream), i.e. under "
                                          "the timed region's schedule; kernel_ms_alone is a lone call (most of the GPU idle at this batch size)",
                         "frac_alone": fl / (kern_alone_ms * 1e-3) / 1e12 / peak_tf if peak_tf else None,
                         "flops_per_launch": fl, "flops_per_solve": fl / B,
                         "flops_def": "algorithmic flops of the dense reference algorithm along each problem's path "
                                      "(SURVEY.md 8d formula, n = 20, m = 120, counted per problem on the device); the kernels exploit "
                                      "G = blockdiag(H, H) and execute fewer",
                         "peak_source": "measured live on this GPU: register-resident DFMA loop (go1mpc_measure_dfma_peak); "
                                        "MEASURED_PEAKS.json has no FP64 figure",
                         "hbm": {"achieved": io_bytes / (kern_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                 "peak_source": f"{hbm_src} (MEASURED_PEAKS.json)", "bytes_per_launch": io_bytes},
                         "other_kernels": {"step_timing_kernel": {
                             "kernel_ms": float(sqp_ov_ms), "kernel_ms_alone": float(np.mean(sqp_ms)), "flops_per_launch": sqp_fl,
                             "achieved": sqp_fl / (float(sqp_ov_ms) * 1e-3) / 1e12,
                             "frac": sqp_fl / (float(sqp_ov_ms) * 1e-3) / 1e12 / peak_tf if peak_tf else None,
                             "note": "thread-per-planner scalar kernel (3 QPs of 4 variables + front-end per tick): local-memory and "
                                     "latency bound (profiles/r01_summary.md); the body tick holds 94 % of the step's algorithmic flops, "
                                     "this kernel most of its time -- the next kernel to restructure"}}},
            "solver": {"body_mean_outer": float(mean_iters[0]), "body_mean_add": float(mean_iters[1]),
                       "body_mean_drop": float(mean_iters[2]), "body_mean_degen": float(mean_iters[3]),
                       "body_mean_l2a": float(mean_l2a), "sqp_solves_per_robot": float(np.mean(sqp_solves)) / B,
                       "sqp_converged_frac": float((sqp_status == 0).mean()), "sqp_infeasible_frac": float((sqp_status == 2).mean())},
            "clocks": clk, "gpu_launches": int(launches), "e2e": e2e,
        }
        if world == 1 and not a.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(a)
        print(json.dumps(line), flush=True)

    if a.sweep and rank == 0:
        for Bs in (256, 1024, 4096, 16384, 65536):
            d = synth.body_mpc_inputs(Bs, nh, seed=1)
            r = torch.from_numpy(q.pack_body_inputs(nh, d["tick"], d["tx"], d["theta"], d["bstate"], d["x_warm"], d["refs"])).to(dev)
            o = torch.zeros(Bs, out_s, dtype=torch.float64, device=dev)
            tk, ss, ii = synth.step_timing_inputs(Bs, mpc.step_default_state(), seed=1)
            tkd = torch.from_numpy(tk).to(dev); ssd = torch.from_numpy(np.ascontiguousarray(ss.T)).to(dev)
            iid = torch.from_numpy(np.ascontiguousarray(ii.T)).to(dev); sso = torch.zeros_like(ssd)
            ood = torch.zeros(q.STEP_OUT, Bs, dtype=torch.float64, device=dev)
            torch.cuda.synchronize()
            res = []
            for fn in (lambda: mpc.body_mpc_step(nh, Bs, r, o, None),
                       lambda: (mpc.lib.go1mpc_copy_device_async(mpc.h, sso.data_ptr(), ssd.data_ptr(), ssd.numel() * 8, None), mpc.step_timing_step(3, Bs, tkd, sso, iid, ood, None))):
                for _ in range(3):
                    fn()
                mpc.synchronize()
                with torch.cuda.stream(stream):
                    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                    e0.record(stream)
                    for _ in range(20):
                        fn()
                    e1.record(stream)
                e1.synchronize()
                res.append(e0.elapsed_time(e1) / 20)
            print(f"sweep B={Bs}: body {res[0] * 1e3:.1f} us ({Bs / res[0] * 1e3:.3e} solves/s), "
                  f"step-timing SQP {res[1] * 1e3:.1f} us ({3 * Bs / res[1] * 1e3:.3e} solves/s)", file=sys.stderr)
    mpc.close()
    if dist:
        dist.destroy_process_group()


def main():
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)


if __name__ == "__main__":
    main()
