#!/usr/bin/env python3
"""bench.py -- PLONK proofs/sec on the BASELINE.json workload (2^20-constraint BN254 circuit).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--log2 20] [--curve BN254]

One "step" = one proof (b2p_prove: 9 MSMs of ~n points, 6 NTTs of size 4n + 5 of size n, quotient, openings)
of the synthetic squaring-chain circuit of SURVEY 8d on a known-tau SRS.  N > 1 runs one replica per GPU
(proofs are independent: no data-path collective, "scaling": "weak"); timing is CUDA events on the library's
stream, max over ranks.  Rank 0 prints ONE JSON line.  At N > 1 the line also carries "msm_sharded": one
2^log2-point MSM with the point set split over the N GPUs (one all_gather of a point per rank over NCCL,
DESIGN.md section 7), timed the same way.

  value      proofs/s with L, R, O already resident in HBM (b2p_prove_dev)
  e2e        proofs/s through the reference-facing C-ABI call b2p_prove with pinned HOST buffers
             (H2D of 3*32*n bytes and D2H of the proof inside the timed region)
  roofline   dominant kernel (k_msm_accumulate): algorithmic bytes n*(64+32) per launch
             / CUDA-event launch duration, against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline  the C++ CPU oracle (oracle/cpu_plonk.cpp, OpenMP) on a bounded sample, rank 0, N = 1

--impl reference times that CPU oracle alone (the reference is pure Go over un-vendored gnark and no Go
toolchain exists here, so oracle/_ref cannot be built: kind "port").
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "plonk_proofs_per_sec"
UNIT = "proofs/s"
TAU = 0x1234567890ABCDEF1234567890ABCDEF1234567890ABCDEF
# dram__bytes_read.sum + dram__bytes_write.sum of ONE k_msm_accumulate launch, from `ncu --set full` captures
# (profiles/ncu_r2_accum_summary.csv; round 1's capture read the same), keyed by (curve, log2 rows): the gather reads W = 13
# table points per scalar, so ~13x the algorithmic bytes.  A configuration nobody captured reports null and says so.
ACCUM_DRAM_TRAFFIC = {("BN254", 20): 1_873_200_000,        # 1.8033 GB read + 69.9 MB written
                      ("BLS12_381", 20): 2_724_300_000}    # 2.6286 GB read + 95.7 MB written
# The roof that binds the accumulation: the SM's multiplier pipe.  IMAD issues at 1.84e13 lanes/s on 148 SMs
# (tools/microbench.cu, profiles/microbench_r2.json); a 32x32->64 product takes two such slots in whichever form ptxas
# emits it (IMAD.WIDE.U32.X, measured at half the IMAD rate, or IMAD + IMAD.HI).  One XYZZ mixed addition executes
# 6 products + 2 squarings + one a*b - c*d = 2 326 slots on BN254 (8 limbs: 256 / 203 / 384 per operation), 5 214 on
# BLS12-381 (12 limbs: 576 / 447 / 864) -- SASS-counted, profiles/sass_counts_r2.json.
IMAD_SLOTS_PER_S = 1.84e13
MADD_SLOTS = {"BN254": 2326, "BLS12_381": 5214}
MADD_ROOFLINE = {c: IMAD_SLOTS_PER_S / s for c, s in MADD_SLOTS.items()}

# ---------------------------------------------------------------------------------------
# pieces shared with tests/test_bench_host.py
# ---------------------------------------------------------------------------------------
def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def reduce_over_ranks(local_ms: float, local_units: float, world: int, device=None):
    """Whole-job figures: elapsed = max over ranks, units = sum over ranks (replicas, no data collective)."""
    if world == 1:
        return local_ms, local_units
    import torch
    import torch.distributed as dist
    t = torch.tensor([local_ms], dtype=torch.float64, device=device)
    u = torch.tensor([local_units], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(u, op=dist.ReduceOp.SUM)
    return float(t.item()), float(u.item())


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def msm_algorithmic_bytes(n_scalars: int, curve: str) -> int:
    """MSM(n) moves n * (affine point + scalar) bytes."""
    return n_scalars * ((64 if curve == "BN254" else 96) + 32)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
            # nvidia-smi's own start-up (NVML initialisation) holds driver locks for a few hundred ms and would land in
            # the first timed pass: wait until it delivers its first sample
            t0 = time.perf_counter()
            while not self.lines and time.perf_counter() - t0 < 3.0 and self.proc.poll() is None:
                time.sleep(0.02)
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        time.sleep(0.05)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------
# workload
# ---------------------------------------------------------------------------------------
def build_workload(curve: str, log2_rows: int):
    """Synthetic squaring-chain circuit with exactly 2^log2_rows rows."""
    from algoplonk_b200 import frontend as fe
    cs, values = fe.squaring_chain(curve, log2_rows)
    tc = fe.build_trace(cs)
    L, R, O = fe.solve_lro(cs, values, tc.n)
    return cs, tc, L, R, O


# ---------------------------------------------------------------------------------------
# CPU oracle timing (cpu_baseline leg and --impl reference)
# ---------------------------------------------------------------------------------------
def host_cores() -> int:
    """The cores this process may run on (cgroup / affinity aware), NOT what OMP_NUM_THREADS says:
    torch.distributed.run exports OMP_NUM_THREADS=1 to every rank, which is right for ranks that drive a GPU
    and wrong for the CPU arm, whose whole point is to use the box's host cores."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


class CpuProver:
    """The CPU arm: oracle/cpu_plonk.cpp (kind "port": gnark itself needs Go, absent here) proving the SAME
    full-size workload as the GPU arm, on all host cores.  Nothing is ever extrapolated from a smaller circuit."""

    def __init__(self, curve: str, log2_rows: int):
        from oracle import cpu_oracle as co
        self.co = co
        self.cores = host_cores()
        co.set_threads(self.cores)
        cid = co.CURVE_ID[curve]
        cs, tc, L, R, O = build_workload(curve, log2_rows)
        srs = co.srs_from_tau_bytes(cid, TAU, tc.n + 3)
        self.circ = co.Circuit(cid, tc.n, tc.nb_public, tc.ql, tc.qr, tc.qm, tc.qo, tc.qk, tc.perm, (), (), srs)
        self.cols = [co.scalars_le(c) for c in (L, R, O)]
        self.bl = co.scalars_le(range(1, 10))
        self.threads = co.threads()

    def prove_seconds(self) -> float:
        t0 = time.perf_counter()
        self.circ.prove(*self.cols, self.bl)
        return time.perf_counter() - t0

    def free(self):
        self.circ.free()


def cpu_baseline(curve: str, target_log2: int, budget_s: float = 30.0):
    """cpu_baseline leg of the GPU arm's line (rank 0, N = 1): full-size proofs on all host cores, as many as
    fit ~budget_s (at least one)."""
    cp = CpuProver(curve, target_log2)
    times = [cp.prove_seconds()]
    while sum(times) + times[-1] <= budget_s:
        times.append(cp.prove_seconds())
    cp.free()
    t = min(times)
    return {"value": 1.0 / t, "unit": UNIT, "cores": cp.threads, "kind": "port",
            "sample": f"{len(times)} full 2^{target_log2}-row {curve} proof(s) (the bench workload itself, nothing "
                      f"scaled), best {t:.2f} s, OpenMP over {cp.threads} host threads"}


REFERENCE_BUDGET_S = 240.0      # wall-clock bound of the timed + warm-up proofs of --impl reference


def plan_reference_steps(step_s: float, steps: int, warmup: int, budget_s: float = REFERENCE_BUDGET_S):
    """(warm-up proofs, timed proofs) the CPU arm actually runs: the requested counts when they fit the
    budget, else fewer -- reported as run, never scaled up."""
    if (steps + warmup) * step_s <= budget_s:
        return warmup, steps
    w = 1 if warmup else 0              # the first proof (page faults, thread start-up) is the warm-up that matters
    k = int((budget_s - w * step_s) // step_s)
    return w, max(1, min(steps, k))


def run_reference(args):
    rank, _, world = dist_env()
    if rank != 0:
        return
    cp = CpuProver(args.curve, args.log2)
    first = cp.prove_seconds()                   # full-size proof: calibrates the plan and is the first warm-up
    w, k = plan_reference_steps(first, args.steps, args.warmup)
    for _ in range(max(0, w - 1)):
        cp.prove_seconds()
    if w == 0:
        times = [first] + [cp.prove_seconds() for _ in range(k - 1)]
    else:
        times = [cp.prove_seconds() for _ in range(k)]
    cp.free()
    per_step = sum(times) / len(times)
    value = 1.0 / per_step
    sample = (f"each step = one FULL 2^{args.log2}-row {args.curve} proof (the GPU arm's workload, nothing scaled) on "
              f"{cp.threads} host threads; {k} timed + {w} warm-up proofs run"
              + ("" if (w, k) == (args.warmup, args.steps) else
                 f" (requested {args.steps} + {args.warmup}: reduced to fit {REFERENCE_BUDGET_S:.0f} s of CPU time)"))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": k, "warmup": w, "steps_requested": args.steps, "warmup_requested": args.warmup,
        "ms_per_step": per_step * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cp.threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "step_seconds": {"min": min(times), "max": max(times)},
        "note": "CPU restatement of gnark's prover (oracle/cpu_plonk.cpp, OpenMP); gnark itself needs Go, absent here. "
                "One CPU prover on the whole host at every --gpus N (the GPU arm's value is the N-replica aggregate).",
    }
    print(json.dumps(line), flush=True)


def workload_config(args):
    return {"workload": f"synthetic squaring-chain circuit, 2^{args.log2} constraints, {args.curve}, "
                        f"known-tau SRS of 2^{args.log2}+3 points, k=0",
            "log2_constraints": args.log2, "curve": args.curve, "parallelism": f"replicas x{args.gpus}",
            "inflight_per_gpu": max(1, args.inflight),
            "l2": "inputs larger than L2: per-proof working set (13-window SRS table + 4n evaluations) "
                  "is > 1 GB vs 126 MB L2"}


# ---------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    from algoplonk_b200 import _lib, api
    rank, local_rank, world = dist_env()
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=device)
    _lib.init(local_rank)          # raises without a usable GPU: there is no CPU fallback
    lib = _lib.load()
    curve = args.curve
    F = max(1, args.inflight)

    cs, tc, L, R, O = build_workload(curve, args.log2)
    n = tc.n
    setup = api.SetupName.TestOnlyBN254 if curve == "BN254" else api.SetupName.TestOnlyBLS12381
    t0 = time.perf_counter()
    # one proving key (SRS table + circuit + workspace + stream) per proof in flight
    ccs = [api.Compile(cs, curve, setup) for _ in range(F)]
    load_s = (time.perf_counter() - t0) / F
    c_bits, windows, buckets = ccs[0].srs.msm_params()

    # pinned host buffers (what the cgo shim would pass) and their device-resident copies
    def pinned(data: bytes):
        return torch.frombuffer(bytearray(data), dtype=torch.uint8).pin_memory()
    hL, hR, hO = (pinned(api.fr_to_mont_bytes(curve, col)) for col in (L, R, O))
    dL, dR, dO = (t.to(device) for t in (hL, hR, hO))
    blinding = C.create_string_buffer(api.fr_to_mont_bytes(curve, list(range(1, 10))))
    cid = api.CURVE_ID[curve]
    outs = [C.create_string_buffer(lib.b2p_proof_raw_size(cid, 0)) for _ in range(F)]
    streams = [torch.cuda.ExternalStream(lib.b2p_circuit_stream(cc.handle), device=device) for cc in ccs]
    torch.cuda.synchronize()

    def prove_dev(i):
        _lib.check(lib.b2p_prove_dev(ccs[i].handle, dL.data_ptr(), dR.data_ptr(), dO.data_ptr(), None, None,
                                     blinding, outs[i]))

    def prove_host(i):
        _lib.check(lib.b2p_prove(ccs[i].handle, hL.data_ptr(), hR.data_ptr(), hO.data_ptr(), None, None,
                                 blinding, outs[i]))

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    # one persistent host thread per lane (a prover service keeps its workers alive; the first CUDA call of a
    # fresh thread costs ~100 ms of one-time initialisation, which the warm-up absorbs)
    from concurrent.futures import ThreadPoolExecutor
    pool = ThreadPoolExecutor(max_workers=F)

    def timed(fn, steps, lanes):
        """EXACTLY `steps` proofs, spread over `lanes` host threads (one proving key and stream each);
        CUDA events on the library's streams, recorded by the workers: first start to last end."""
        counts = [steps // lanes + (1 if i < steps % lanes else 0) for i in range(lanes)]
        barrier()
        starts = [torch.cuda.Event(enable_timing=True) for _ in range(lanes)]
        ends = [torch.cuda.Event(enable_timing=True) for _ in range(lanes)]
        gate = threading.Barrier(lanes)

        def work(i):
            gate.wait()
            starts[i].record(streams[i])
            tw = time.perf_counter()
            for _ in range(counts[i]):
                fn(i)
            ends[i].record(streams[i])
            if os.environ.get("B2P_BENCH_DEBUG"):
                print(f"[lane {i}] {counts[i]} x {fn.__name__}: {(time.perf_counter() - tw) * 1e3:.1f} ms",
                      file=sys.stderr)

        for f in [pool.submit(work, i) for i in range(lanes)]:
            f.result()                   # re-raises a worker's exception
        for e in ends:
            e.synchronize()
        barrier()
        if os.environ.get("B2P_BENCH_DEBUG"):
            print("[events]", [[round(s.elapsed_time(e), 1) for e in ends] for s in starts], file=sys.stderr)
        return max(s.elapsed_time(e) for s in starts for e in ends)

    def warm(i):
        for _ in range(args.warmup):
            prove_dev(i)
    for f in [pool.submit(warm, i) for i in range(F)]:
        f.result()
    sampler = ClockSampler(local_rank)
    sampler.start()
    # pass 1: one proof at a time, CUDA-event spans around MSM / accumulate / NTT / quotient (no extra syncs)
    ccs[0].set_profiling(True)
    for _ in range(2):                 # the wait for the sampler idled the GPU: back to boost clocks before timing
        prove_dev(0)
    ms_single = timed(prove_dev, args.steps, 1)
    stats = ccs[0].stats()         # spans of the last timed proof
    ccs[0].set_profiling(False)
    proof_resident = bytes(outs[0].raw)
    # pass 2 (value): F proofs in flight
    launches0 = lib.b2p_launch_count()
    ms = timed(prove_dev, args.steps, F)
    launches = lib.b2p_launch_count() - launches0
    # pass 3 (e2e): the same through b2p_prove with host buffers
    for i in range(F):
        prove_host(i)
    ms_e2e = timed(prove_host, args.steps, F)
    clocks = sampler.stop()
    for o in outs:
        assert bytes(o.raw) == proof_resident, "host-buffer and resident-buffer proofs differ"
    # pass 4 (boundary cost, reported next to e2e): the columns as a Go caller holds them -- PAGEABLE memory.
    #   pageable: b2p_prove straight on pageable buffers (the driver stages them; no overlap with compute)
    #   shim:     what go/gpuplonk does per proof: memcpy of the three columns into the proving key's own
    #             page-locked columns (allocated once per key), then b2p_prove on those
    pageable = [C.create_string_buffer(bytes(t.numpy().tobytes()), t.numel()) for t in (hL, hR, hO)]
    lane_pinned = [[torch.empty_like(hL).pin_memory() for _ in range(3)] for _ in range(F)]

    def prove_pageable(i):
        _lib.check(lib.b2p_prove(ccs[i].handle, pageable[0], pageable[1], pageable[2], None, None, blinding, outs[i]))

    def prove_shim(i):
        for dst, src in zip(lane_pinned[i], pageable):
            C.memmove(dst.data_ptr(), src, dst.numel())
        _lib.check(lib.b2p_prove(ccs[i].handle, lane_pinned[i][0].data_ptr(), lane_pinned[i][1].data_ptr(),
                                 lane_pinned[i][2].data_ptr(), None, None, blinding, outs[i]))
    boundary = {}
    for name, fn in (("pageable", prove_pageable), ("shim", prove_shim)):
        for i in range(F):
            fn(i)
        ms_b = timed(fn, args.steps, F)
        tot_b, _ = reduce_over_ranks(ms_b, args.steps, world, device)
        boundary[name] = tot_b
        for o in outs:
            assert bytes(o.raw) == proof_resident, f"{name}: proof differs"

    # pass 5 ("end-to-end incl. solve"): the caller hands over the circuit's INPUTS only; the library's solver
    # (b2p_solver_solve_dev, placement by its own cost model: a dependency chain like this workload is solved on a host
    # thread, a wide circuit on the GPU) fills L, R, O in HBM and b2p_prove_dev proves them.  One solver per lane.
    from_inputs = None
    if getattr(cs, "input_vars", None) is not None and not cs.commitments and not args.no_solver_leg:
        solvers, setup_error = [], None
        try:
            import numpy as np
            from algoplonk_b200 import frontend as fe
            cols_b = [C.create_string_buffer(api.fr_to_mont_bytes(curve, c)) for c in (tc.ql, tc.qr, tc.qm, tc.qo, tc.qk)]
            wires = [np.asarray(w, dtype=np.uint32) for w in fe.solver_wires(cs, n)]
            ids = np.asarray(cs.input_vars, dtype=np.uint32)
            values_in = C.create_string_buffer(api.fr_to_mont_bytes(curve, [L[0], L[tc.nb_public]]))
            assert list(cs.input_vars) == [0, 1], "the squaring chain assigns y and x0"
            # two solver handles per lane: a handle's L, R, O stay valid until its next solve, so while proof k is
            # proved from one handle's columns a helper thread solves proof k + 1 on the other (what a prover service
            # does with a sequential solver: the chain's 27 ms per proof hide behind the GPU work)
            for _ in range(2 * F):
                h = C.c_void_p()
                _lib.check(lib.b2p_solver_create(cid, n, tc.nb_public, cs.nb_variables, ids.ctypes.data, len(ids), *cols_b,
                                                 *[w.ctypes.data for w in wires], C.byref(h)))
                solvers.append(h)
            helpers = [ThreadPoolExecutor(max_workers=1) for _ in range(F)]

            def solve(i, k):
                ptrs = [C.c_void_p() for _ in range(3)]
                _lib.check(lib.b2p_solver_solve_dev(solvers[2 * i + (k & 1)], values_in, _lib.SOLVE_AUTO,
                                                    *[C.byref(p) for p in ptrs]))
                return ptrs
            turn = [0] * F
            pending = [helpers[i].submit(solve, i, 0) for i in range(F)]

            def prove_from_inputs(i):
                ptrs = pending[i].result()
                turn[i] += 1
                pending[i] = helpers[i].submit(solve, i, turn[i])
                _lib.check(lib.b2p_prove_dev(ccs[i].handle, ptrs[0], ptrs[1], ptrs[2], None, None, blinding, outs[i]))
            for i in range(F):
                prove_from_inputs(i)
        except Exception as e:  # noqa: BLE001 -- reported in the line
            setup_error = f"{type(e).__name__}: {e}"[:300]
        # the timed pass holds barriers: every rank runs it, or none does
        ready = 0 if setup_error else 1
        if world > 1:
            import torch.distributed as dist
            flag = torch.tensor([ready], device=device, dtype=torch.int32)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            ready = int(flag.item())
        if ready:
            ms_in = timed(prove_from_inputs, args.steps, F)
            tot_in, units_in = reduce_over_ranks(ms_in, args.steps, world, device)
            same = all(bytes(o.raw) == proof_resident for o in outs)
            info = (C.c_uint64 * 8)()
            lib.b2p_solver_info(solvers[0], info)
            for f in pending:
                f.result()               # the one solve per lane that ran ahead
            from_inputs = {"value": units_in / (tot_in / 1e3), "unit": UNIT, "ms_per_step": tot_in / args.steps,
                           "same_proof_bytes": same, "pipelining": "per lane: proof k is proved while k + 1 is solved",
                           "solver": {"levels": int(info[0]), "widest_level": int(info[1]),
                                      "ran_on": "device" if info[7] == _lib.SOLVE_DEVICE else "host thread",
                                      "last_solve_ms": info[6] / 1e3},
                           "what": "circuit inputs (64 bytes) -> b2p_solver_solve_dev -> b2p_prove_dev: solving included"}
        else:
            from_inputs = {"error": setup_error or "another rank could not set the solver up"}
        if setup_error:
            for f in locals().get("pending", []):
                try:
                    f.result()
                except Exception:  # noqa: BLE001
                    pass
        for h in solvers:
            lib.b2p_solver_free(h)

    sharded_line = measure_sharded_msm(args, rank, world, device) if world > 1 else None
    # one proof over all the GPUs (commitments sharded over the point set, native peer-memory path)
    proof_sharded_line = None
    if world > 1 and not args.no_proof_sharded:
        proof_sharded_line = measure_sharded_proof(args, rank, world, device, cs, (hL, hR, hO), blinding,
                                                   proof_resident, ms_single / args.steps)

    tot_ms, units = reduce_over_ranks(ms, args.steps, world, device)
    tot_ms_e2e, _ = reduce_over_ranks(ms_e2e, args.steps, world, device)
    tot_ms_single, _ = reduce_over_ranks(ms_single, args.steps, world, device)
    # last, after every number of the line above is final: the domain-sharded NTT (both exchange modes)
    ntt_sharded_line = None
    if world in (2, 4, 8) and not args.no_ntt_sharded:
        ntt_sharded_line = measure_sharded_ntt(args, rank, world, device)
    if rank != 0:
        return
    value = units / (tot_ms / 1e3)
    e2e_value = units / (tot_ms_e2e / 1e3)

    hbm_peak, peak_src = peaks()
    msm_calls = int(stats["msm_calls"])
    accum_ms = stats["msm_accum_ms"] / msm_calls
    alg_bytes = msm_algorithmic_bytes(n + 2, curve)
    achieved = alg_bytes / (accum_ms * 1e-3) / 1e9
    traffic = ACCUM_DRAM_TRAFFIC.get((curve, args.log2))
    roofline = {"bound": "hbm", "kernel": "k_msm_accumulate", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                "frac": achieved / hbm_peak, "traffic": traffic,
                "traffic_source": ("ncu --set full: dram__bytes_read.sum + dram__bytes_write.sum of one launch "
                                   "(profiles/ncu_r2_accum_summary.csv; round 1 read the same: ncu_r1_full_summary.csv)" if traffic else
                                   f"null: no ncu --set full capture exists for ({curve}, 2^{args.log2})"),
                "peak_source": peak_src, "launch_ms": accum_ms, "algorithmic_bytes_per_launch": alg_bytes,
                "measured_in": "the one-proof-at-a-time pass (CUDA events around every launch of the kernel)",
                "note": "254/381-bit modular arithmetic makes this kernel multiplier-bound, not HBM-bound (ncu: "
                        "sm__pipe_fmaheavy_cycles_active 85 %, DESIGN.md section 4); binding_roof is the roof that binds"}
    # G1 mixed additions / s against the rate at which the multiplier pipe saturates (constants above)
    adds_per_sec = stats["msm_accum_adds"] / (stats["msm_accum_ms"] * 1e-3)
    if MADD_ROOFLINE.get(curve):
        roofline["binding_roof"] = {"bound": "int32 multiplier pipe (fmaheavy)", "achieved": adds_per_sec,
                                    "peak": MADD_ROOFLINE[curve], "unit": "G1 mixed additions/s",
                                    "frac": adds_per_sec / MADD_ROOFLINE[curve],
                                    "peak_source": f"measured IMAD issue rate {IMAD_SLOTS_PER_S:.3g} slots/s (profiles/"
                                                   f"microbench_r2.json) / {MADD_SLOTS[curve]} slots per mixed addition "
                                                   "(SASS-counted, profiles/sass_counts_r2.json); ncu reports the same "
                                                   "fraction as sm__pipe_fmaheavy_cycles_active"}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": tot_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32", "data": "synthetic", "config": workload_config(args),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 3 * 32 * n + 9 * 32,
                "d2h_bytes_per_step": int(lib.b2p_proof_raw_size(cid, 0)), "ms_per_step": tot_ms_e2e / args.steps},
        "e2e_pageable": {"value": units / (boundary["pageable"] / 1e3), "ms_per_step": boundary["pageable"] / args.steps,
                         "unit": UNIT, "what": "b2p_prove on PAGEABLE host columns (a Go slice handed over as it is)"},
        "e2e_shim": {"value": units / (boundary["shim"] / 1e3), "ms_per_step": boundary["shim"] / args.steps, "unit": UNIT,
                     "what": "go/gpuplonk's per-proof path: memcpy of the pageable columns into the key's page-locked "
                             "columns (allocated once per key), then b2p_prove"},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "clocks": clocks,
        "single_stream": {"value": units / (tot_ms_single / 1e3), "ms_per_proof": tot_ms_single / args.steps,
                          "note": "one proof at a time on one stream (latency); `value` keeps "
                                  f"{F} proofs in flight on {F} streams"},
        "msm": {"c": c_bits, "windows": windows, "buckets": buckets, "calls_per_proof": msm_calls,
                "accum_adds_per_proof": stats["msm_accum_adds"],
                "msm_g1_adds_per_sec": stats["msm_accum_adds"] / (stats["msm_accum_ms"] * 1e-3),
                "madd_roofline_per_sec": MADD_ROOFLINE.get(curve),
                "msm_ms_per_proof": stats["msm_ms"]},
        "phases_ms": {"msm": stats["msm_ms"], "msm_accumulate": stats["msm_accum_ms"], "ntt": stats["ntt_ms"],
                      "quotient": stats["quotient_ms"], "total_host_wall": stats["total_ms"]},
        "circuit_load_s": load_s,
    }
    # plonk.Verify (algoplonk.go:93) on a proof of the timed region: outside every timed region, host arithmetic
    # of the library (b2p_verify); reported so the line says its proofs are valid, never raised
    try:
        blob = api.MarshalProof(api.Proof(curve, 0, proof_resident))
        pub = api.MarshalPublicInputs(curve, L[: tc.nb_public])
        ccs[0].VerifyProof(blob, pub)
        t0 = time.perf_counter()
        for _ in range(5):
            ccs[0].VerifyProof(blob, pub)
        one_ms = (time.perf_counter() - t0) / 5 * 1e3
        t0 = time.perf_counter()
        ccs[0].VerifyProofs([blob] * 16, [pub] * 16)
        batch16_ms = (time.perf_counter() - t0) / 16 * 1e3
        ccs[0].VerifyProofs([blob] * 8, [pub] * 8, device=True)            # module load, stream
        t0 = time.perf_counter()
        ccs[0].VerifyProofs([blob] * 1024, [pub] * 1024, device=True)
        dev_us = (time.perf_counter() - t0) / 1024 * 1e6
        line["verify"] = {"accepted": True, "ms_per_proof": one_ms,
                          "ms_per_proof_batch_of_16": batch16_ms,
                          "us_per_proof_device_batch_of_1024": dev_us,
                          "where": "b2p_verify / b2p_verify_batch: host threads; b2p_verify_batch_dev: point "
                                   "combinations on the GPU, transcript on host threads; none of it is part of value / "
                                   "e2e (the reference arm times plonk.Prove only)"}
    except Exception as e:  # noqa: BLE001 -- reported in the line
        line["verify"] = {"accepted": False, "error": f"{type(e).__name__}: {e}"[:300]}
    if from_inputs is not None:
        line["e2e_from_inputs"] = from_inputs
    if sharded_line is not None:
        line["msm_sharded"] = sharded_line
    if proof_sharded_line is not None:
        line["proof_sharded"] = proof_sharded_line
    if ntt_sharded_line is not None:
        line["ntt_sharded"] = ntt_sharded_line
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(curve, args.log2)
    else:
        line["cpu_baseline"] = None
    print(json.dumps(line), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


def measure_sharded_msm(args, rank, world, device, iters: int = 10):
    """One MSM with the point set sharded over the ranks, native path (algoplonk_b200/shard_group.py,
    b2p_shard_group_msm): rank 0 holds the n scalars in HBM, every other rank's Pippenger kernels read their slice of them
    over NVLink, the partial sums come back as peer stores and are added on rank 0's device.  Sizes: the workload's
    2^log2 points and 4x that (the judge's 2^22 at the default).  CUDA events on rank 0's stream around `iters`
    blocking calls (a call ends with the D2H of the sum), against the same MSM on rank 0 alone with the whole SRS.
    Every failure is reported in the line instead of raised."""
    import torch
    import torch.distributed as dist
    from algoplonk_b200 import _lib, api, shard_group as sg
    lib = _lib.load()
    curve = args.curve
    out = {"collective": "none: peer loads of 32 n/G bytes of scalars per rank, one XYZZ point per rank stored into "
                         "rank 0's mailbox, flags in peer memory (csrc/shard_group.cuh)", "sizes": []}
    for log2 in (args.log2, args.log2 + 2):
        n = 1 << log2
        res = {"points": n}
        grp = None
        try:
            grp = sg.ShardGroup(curve, n, device=device)
            grp.connect()
        except Exception as e:  # noqa: BLE001
            res["error"] = f"setup: {type(e).__name__}: {e}"[:300]
        ok = torch.tensor([0 if grp is not None and "error" not in res else 1], dtype=torch.int32, device=device)
        dist.all_reduce(ok)
        if int(ok.item()) != 0:
            res.setdefault("error", "setup failed on another rank")
            if grp is not None:
                try:
                    grp.free()
                except Exception:  # noqa: BLE001
                    pass
            out["sizes"].append(res)
            continue
        if rank != 0:
            try:
                grp.serve()
            finally:
                grp.free()
            continue
        whole = None
        try:
            gen = torch.Generator(device="cpu").manual_seed(0xB200 + log2)
            raw = torch.randint(-(1 << 31), 1 << 31, (n, 8), generator=gen, dtype=torch.int64).to(torch.int32)
            raw[:, 7] &= 0x0FFFFFFF                      # < 2^252 < r: valid field elements (Montgomery form of something)
            d_scalars = raw.to(device).contiguous()
            stream = torch.cuda.ExternalStream(lib.b2p_srs_stream(grp.shard.handle), device=device)
            grp.announce_msm(n, iters + 3)
            first = grp.msm_dev_raw(d_scalars.data_ptr(), n, announce=False)
            for _ in range(2):
                grp.msm_dev_raw(d_scalars.data_ptr(), n, announce=False)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(iters):
                got = grp.msm_dev_raw(d_scalars.data_ptr(), n, announce=False)
            e1.record(stream)
            e1.synchronize()
            ms = e0.elapsed_time(e1) / iters
            # the same MSM on this GPU alone
            whole = api.SRS.unsafe(curve, n)
            o = C.create_string_buffer(2 * api.FP_BYTES[curve])
            ws = torch.cuda.ExternalStream(lib.b2p_srs_stream(whole.handle), device=device)
            for _ in range(3):
                _lib.check(lib.b2p_msm_g1_dev(whole.handle, _lib.BASIS_CANONICAL, d_scalars.data_ptr(), n, o))
            e0.record(ws)
            for _ in range(iters):
                _lib.check(lib.b2p_msm_g1_dev(whole.handle, _lib.BASIS_CANONICAL, d_scalars.data_ptr(), n, o))
            e1.record(ws)
            e1.synchronize()
            ms1 = e0.elapsed_time(e1) / iters
            c_bits, windows, _ = api.SRS(curve, grp.shard.handle).msm_params()
            res.update({"points_per_gpu": grp.shard.count, "ms_per_msm": ms, "one_gpu_ms_per_msm": ms1,
                        "speedup": ms1 / ms, "efficiency": ms1 / ms / world, "equal_to_one_gpu_result": got == o.raw == first,
                        "g1_adds_per_sec": n * windows / (ms * 1e-3), "shard_c": c_bits, "shard_windows": windows})
        except Exception as e:  # noqa: BLE001
            res["error"] = f"{type(e).__name__}: {e}"[:300]
        finally:
            try:
                grp.stop()
                grp.free()
                if whole is not None:
                    whole.free()
            except Exception as e:  # noqa: BLE001
                res["close_error"] = f"{type(e).__name__}: {e}"[:200]
        out["sizes"].append(res)
    return out if rank == 0 else None


def measure_sharded_proof(args, rank, world, device, cs, host_cols, blinding, want_raw: bytes, ms_one_gpu: float,
                          shard_ntt: bool = True):
    """ONE proof at a time with its 9 commitments spread over all the GPUs (algoplonk_b200/shard_group.py,
    csrc/shard_group.cuh: peer loads of the scalars, peer stores of the partial sums, flags; BASELINE configs[2]).
    Rank 0 proves through the reference-facing b2p_prove with pinned host columns, the other ranks serve.  CUDA
    events on rank 0's proving stream (a proof ends with its D2H on that stream), and the host wall clock around the
    same blocking calls.  The proof bytes must equal the single-GPU proof of the timed region above.
    Every failure is reported in the line instead of raised."""
    import torch
    import torch.distributed as dist
    from algoplonk_b200 import _lib, api, shard_group as sg
    lib = _lib.load()
    curve = args.curve
    setup = api.SetupName.TestOnlyBN254 if curve == "BN254" else api.SetupName.TestOnlyBLS12381
    res, sp = None, None
    old_c = os.environ.get("B2P_MSM_C")
    try:
        if args.shard_c:
            os.environ["B2P_MSM_C"] = str(args.shard_c)
        sp = sg.ShardedProver(cs, curve, setup, shard_ntt=shard_ntt)
    except Exception as e:  # noqa: BLE001
        res = {"error": f"setup: {type(e).__name__}: {e}"[:300]}
    finally:
        if old_c is None:
            os.environ.pop("B2P_MSM_C", None)
        else:
            os.environ["B2P_MSM_C"] = old_c
    ok = torch.tensor([0 if sp is not None else 1], dtype=torch.int32, device=device)
    dist.all_reduce(ok)
    if int(ok.item()) != 0:
        if sp is not None:
            try:
                sp.grp.free()
            except Exception:  # noqa: BLE001
                pass
        return res or {"error": "setup failed on another rank"}
    if rank != 0:
        try:
            sp.serve()
        finally:
            sp.close()
        return None
    try:
        hL, hR, hO = host_cols
        cid = api.CURVE_ID[curve]
        out = C.create_string_buffer(lib.b2p_proof_raw_size(cid, 0))
        stream = torch.cuda.ExternalStream(lib.b2p_circuit_stream(sp.cc.handle), device=device)

        def one():
            sp.grp.announce(sp.n)
            _lib.check(lib.b2p_prove(sp.cc.handle, hL.data_ptr(), hR.data_ptr(), hO.data_ptr(), None, None,
                                     blinding, out))
        for _ in range(max(2, args.warmup)):
            one()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            one()
        wall_ms = (time.perf_counter() - t0) * 1e3 / args.steps
        e1.record(stream)
        e1.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        c_bits, windows, buckets = api.SRS(curve, sp.grp.shard.handle).msm_params()
        res = {"ms_per_proof": ms, "proofs_per_sec": 1e3 / ms, "host_wall_ms_per_proof": wall_ms,
               "byte_identical_to_one_gpu_proof": bytes(out.raw) == want_raw,
               "one_gpu_ms_per_proof": ms_one_gpu, "speedup_vs_one_gpu": ms_one_gpu / ms, "n_gpus": world,
               "points_per_gpu": sp.grp.shard.count, "shard_c": c_bits, "shard_windows": windows,
               "transforms_sharded": bool(shard_ntt),
               "exchange": "scalars: peer loads of 32 n/G bytes per rank and commitment out of rank 0's HBM; partial "
                           "sums: one XYZZ point per rank and commitment stored into rank 0's mailbox; flags in peer "
                           "memory, no collective library on the data path",
               "timing": "CUDA events on rank 0's proving stream around `steps` blocking b2p_prove calls (pinned "
                         "host columns, H2D and D2H inside)"}
    except Exception as e:  # noqa: BLE001
        res = {"error": f"{type(e).__name__}: {e}"[:300]}
    finally:
        try:
            sp.close()                      # STOP: the other ranks leave serve()
        except Exception as e:  # noqa: BLE001
            res = dict(res or {}, close_error=f"{type(e).__name__}: {e}"[:200])
    return res


def measure_sharded_ntt(args, rank, world, device, iters: int = 10):
    """The largest transform of a proof -- the coset NTT of size 4n of the quotient step -- with its domain
    sharded over the ranks (algoplonk_b200/sharded_ntt.py): cyclic coefficients in, block bit-reversed
    evaluations out, and back.  Both exchange modes: "staged" (one NCCL all_to_all_single between the two
    launches) and "p2p" (the combine / split kernel loads / stores the peers' memory over NVLink; the only
    collective is a one-element all_reduce that orders the ranks).  CUDA events on torch's current stream (the
    library launches there and NCCL orders itself on it), max over ranks.  Checked in the run: the inverse of
    the forward is the input bit for bit, and both modes produce the same evaluations.
    Every failure is reported in the line instead of raised: this leg must never cost the headline numbers."""
    import torch
    import torch.distributed as dist
    from algoplonk_b200 import sharded_ntt as sn
    curve, n = args.curve, 1 << (args.log2 + 2)
    ln = n // world
    res = {"elements": n, "elements_per_gpu": ln, "transform": "coset NTT / coset iNTT of size 4n",
           "exchange_bytes_per_gpu": 32 * ln * (world - 1) // world}

    def agree(ok: bool) -> bool:
        t = torch.tensor([0 if ok else 1], dtype=torch.int32, device=device)
        dist.all_reduce(t)
        return int(t.item()) == 0

    gen = torch.Generator(device="cpu").manual_seed(0xB200 + rank)
    coeffs = torch.randint(0, 1 << 60, (ln, 4), generator=gen, dtype=torch.int64).to(device)   # < r: canonical
    evals = {}
    for mode in ("staged", "p2p"):
        nt, err = None, ""
        try:
            nt = sn.ShardedNtt(curve, n, rank=rank, world=world, mode=mode, device=device)
        except Exception as e:  # noqa: BLE001 -- reported, see the docstring
            err = f"{type(e).__name__}: {e}"
        if not agree(nt is not None):
            res[mode] = {"error": err or "setup failed on another rank"}
            if nt is not None:
                try:
                    nt.free()
                except Exception:  # noqa: BLE001
                    pass
            continue
        try:
            ev = nt.forward(coeffs, coset=True)
            back = nt.inverse(ev, coset=True)
            ok = bool(torch.equal(back, coeffs))
            for _ in range(2):
                nt.inverse(nt.forward(coeffs, coset=True), coset=True)
            dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                nt.inverse(nt.forward(coeffs, coset=True), coset=True)
            e1.record()
            e1.synchronize()
            dist.barrier()
            ms, _ = reduce_over_ranks(e0.elapsed_time(e1) / (2 * iters), 0, world, device)
            evals[mode] = ev
            res[mode] = {"ms_per_transform": ms, "transforms_per_sec": 1e3 / ms, "round_trip_exact": ok,
                         "exchange_GBps_per_gpu_if_exchange_only": res["exchange_bytes_per_gpu"] / (ms * 1e-3) / 1e9}
        except Exception as e:  # noqa: BLE001
            res[mode] = {"error": f"{type(e).__name__}: {e}"}
        finally:
            try:
                nt.free()
            except Exception:  # noqa: BLE001
                pass
    both = agree(len(evals) == 2)            # collectively, so that every rank takes the same branch
    same = agree(both and bool(torch.equal(evals["staged"], evals["p2p"])))
    if both:
        res["modes_agree"] = same
    # the same transform on one GPU (local passes only), for the speed-up figure
    try:
        one = sn.CudaSteps(curve, n, 1, 0)
        full = torch.randint(0, 1 << 60, (n, 4), generator=gen, dtype=torch.int64).to(device)
        x = torch.empty_like(full)
        for _ in range(3):
            one.forward_local(full, n, True, x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            one.forward_local(full, n, True, x)
        e1.record()
        e1.synchronize()
        res["single_gpu_ms_per_transform"] = e0.elapsed_time(e1) / iters
        one.free()
    except Exception as e:  # noqa: BLE001
        res["single_gpu_error"] = f"{type(e).__name__}: {e}"
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--log2", type=int, default=20, help="log2 of the constraint count (BASELINE: 20)")
    ap.add_argument("--curve", default="BN254", choices=["BN254", "BLS12_381"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-solver-leg", action="store_true", help="skip the e2e_from_inputs leg (solver + prover)")
    ap.add_argument("--no-proof-sharded", action="store_true",
                    help="skip the one-proof-over-all-GPUs leg of a multi-GPU run")
    ap.add_argument("--shard-c", type=int, default=0,
                    help="window bits of the per-rank SRS blocks of the sharded proof (0: planned for the block size)")
    ap.add_argument("--no-ntt-sharded", action="store_true",
                    help="skip the domain-sharded NTT leg of a multi-GPU run")
    ap.add_argument("--inflight", type=int, default=3,
                    help="proofs in flight per GPU (one proving key + stream each; SURVEY 8d timing protocol)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        print(f"note: --warmup {args.warmup} < 3 breaks the timing rules", file=sys.stderr)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
