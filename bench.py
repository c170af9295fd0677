#!/usr/bin/env python
"""bench.py -- the headline benchmark: Fiat-Shamir sum-check prover throughput (Melem/s) on a degree-3 product of
three 2^28-entry multilinear tables (BASELINE.json configs[4]) over the reference's own field F_1572869
(triangle-counting/src/lib.rs:272-277), tables resident in HBM.

  python bench.py --gpus N --steps K --warmup W            # this engine (N>1: launched under torchrun)
  python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU path (C restatement, all host cores)

A "step" is one complete proof: Prover::new (c_1) + all v prover rounds + the Fiat-Shamir chain that supplies the
challenges (fiat_shamir::generate_transcript), i.e. first message ... last message.

N>1, headline line: WEAK scaling -- every rank holds a 2^28-entry slab of each table (total 2^(28+log2 N) entries),
one tiny exchange per pass inside the kernels.  The same line carries the STRONG-scaling curve of BASELINE.json's own
metric ("2^28 table at 1/2/4/8 B200") as `strong_scaling`: ONE 2^28-entry table set split over the N GPUs, timed the
same way (--scaling strong makes that the headline instead).

Every proof that is timed is checked: `verified` = the last timed transcript passes verify_transcript (N>1: the
sharded verifier, whose final oracle query is the sharded MLE evaluation); `sharded_equals_single` (N>1) = a sharded
proof of a 2^20-per-rank instance equals, byte for byte, the proof rank 0 computes alone from the concatenated tables.

Prints ONE JSON line on rank 0 (contract in the task statement).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

MODULUS = 1572869
BLS12_381_FR = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
K_TABLES = 3
METRIC = "sumcheck_prover_throughput"
UNIT = "Melem/s"
POLICY_NAME = {0: "small-prime 32-bit", 1: "generic 64-bit", 4: "4-limb"}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--vars", type=int, default=28, help="log2 of table entries PER GPU (weak) / in total (strong)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"], help="which curve is the headline at N>1")
    ap.add_argument("--cpu-vars", type=int, default=24, help="log2 table size of the bounded sample in the cpu_baseline leg")
    ap.add_argument("--ref-vars", type=int, default=0, help="--impl reference: log2 table size per step (0 = the workload's own size if it fits)")
    ap.add_argument("--e2e-steps", type=int, default=10)  # mean over 10 steps (every step time is printed): one host hiccup in 5 moved the mean by 7 %
    ap.add_argument("--e2e-warmup", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-strong", action="store_true", help="N>1: skip the second (strong-scaling) curve")
    ap.add_argument("--no-fields", action="store_true", help="N=1: skip the 4-limb (BLS12-381 Fr) record")
    ap.add_argument("--no-trait-leg", action="store_true", help="N=1: skip the trait-only e2e leg")
    ap.add_argument("--modulus", type=int, default=MODULUS)
    return ap.parse_args()


def workload_name(v, n_gpus, p, scaling="weak"):
    per = v if scaling == "weak" else v - (n_gpus.bit_length() - 1)
    return (f"fiat-shamir sum-check, ProductMLE<3> (degree-3 product of 3 multilinear tables), 2^{per} entries per GPU x "
            f"{n_gpus} GPU, field F_{p} (Fp64 Montgomery, 8 B/element)")


def config_dict(v_arg, n_gpus, p, n_limbs, policy, exchange="p2p", scaling="weak"):
    """The workload both arms are quoted on (same dict from `--impl ours` and `--impl reference`).  v_arg = --vars:
    per-GPU table variables for weak scaling, total variables for strong scaling."""
    E = 8 * n_limbs
    lg = n_gpus.bit_length() - 1
    v_per_gpu = v_arg if scaling == "weak" else v_arg - lg
    return {"workload": workload_name(v_arg, n_gpus, p, scaling), "tables": K_TABLES, "vars_per_gpu": v_per_gpu, "total_vars": v_per_gpu + lg,
            "field_modulus": p, "bytes_per_element": E, "arith_policy": POLICY_NAME[policy],
            "l2": "inputs (%.1f GB per GPU) larger than L2; no flush needed" % (K_TABLES * (1 << v_per_gpu) * E / 1e9),
            "parallelism": f"tables sharded by top variables over {n_gpus} GPU(s)" + ("" if n_gpus == 1 else f"; per-pass exchange: {exchange}")}


def policy_of(p):
    return 4 if p.bit_length() > 64 else (0 if p.bit_length() <= 28 else 1)


def host_threads():
    """Threads the CPU legs use: the cores this process may run on.  NOT omp_get_max_threads(): torchrun exports
    OMP_NUM_THREADS=1, which made round 1's N>1 reference arm single-threaded."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index
        self.nvml, self.stop_flag, self.t = None, False, None

    def start(self):
        # NVML in-process (a sample costs microseconds, so even a 30 ms timed region gets several); nvidia-smi otherwise
        try:
            import pynvml

            pynvml.nvmlInit()
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            pynvml.nvmlDeviceGetClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.nvml = pynvml
            self.t = threading.Thread(target=self._poll_nvml, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _poll_nvml(self):
        n = self.nvml
        bits = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))
        while not self.stop_flag:
            try:
                sm = n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)
                mx = n.nvmlDeviceGetMaxClockInfo(self.h, n.NVML_CLOCK_SM)
                mask = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.rows.append([str(sm), str(mx), "0"] + ["Active" if mask & b else "Not Active" for _, b in bits])
            except Exception:
                pass
            time.sleep(0.002)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.nvml is not None:
            self.stop_flag = True
            self.t.join(timeout=2)
        elif not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        else:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for n, val in zip(names, r[3:7]):
                    if val.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm), "source": "nvml" if self.nvml is not None else "nvidia-smi"}


def resident_stats(T):
    import ctypes as C

    n, ms, rounds = C.c_uint64(), C.c_double(), C.c_uint32()
    work, turn = (C.c_double * 40)(), (C.c_double * 40)()
    T.lib.scb_resident_stats(C.byref(n), C.byref(ms), C.byref(rounds), work, turn, 40)
    return {"launches": n.value, "total_ms": ms.value, "last_rounds": rounds.value,
            "work_us": list(work)[: rounds.value], "turn_us": list(turn)[: rounds.value]}


def pair_pass_stats(T):
    import ctypes as C

    n, ms = C.c_uint64(), C.c_double()
    T.lib.scb_pair_pass_stats(C.byref(n), C.byref(ms))
    return {"launches": n.value, "total_ms": ms.value}


def grid_pass_stats(T):
    import ctypes as C

    n, ms, wg, wp = C.c_uint64(), C.c_double(), C.c_uint64(), C.c_uint64()
    T.lib.scb_grid_pass_stats(C.byref(n), C.byref(ms), C.byref(wg), C.byref(wp))
    return {"launches": n.value, "total_ms": ms.value, "w21_grid": wg.value, "w21_pair": wp.value}


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(kernel, v, K, p):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the committed `ncu --set full` captures
    (profiles/traffic.json lists them with the capture they come from).  None for a shape that was never captured."""
    try:
        tab = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    except Exception:
        return None
    for e in tab.get("captures", []):
        if e["kernel"] == kernel and e["vars"] == v and e["tables"] == K and e["modulus"] == p:
            return float(e["dram_bytes_per_launch"])
    return None


# ------------------------------------------------------------------------------------------------ CPU legs (oracle)
def cpu_prove_sample(p, v, threads, steps=1, tabs=None):
    """Times the C restatement of the reference's prover (oracle/oracle.c::orc_product_prove: Prover::new + v rounds,
    separate fold and message passes, table copies included) on 2^v-entry tables.  Returns (Melem/s, seconds/step)."""
    from oracle.coracle import CField

    cf = CField(p)
    if tabs is None:
        tabs = [cf.synth(0xB200 + k, 0, 1 << v) for k in range(K_TABLES)]
    ch = cf.synth(0xC4A1, 0, max(v - 1, 1))
    best = None
    for _ in range(steps):
        t0 = time.perf_counter()
        cf.product_prove(tabs, ch, K_TABLES + 1, threads=threads)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return (1 << v) / best / 1e6, best


def mem_available_bytes():
    try:
        for line in open("/proc/meminfo"):
            if line.startswith("MemAvailable:"):
                return int(line.split()[1]) * 1024
    except Exception:
        pass
    return 0


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path.  The Rust crates cannot be built in this
    image (no cargo/rustc; arkworks is not vendored), so this times oracle/oracle.c -- a C restatement in the
    reference's structure -- with ALL host threads (counted from the affinity mask, whatever OMP_NUM_THREADS says).
    Same `config` as the GPU arm.  Each step is one full proof of the per-GPU share of that workload (2^vars entries
    per table: at N = 1 that IS the workload; the per-element throughput of this streaming prover does not depend on the
    table size) unless host memory or the time budget forces a smaller sample, which `cpu_baseline.sample` then says."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle.coracle import CField

    cores = host_threads()
    p = args.modulus
    n_limbs = 4 if p.bit_length() > 64 else 1
    v = args.ref_vars or args.vars
    # inputs + the folded copies the reference's structure keeps (1 + 1/2 + 1/4 .. of the tables) + slack
    need = lambda vv: int(2.3 * K_TABLES * (1 << vv) * 8 * n_limbs) + (2 << 30)
    avail = mem_available_bytes()
    while avail and need(v) > avail and v > 20:
        v -= 1
    cf = CField(p)
    tabs = [cf.synth(0xB200 + k, 0, 1 << 20) for k in range(K_TABLES)]
    _, t20 = cpu_prove_sample(p, 20, cores, tabs=tabs)  # calibration: keep the whole run within a few minutes
    budget_s = 240.0
    total_steps = max(1, args.steps) + max(0, args.warmup)
    while v > 20 and t20 * (1 << (v - 20)) * total_steps > budget_s:
        v -= 1
    tabs = [cf.synth(0xB200 + k, 0, 1 << v) for k in range(K_TABLES)]
    for _ in range(max(0, args.warmup)):
        cpu_prove_sample(p, v, cores, tabs=tabs)
    times = []
    for _ in range(max(1, args.steps)):
        _, dt = cpu_prove_sample(p, v, cores, tabs=tabs)
        times.append(dt)
    dt = sum(times) / len(times)
    val = (1 << v) / dt / 1e6
    full = v == args.vars
    sample = (f"oracle/oracle.c orc_product_prove, 2^{v}-entry tables per step "
              + ("(the workload's per-GPU size)" if full else f"(bounded sample of the 2^{args.vars} workload: host memory / time budget)")
              + f", {cores} OpenMP threads over pair ranges")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": config_dict(args.vars, args.gpus, p, n_limbs, policy_of(p), scaling=args.scaling),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample, "sample_vars": v,
                         "sample_is_full_per_gpu_workload": full},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "host_threads": cores, "omp_num_threads_env_ignored": os.environ.get("OMP_NUM_THREADS"),
        "gpu_launches": 0,
    }))


# ------------------------------------------------------------------------------------------------ GPU arm
def pass_bytes(v, K, E):
    """Algorithmic bytes of the pair passes of a 2^v proof (small-prime policy): pass over tables of 2^m entries reads
    K * 2^m entries (E bytes the first time, 4 afterwards) and writes K * 2^(m-2) packed entries (nothing after the
    last fold).  DESIGN.md section 4a."""
    out, m, in_b = [], v, E
    while m >= 3:
        out.append(K * ((1 << m) * in_b + ((1 << (m - 2)) * 4 if m >= 4 else 0)))
        m, in_b = m - 2, 4
    return out


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import thaler_study_b200 as T
    from thaler_study_b200.distributed import (CudaProductEngine, Peers, prove_sharded, prove_sharded_p2p, verify_transcript_sharded)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    # stdout carries exactly one JSON line: libraries that print banners to fd 1 (NCCL's version line) go to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    T.options_from_env()  # harness opt-in: SCB_<OPTION>=n variables become scb_set_option calls (the library reads no environment)
    n_gpus = world
    lg = world.bit_length() - 1
    p, K = args.modulus, K_TABLES
    F = T.Field(p)
    E = 8 * F.n
    headline = args.scaling if world > 1 else "weak"
    v_weak, v_strong = args.vars, args.vars - lg  # per-GPU variables of the two curves

    exchange = os.environ.get("SCB_EXCHANGE", "p2p")  # p2p: in-kernel exchange over NVLink peer memory; nccl: all-gather
    peers = None
    if world > 1 and exchange == "p2p":
        # CUDA IPC windows need peer access between the ranks' GPUs; every rank must take the same path, so the
        # outcome is agreed on with an all-reduce and NCCL all-gather is the (slower) alternative exchange
        try:
            peers = Peers()
            ok_local = 1
        except Exception as ex:  # noqa: BLE001
            print(f"[bench] rank {rank}: peer windows unavailable ({ex}); using the NCCL exchange", file=sys.stderr)
            ok_local = 0
        flag = torch.tensor([ok_local], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if flag.item() == 0:
            peers, exchange = None, "nccl"

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def all_true(ok):
        if world == 1:
            return bool(ok)
        t = torch.tensor([1 if ok else 0], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return bool(t.item())

    def prove(poly):
        if world == 1:
            return T.generate_transcript(T.Prover(poly))
        if peers is not None:
            return prove_sharded_p2p(poly, peers, consolidate_at=0)[1]
        return prove_sharded(CudaProductEngine(poly.clone()))[1]

    def verify(transcript, poly, v_local):
        if world == 1:
            return T.verify_transcript(transcript, T.Verifier(v_local, poly))
        if peers is not None:
            return verify_transcript_sharded(transcript, poly, peers)
        return len(transcript) == v_local + lg  # NCCL fallback path: no sharded oracle query available

    def timed_run(v_local, steps, warmup, collect_stats):
        """Builds the slabs (rank g owns entries [g*2^v_local, (g+1)*2^v_local) of every table: the synthetic stream is
        indexed globally), W warm-up proofs, then K proofs between barriers, CUDA events, max over ranks."""
        tabs = [T.DenseMultilinearExtension.synthetic(F, v_local, 0xB200 + k, start=rank << v_local) for k in range(K)]
        g = T.ProductMLE.new(tabs)
        T.synchronize()
        for _ in range(max(warmup, 3)):
            transcript = prove(g)
        assert len(transcript) == v_local + lg
        barrier()
        sampler = ClockSampler(local_rank) if (collect_stats and rank == 0) else None
        if sampler:
            sampler.start()
        T.launch_count(reset=True)
        T.lib.scb_resident_stats_reset()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record()
        for _ in range(steps):
            transcript = prove(g)
        ev1.record()
        barrier()
        launches = T.launch_count()
        clocks = sampler.stop() if sampler else None  # samples cover exactly the timed region
        res = resident_stats(T)  # CUDA-event time of the resident kernels launched inside the timed region
        pst = pair_pass_stats(T)
        gst = grid_pass_stats(T)
        ms = ev0.elapsed_time(ev1) / steps
        if world > 1:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        verified = all_true(verify(transcript, g, v_local))  # the LAST TIMED proof, not a separate run
        return {"ms": ms, "launches": launches, "clocks": clocks, "res": res, "pass": pst, "gridp": gst, "verified": verified, "g": g, "tabs": tabs,
                "transcript_bytes": sum(map(len, transcript))}

    # ---- the two curves
    runs = {}
    order = ["weak"] if world == 1 else ([headline] + ([] if args.no_strong and headline == "weak" else [("strong" if headline == "weak" else "weak")]))
    for which in order:
        v_local = v_weak if which == "weak" else v_strong
        r = timed_run(v_local, args.steps, args.warmup, collect_stats=(which == headline))
        r["v_local"] = v_local
        if which != headline:  # free the second curve's tables right away
            r["g"], r["tabs"] = None, None
        runs[which] = r
    main = runs[headline]
    v, ms, g, tabs = main["v_local"], main["ms"], main["g"], main["tabs"]
    total_entries = 1 << (v + lg)
    value = total_entries / (ms * 1e-3) / 1e6
    res_stats, pass_stats, grid_stats = main["res"], main["pass"], main["gridp"]

    # ---- sharded == single, byte for byte (N > 1): a 2^20-per-rank instance proved sharded and by rank 0 alone
    sharded_equals_single = None
    if world > 1:
        sv = 20
        slabs = [T.DenseMultilinearExtension.synthetic(F, sv, 0x51AB + k, start=rank << sv) for k in range(K)]
        sharded_tr = prove(T.ProductMLE.new(slabs))
        ok = True
        if rank == 0:
            full = T.ProductMLE.new([T.DenseMultilinearExtension.synthetic(F, sv + lg, 0x51AB + k) for k in range(K)])
            single_tr = T.generate_transcript(T.Prover(full))
            ok = single_tr == sharded_tr and T.verify_transcript(single_tr, T.Verifier(sv + lg, full))
            del full
        sharded_equals_single = all_true(ok)
        del slabs

    # ---- roofline of the dominant kernel
    # Small-prime fields (the default workload): two rounds per pass (csrc/pairs.cuh).  A proof of 2^28-entry tables
    # (per GPU) is three launches: k_grid_sp_pf (Prover::new: the (K+1)^2 grid sums that yield c_1, g_1 and g_2),
    # k_pair_pass_sp (the pair pass over the caller's 8-byte tables: folds two variables, accumulates the next grid;
    # sharded: its finishing thread also exchanges the 16 sums with the peers) and the resident k_persist_pairs_sp (the
    # remaining passes over packed tables).  The pair pass over the full tables is the dominant kernel; the library
    # times it with CUDA events around the launch, on its stream, for every launch of the timed region
    # (scb_pair_pass_stats); likewise the resident kernel(s) (scb_resident_stats).  The same code describes N > 1: these
    # are the kernels the sharded proof runs.
    # Other fields: one round per pass; the dominant launch is the fused fold + message kernel of round 1, timed alone.
    roof = None
    steps = args.steps
    first_alone = pass_stats["launches"] == steps  # the pass over the 8-byte tables ran as its own launch
    # option pair_w21 (K = 3, p < 2^21, first pair pass alone): the grid pass also WRITES one 8-byte word per index (the three
    # 21-bit entries) and the first pair pass reads those words instead of the 8-byte tables (csrc/pairs.cuh)
    w21 = first_alone and grid_stats["w21_grid"] == steps and grid_stats["w21_pair"] == steps
    if rank == 0 and F.policy == 0 and T.get_option("pairs") != 0 and res_stats["launches"] >= steps:
        peak, peak_src = hbm_peak()
        pbytes = pass_bytes(v, K, E)  # local passes over this rank's slab (sharded: until consolidation, then replicated)
        if w21:
            pbytes[0] = (1 << v) * 8 + K * (1 << (v - 2)) * 4
        res_ms = res_stats["total_ms"] / steps  # all resident launches of a proof (sharded: before + after consolidation)
        res_bytes = sum(pbytes[1:]) if first_alone else sum(pbytes)
        grid_bytes = K * (1 << v) * E + ((1 << v) * 8 if w21 else 0)
        if grid_stats["launches"] == steps:  # Prover::new's grid pass as the proof ran it
            gms = grid_stats["total_ms"] / steps
            g_timing = "CUDA events around the launch on its stream, all %d launches of the timed region (rank 0)" % steps
        else:
            times = []
            for i in range(3 + max(steps, 5)):  # timed alone (call includes one sync + 128 B D2H)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                g.grid_evals()
                b.record()
                torch.cuda.synchronize()
                if i >= 3:
                    times.append(a.elapsed_time(b))
            gms = sum(times) / len(times)
            g_timing = "CUDA events around the call, timed alone after the timed region (call includes one sync + 128 B D2H)"
        proof_bytes = grid_bytes + sum(pbytes)  # per GPU and entry-column: 8 (grid) + 8+1 (first pair pass) + 1.25*(1+1/4+..) = 18.7 B; w21: 8 B per index (2.7 B per entry-column) less
        # what any schedule must move: the caller's 8-byte tables once before r_1 exists; after r_1 either the same bytes again or a
        # narrower copy written and read back -- K entries of a b-bit field need ceil(K b / 8) bytes per index each way
        bits = int(p).bit_length()
        narrow = 2 * ((K * bits + 7) // 8) * (1 << v)
        floor_bytes = K * (1 << v) * E + min(K * (1 << v) * E, narrow)
        gname = "k_grid_sp_pf_w21" if w21 else "k_grid_sp_pf"
        grid = {"kernel": "%s (Prover::new: the 16 grid sums that yield c_1, g_1 and g_2; u64 input, register double buffer%s), 2^%d-entry tables"
                          % (gname + ("" if w21 else "<3>"), "; also writes one 8-byte word of three 21-bit entries per index" if w21 else "", v),
                "kernel_ms": gms, "achieved": grid_bytes / (gms * 1e-3) / 1e9, "frac": grid_bytes / (gms * 1e-3) / 1e9 / peak,
                "algorithmic_bytes_per_launch": grid_bytes, "traffic": ncu_traffic(gname, v, K, p), "share_of_step": gms / ms, "timing": g_timing}
        resident = {"kernel": "k_persist_pairs_sp<3> (the remaining pair passes, %d resident launch(es) per proof)" % (res_stats["launches"] // steps),
                    "kernel_ms": res_ms, "algorithmic_bytes_per_launch": res_bytes, "achieved": res_bytes / (res_ms * 1e-3) / 1e9,
                    "frac": res_bytes / (res_ms * 1e-3) / 1e9 / peak, "share_of_step": res_ms / ms,
                    "latency_us": {"host_turnaround_total_last_launch": sum(res_stats["turn_us"]), "device_passes_total_last_launch": sum(res_stats["work_us"])},
                    "traffic": None, "traffic_note": "ncu serialises kernel and host, so a resident kernel cannot run under it"}
        proof = {"proof_bytes_moved_per_gpu": proof_bytes, "proof_frac_of_hbm_roofline": (proof_bytes / (ms * 1e-3) / 1e9) / peak,
                 "proof_floor_bytes_per_gpu": floor_bytes, "proof_frac_vs_floor": (floor_bytes / (ms * 1e-3) / 1e9) / peak,
                 "proof_floor_note": "any schedule must read the caller's 8-byte tables once before r_1 exists (K*2^v*E) and afterwards either read them "
                                     "again or write and read back a narrower copy (2*ceil(K*bits(p)/8)*2^v)",
                 "proof_survey_bytes": 4.0 * K * (1 << v) * E,
                 "proof_survey_note": "SURVEY 8d's 4*K*2^v*E assumes one round per pass and 8-byte intermediates; two rounds per pass with "
                                      "uint32 intermediates move fewer bytes, so a fraction against it can exceed 1 and is not reported",
                 "w21_triples": bool(w21)}
        if first_alone:
            pms = pass_stats["total_ms"] / pass_stats["launches"]
            pname = "k_pair_pass_sp_w21" if w21 else "k_pair_pass_sp"
            pair = {"kernel": "%s (rounds 3-4 of the proof: folds two variables of the 2^%d-entry tables and accumulates the 16 grid sums of the next two "
                              "messages%s)" % ("k_pair_pass_sp_w21 <in = 21-bit triples, out = u32>" if w21 else "k_pair_pass_sp<3,in=u64,out=u32>", v,
                                               "" if world == 1 else "; finishing thread exchanges them with the peers"),
                    "achieved": pbytes[0] / (pms * 1e-3) / 1e9, "frac": pbytes[0] / (pms * 1e-3) / 1e9 / peak,
                    "traffic": ncu_traffic(pname, v, K, p), "kernel_ms": pms, "algorithmic_bytes_per_launch": pbytes[0],
                    "timing": "CUDA events around the launch on its stream, all %d launches of the timed region (rank 0)" % pass_stats["launches"],
                    "share_of_step": pms / ms}
            # the dominant kernel of the step heads the block; the other two follow
            top, other_key, other = (grid, "first_pair_pass_kernel", pair) if gms >= pms else (pair, "grid_kernel", grid)
            roof = {"bound": "hbm", "kernel": top["kernel"], "achieved": top["achieved"], "peak": peak, "unit": "GB/s", "frac": top["frac"],
                    "traffic": top["traffic"], "kernel_ms": top["kernel_ms"], "algorithmic_bytes_per_launch": top["algorithmic_bytes_per_launch"],
                    "peak_source": peak_src, "timing": top["timing"], "share_of_step": top["share_of_step"], other_key: other,
                    "resident_kernel": resident}
            if roof["frac"] > 1.0:
                roof["frac_note"] = ("MEASURED_PEAKS.json hbm_gbs is a copy (1 : 1 read/write mix); this pass reads %d bytes for every byte it writes, "
                                     "and read-heavy streams run a few per cent above the copy figure on these boxes (the read-only grid pass without "
                                     "the triples: 0.99-1.00 of it)" % (K * E // 8) if w21 else "read-only stream measured against the copy (1 : 1 read/write) figure")
        else:
            roof = {"bound": "hbm", "kernel": resident["kernel"], "achieved": resident["achieved"], "peak": peak, "unit": "GB/s",
                    "frac": resident["frac"], "traffic": None, "traffic_note": resident["traffic_note"], "kernel_ms": res_ms,
                    "algorithmic_bytes_per_launch": res_bytes, "peak_source": peak_src,
                    "timing": "CUDA events around the launches on their stream, all %d launches of the timed region (rank 0)" % res_stats["launches"],
                    "share_of_step": res_ms / ms, "latency_us": resident["latency_us"], "grid_kernel": grid}
        roof.update(proof)
    elif rank == 0:
        roof = generic_roofline(T, torch, F, g, v, K, p, ms, steps, res_stats, world)

    # ---- e2e: the same proof through the C ABI with HOST tables (pinned), H2D inside the timed region
    e2e = None
    if not args.no_e2e:
        host = []
        for t in tabs:
            ht = torch.empty([1 << v, F.n], dtype=torch.int64, pin_memory=True)
            dt_ = torch.empty([1 << v, F.n], dtype=torch.int64, device="cuda")  # device -> pinned host staging (outside the timed region)
            T._lib.check(T.lib.scb_mle_copy_to_device(t._h, dt_.data_ptr()))
            ht.copy_(dt_)
            del dt_
            host.append(ht.numpy().view(np.uint64))
        torch.cuda.synchronize()

        # One GPU: scb_poly_product_from_host -- for the small-prime field the tables cross PCIe as packed uint32 where
        # the host cores keep up and as 8-byte entries where not (narrowed on the device).  Sharded runs keep one plain
        # cudaMemcpy per table: the ranks of a box share its host memory system, which is what bounds the narrowing
        # upload (profiles/r01_upload_ab.md; SCB_BENCH_E2E_UPLOAD=1 selects it anyway).
        one_call = (world == 1 or os.environ.get("SCB_BENCH_E2E_UPLOAD", "0") == "1") and os.environ.get("SCB_BENCH_E2E_PLAIN", "0") == "0"

        split = {"upload_ms": 0.0, "prove_ms": 0.0}  # host clock of the last step's two parts (both block until done)

        def e2e_step(tables):
            t0 = time.perf_counter()
            if one_call:
                gg = T.ProductMLE.from_host_tables(F, v, tables)
            else:
                hs = [T.DenseMultilinearExtension.from_evaluations_vec(F, v, h) for h in tables]  # cudaMemcpy H2D
                gg = T.ProductMLE.new(hs)
            t1 = time.perf_counter()
            tr = prove(gg)  # messages come back device -> host every round
            split["upload_ms"], split["prove_ms"] = (t1 - t0) * 1e3, (time.perf_counter() - t1) * 1e3
            return tr, gg

        step_log = {}

        def time_e2e(tables, n_steps, n_warm=1, tag="pinned"):
            warm = []
            for _ in range(max(1, n_warm)):
                tw = time.perf_counter()
                e2e_step(tables)
                warm.append(round((time.perf_counter() - tw) * 1e3, 2))
            barrier()
            t0 = time.perf_counter()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            per = []
            tr = gg = None
            for _ in range(n_steps):
                tr = gg = None  # the previous step's handles are dropped before the next upload, as a caller's loop would
                ts = time.perf_counter()
                tr, gg = e2e_step(tables)
                per.append(round((time.perf_counter() - ts) * 1e3, 2))
            b.record()
            barrier()
            step_log[tag] = {"warmup_ms": warm, "timed_ms": per}
            wall = (time.perf_counter() - t0) / n_steps
            ems = max(a.elapsed_time(b) / n_steps, wall * 1e3)
            if world > 1:
                t = torch.tensor([ems], device="cuda", dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ems = float(t.item())
            return ems, all_true(verify(tr, gg, v))

        ems, e2e_ok = time_e2e(host, args.e2e_steps, args.e2e_warmup)
        rounds = v + lg
        h2d = K * (1 << v) * E * n_gpus
        note = "scb_mle_from_host x%d (pinned host tables, H2D) + Prover::new + all rounds + Fiat-Shamir, per step" % K
        upload = None
        if one_call:
            import ctypes

            pc, rc_, hb = ctypes.c_uint64(), ctypes.c_uint64(), ctypes.c_uint64()
            T._lib.check(T.lib.scb_host_pack_stats(ctypes.byref(pc), ctypes.byref(rc_), ctypes.byref(hb)))
            note = "scb_poly_product_from_host (pinned host tables, H2D) + Prover::new + all rounds + Fiat-Shamir, per step"
            if F.policy == 0 and pc.value + rc_.value > 0 and T.get_option("host_pack") != 0:
                h2d = hb.value  # bytes of the copies the last step queued, counted by the library as it queued them
                if world > 1:
                    tb = torch.tensor([h2d], device="cuda", dtype=torch.int64)
                    dist.all_reduce(tb, op=dist.ReduceOp.SUM)
                    h2d = int(tb.item())
                wire21 = p < (1 << 21) and T.get_option("host_pack_wire") == 21
                upload = {"chunk_entries": K * (1 << v) // (pc.value + rc_.value), "chunks_narrowed_on_host": pc.value,
                          "chunks_narrowed_on_device": rc_.value,
                          "host_lane_wire_format": "three 21-bit entries per 64-bit word" if wire21 else "uint32",
                          "host_pack_threads_per_rank": min(32, T.get_option("host_pack_threads") or max(1, host_threads() // world))}
        e2e = {"value": total_entries / (ems * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": ems, "steps": args.e2e_steps, "warmup": max(1, args.e2e_warmup),
               "step_ms_rank0": step_log["pinned"], "verified": e2e_ok, "last_step_split_rank0": dict(split),
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": rounds * (K + 1) * E * n_gpus, "note": note}
        if upload:
            e2e["upload"] = upload
        # what a Rust caller has: tables in PAGEABLE memory (Vec<F>); the device-side narrowing lane needs pinned memory
        # and is skipped, the host lane stages through the library's pinned buffers
        if world == 1 and one_call:
            pageable = [np.array(h, copy=True) for h in host]
            pms_, pok = time_e2e(pageable, max(1, args.e2e_steps - 1), 2, "pageable")
            e2e["pageable_host_tables"] = {"value": total_entries / (pms_ * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": pms_, "verified": pok, "step_ms_rank0": step_log["pageable"],
                                           "note": "same call with the tables in pageable memory (what `Vec<F>::as_ptr()` gives a Rust shim)"}
            del pageable
        # The drop-in through the TRAIT ONLY: exactly the calls an unmodified sum_check_protocol::Prover<F, GpuPoly> and
        # fiat_shamir::generate_transcript issue (sum-check-protocol/src/lib.rs:88-112): Prover::new = to_evaluations()
        # (the 2^v-entry product table device -> host) + a host sum; per round fix_variables then to_univariate as two
        # separate calls; the hash chain on the host.  No fused call, no pair passes, no resident kernel, no packing.
        if world == 1 and not args.no_trait_leg:
            e2e["trait_only"] = trait_only_leg(T, np, torch, F, v, K, host, total_entries)
        del host

    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        cv = args.cpu_vars
        v1, t1 = cpu_prove_sample(p, cv, 1)
        cores = host_threads()
        vN, tN = cpu_prove_sample(p, cv, cores)
        cpu = {"value": vN, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"oracle/oracle.c orc_product_prove on 2^{cv}-entry tables (bounded sample), {cores} OpenMP threads; "
                         f"single-thread (reference-faithful) = {v1:.2f} {UNIT}",
               "single_thread_value": v1, "seconds": tN}

    # ---- second field: the 4-limb path (north_star kernel #1, "multi-limb Montgomery"), N = 1 only
    fields = None
    if world == 1 and not args.no_fields and p == MODULUS:
        del g, tabs
        main["g"] = main["tabs"] = None
        torch.cuda.empty_cache()
        fields = {"bls12_381_fr": field_record(T, torch, BLS12_381_FR, args.vars, K, max(3, min(args.steps, 5)), not args.no_cpu_baseline)}

    if rank == 0:
        cfg = config_dict(args.vars, n_gpus, p, F.n, F.policy, exchange, scaling=headline)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms, "higher_is_better": True, "scaling": headline, "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": cfg, "verified": main["verified"], "sharded_equals_single": sharded_equals_single,
            "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": main["launches"], "clocks": main["clocks"],
            "host_threads": host_threads(), "transcript_bytes": main["transcript_bytes"],
        }
        other = "strong" if headline == "weak" else "weak"
        if other in runs:
            o = runs[other]
            tot = 1 << (o["v_local"] + lg)
            line[other + "_scaling"] = {"scaling": other, "value": tot / (o["ms"] * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": o["ms"],
                                        "vars_per_gpu": o["v_local"], "total_vars": o["v_local"] + lg, "verified": o["verified"],
                                        "gpu_launches": o["launches"], "steps": args.steps,
                                        "note": ("ONE 2^%d-entry table set split over the %d GPUs (BASELINE.json metric)" % (args.vars, n_gpus)) if other == "strong"
                                                else "2^%d entries per GPU" % o["v_local"]}
        if fields:
            line["fields"] = fields
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        if peers is not None:
            peers.close()
        dist.destroy_process_group()


def generic_roofline(T, torch, F, g, v, K, p, ms, steps, res_stats, world):
    """One round per pass (generic / 4-limb fields, or option pairs = 0).  Timed alone, on the kernel variant the proof
    runs in round 1: with the small-prime policy the folded tables are packed uint32, so the launch reads 2^v ark
    elements (E bytes) per table and writes 2^(v-1) 4-byte entries; otherwise it writes 2^(v-1) E-byte elements
    (SURVEY 8d's 1.5*K*2^v*E)."""
    E = 8 * F.n
    packed = F.policy == 0 and T.get_option("packed") != 0
    gk = g.clone().allow_packed(packed)
    d_out = torch.empty([K + 1, F.n], dtype=torch.int64, device="cuda")
    r = 123456 % p
    claim = None
    if F.policy == 4:
        # what Prover::round's caller knows before the pass: g_1(r_1) of the first message (the 4-limb kernel uses it)
        claim = T.evals_to_univariate(F, T.KIND_PRODUCT, g.round_evals()).evaluate(r)
    times = []
    for i in range(3 + max(steps, 5)):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        if claim is not None:
            nxt, _ = gk.fix_and_round_evals(r, claim=claim)  # includes the wait for the sums and their host rebuild (microseconds)
        else:
            nxt = gk.fix_and_round_evals_device(r, d_out.data_ptr())
        b.record()
        torch.cuda.synchronize()
        if i >= 3:
            times.append(a.elapsed_time(b))
        del nxt
    kms = sum(times) / len(times)
    alg_bytes = K * (1 << v) * (E + (2 if packed else E / 2))
    proof_bytes = K * (1 << v) * (3.0 * E if packed else 4.0 * E)  # packed: 8+8+2+3*(1+1/2+..)=24 B per entry-column
    peak, peak_src = hbm_peak()
    achieved = alg_bytes / (kms * 1e-3) / 1e9
    kname = {0: "k_fold_round_sp<3,in=u64,out=u32>", 1: "k_fold_round<PolG1,3>", 4: "fused fold+message kernel of the 4-limb policy (option g4_kernel=%d)" % T.get_option("g4_kernel")}[F.policy] \
        if (packed or F.policy != 0) else "k_fold_round<PolSP,3>"
    tkey = {0: "k_fold_round_sp", 1: "k_fold_round_g1", 4: "k_fold_round_g4"}[F.policy]
    # integer roofline: modmuls per launch x 32x32->64 multiply-adds per modmul / their measured issue rate.
    # A wide multiply-add (IMAD.WIDE.U32, with or without carry, IMAD.HI) issues once per 4 cycles per SM sub-partition on
    # the fmaheavy pipe: 148 SMs x 4 x 32 lanes / 4 cycles x 1.965 GHz = 9.31 T/s (scripts/mont29_bench.cu under ncu:
    # fmaheavy 91-95 % active at 0.24 warp instructions per cycle; profiles/r02_mont29.md).  Round 1-2 divided by 17.9 T/s,
    # which was the rate of IADD3 pairs -- ptxas had hoisted that microbenchmark's loop-invariant product.
    imad_peak = 9.31e12
    imads_per_mul = {0: 2.5, 1: 10, 4: 128}[F.policy]  # wide multiply-adds (a plain 32-bit IMAD counts one half)
    modmuls_launch = (2 * K + (K + 1) * (K - 1)) * (1 << v) / 4.0  # fold K per output + message at X = 0..K
    g4_mode = T.get_option("g4_kernel") if F.policy == 4 else 0
    g4_on = g4_mode != 0
    modmuls_executed = ((2 * K + K * (K - 1)) if g4_on else (2 * K + (K + 1) * (K - 1))) * (1 << v) / 4.0  # g4.cuh skips one point
    wide_executed = modmuls_executed * imads_per_mul
    if g4_mode == 3 and K >= 2:
        # last product of each point unreduced (64); p = 1 mod 2^32: 120 per reduced product; folds with the pass's table: 80 / 78
        p0one = (p & 0xFFFFFFFF) == 1 and T.get_option("g4_p0one")
        red, fold = (120, 78) if p0one else (128, 80)
        wide_executed = (2 * K * fold + K * (K - 2) * red + K * 64) * (1 << v) / 4.0
    modmuls_proof = (K * K + K - 1) * float(1 << v)
    # The integer bound is taken on the multiply-adds the best formulation in this repo needs (what the kernels execute):
    # SURVEY's count -- (K^2 + K - 1) 2^v modmuls x 128 -- is reported next to it as `survey_count_*`; the kernels need
    # fewer (one point fewer, unreduced last products, table folds, an interpolated point), so a fraction against the
    # survey count could exceed 1 and is not what the kernel is judged by.
    wide_proof = 2.0 * wide_executed + (K + 1) * (K - 1) * imads_per_mul * (1 << v) / 2.0  # later rounds halve: sum = 2 launches; round 0: X = 0..K
    if g4_mode == 3 and K >= 2:
        r0 = (3 * red + 4 * 64) if K == 3 else (K + 1) * ((K - 2) * red + 64)  # K = 3: first-level product at X = 2 interpolated
        wide_proof = 2.0 * wide_executed + r0 * (1 << v) / 2.0
    t_int_launch = wide_executed / imad_peak
    t_hbm_launch = alg_bytes / (peak * 1e9)
    t_int_proof = wide_proof / imad_peak
    t_hbm_proof = proof_bytes / (peak * 1e9)
    survey_int_launch = modmuls_launch * imads_per_mul / imad_peak
    survey_int_proof = modmuls_proof * imads_per_mul / imad_peak
    alone = {"kernel": kname + ", 2^%d-entry tables, timed alone" % v,
             "achieved": achieved, "frac": achieved / peak, "kernel_ms": kms, "algorithmic_bytes_per_launch": alg_bytes,
             "traffic": ncu_traffic(tkey, v, K, p),
             "modmuls_per_launch": modmuls_launch, "modmuls_executed_per_launch": modmuls_executed, "imad_wide_per_modmul": imads_per_mul, "imad_wide_peak_per_s": imad_peak,
             "imad_wide_executed_per_launch": wide_executed, "frac_of_fmaheavy_issue_peak": wide_executed / imad_peak / (kms * 1e-3),
             "t_integer_bound_ms": t_int_launch * 1e3, "t_hbm_bound_ms": t_hbm_launch * 1e3,
             "frac_of_slower_bound": max(t_int_launch, t_hbm_launch) / (kms * 1e-3),
             "survey_count_t_integer_ms": survey_int_launch * 1e3,
             "survey_count_note": "SURVEY 8d's modmul count x 128 multiply-adds at the measured multiplier peak; the kernels execute fewer (imad_wide_executed_per_launch), so the bound above is taken on what they execute"}
    proof = {"proof_bytes_moved": proof_bytes, "proof_frac_of_hbm_roofline": (proof_bytes / (ms * 1e-3) / 1e9) / peak,
             "proof_modmuls": modmuls_proof, "proof_imad_wide_executed": wide_proof, "proof_t_integer_bound_ms": t_int_proof * 1e3, "proof_t_hbm_bound_ms": t_hbm_proof * 1e3,
             "proof_frac_of_slower_bound": max(t_int_proof, t_hbm_proof) / (ms * 1e-3), "proof_survey_count_t_integer_ms": survey_int_proof * 1e3}
    int_bound = t_int_launch > t_hbm_launch
    roof = {"bound": "hbm", "kernel": alone["kernel"], "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": alone["traffic"], "kernel_ms": kms, "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peak_src,
            "share_of_step": kms / ms,
            "binding_roofline": "integer (IMAD.WIDE)" if int_bound else "hbm",
            "note": "contract fields describe the HBM side; for this policy the integer pipe binds -- see frac_of_slower_bound" if int_bound else None,
            "integer": {k: alone[k] for k in ("modmuls_per_launch", "modmuls_executed_per_launch", "imad_wide_per_modmul", "imad_wide_peak_per_s", "imad_wide_executed_per_launch",
                                             "frac_of_fmaheavy_issue_peak", "t_integer_bound_ms", "t_hbm_bound_ms", "frac_of_slower_bound", "survey_count_t_integer_ms", "survey_count_note")}}
    if world == 1 and res_stats["launches"] >= steps and res_stats["last_rounds"] > 0:
        res_ms = res_stats["total_ms"] / steps
        roof["resident_kernel"] = {"kernel_ms": res_ms, "rounds_last_launch": res_stats["last_rounds"], "share_of_step": res_ms / ms,
                                   "latency_us": {"host_turnaround_total": sum(res_stats["turn_us"]), "device_rounds_total": sum(res_stats["work_us"])}}
    roof.update(proof)
    return roof


def field_record(T, torch, p, v, K, steps, with_cpu):
    """A second record inside the JSON line: the same proof over another field (the 4-limb BLS12-381 Fr =
    ark_ed_on_bls12_381::Fq of the reference's Cargo.toml:20), device-timed, verified, with its own roofline (integer
    vs HBM, the slower one) and CPU port number."""
    F = T.Field(p)
    tabs = [T.DenseMultilinearExtension.synthetic(F, v, 0xB200 + k) for k in range(K)]
    g = T.ProductMLE.new(tabs)
    T.synchronize()
    for _ in range(3):
        tr = T.generate_transcript(T.Prover(g))
    torch.cuda.synchronize()
    T.launch_count(reset=True)
    T.lib.scb_resident_stats_reset()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(steps):
        tr = T.generate_transcript(T.Prover(g))
    ev1.record()
    torch.cuda.synchronize()
    launches = T.launch_count()
    ms = ev0.elapsed_time(ev1) / steps
    res = resident_stats(T)
    ok = T.verify_transcript(tr, T.Verifier(v, g))
    roof = generic_roofline(T, torch, F, g, v, K, p, ms, steps, res, 1)
    rec = {"metric": METRIC, "value": (1 << v) / (ms * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": ms, "steps": steps, "verified": ok,
           "config": {"workload": f"fiat-shamir sum-check, ProductMLE<3>, 2^{v} entries, 4-limb field (BLS12-381 Fr, 32 B/element)",
                      "field_modulus_bits": F.bits, "bytes_per_element": 8 * F.n, "arith_policy": POLICY_NAME[F.policy]},
           "gpu_launches": launches, "roofline": roof}
    if with_cpu:
        cv = 20
        cores = host_threads()
        vN, tN = cpu_prove_sample(p, cv, cores)
        rec["cpu_baseline"] = {"value": vN, "unit": UNIT, "cores": cores, "kind": "port", "seconds": tN,
                               "sample": f"oracle/oracle.c orc_product_prove on 2^{cv}-entry tables (bounded sample), {cores} OpenMP threads"}
    del g, tabs
    torch.cuda.empty_cache()
    return rec


def trait_only_leg(T, np, torch, F, v, K, host, total_entries):
    """Times exactly what an unmodified `Prover<F, GpuPoly>` + `generate_transcript` would do through the five trait
    methods (sum-check-protocol/src/lib.rs:88-112, fiat-shamir/src/lib.rs:75-98), host tables in, transcript out."""
    p = F.p

    def step():
        mles = [T.DenseMultilinearExtension.from_evaluations_vec(F, v, h) for h in host]  # from_evaluations_vec: H2D
        g = T.ProductMLE.new(mles)
        # Prover::new: g.to_evaluations().into_iter().sum()
        n = 1 << v
        ev = np.empty((n, F.n), dtype=np.uint64)
        T._lib.check(T.lib.scb_poly_to_evaluations(g._h, T.api._p64(ev), n))
        if F.n == 1:
            c1_m = int(ev[:, 0].sum(dtype=np.uint64) % np.uint64(p)) if p < (1 << 36) else int(sum(int(x) for x in ev[:, 0]) % p)
            c_1 = F.from_mont(np.array([[c1_m]], dtype=np.uint64))[0]  # Montgomery form is linear: sum of residues, then one conversion
        else:
            c_1 = sum(F.from_mont(ev)) % p
        # generate_transcript: g_1 = (c_1, round(1, 0)); then r_j = hash(transcript so far), round(r_j, j)
        msgs, sofar = [], b""
        for j in range(v):
            if j > 0:
                r = F.hash_to_field(sofar)
                g = g.fix_variables([r])        # self.g = self.g.fix_variables(&[r])
            poly = g.to_univariate()            # self.g.to_univariate()
            m = (c_1.to_bytes(F.ser_bytes, "little") if j == 0 else b"") + poly.serialize_uncompressed()
            msgs.append(m)
            sofar += m
        return msgs, g

    msgs, _ = step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    msgs, _ = step()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    fast = T.generate_transcript(T.Prover(T.ProductMLE.from_host_tables(F, v, host)))
    return {"value": total_entries / dt / 1e6, "unit": UNIT, "ms_per_step": dt * 1e3, "equals_fast_path_bytes": msgs == fast,
            "h2d_bytes_per_step": K * (1 << v) * 8 * F.n, "d2h_bytes_per_step": (1 << v) * 8 * F.n + v * (K + 1) * 8 * F.n,
            "note": "only the five SumCheckPolynomial methods: from_evaluations_vec (H2D), Prover::new = to_evaluations() (2^v-entry D2H) + host sum, "
                    "per round fix_variables then to_univariate (two launches), host hash chain; no fused/pair/resident/packed path"}


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
