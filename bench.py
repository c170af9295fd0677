#!/usr/bin/env python
"""bench.py -- the headline benchmark: Fiat-Shamir sum-check prover throughput (Melem/s) on a degree-3 product of
three 2^28-entry multilinear tables (BASELINE.json configs[4]) over the reference's own field F_1572869
(triangle-counting/src/lib.rs:272-277), tables resident in HBM.

  python bench.py --gpus N --steps K --warmup W            # this engine (N>1: launched under torchrun)
  python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU path (C restatement, all host cores)

A "step" is one complete proof: Prover::new (c_1) + all v prover rounds + the Fiat-Shamir chain that supplies the
challenges (fiat_shamir::generate_transcript), i.e. first message ... last message.  N>1: weak scaling, every rank
holds a 2^28-entry slab of each table (total 2^(28+log2 N) entries), one tiny all-gather per round.
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
K_TABLES = 3
METRIC = "sumcheck_prover_throughput"
UNIT = "Melem/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--vars", type=int, default=28, help="log2 of table entries PER GPU")
    ap.add_argument("--cpu-vars", type=int, default=24, help="log2 table size of the bounded CPU sample")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--modulus", type=int, default=MODULUS)
    return ap.parse_args()


def workload_name(v, n_gpus, p):
    return (f"fiat-shamir sum-check, ProductMLE<3> (degree-3 product of 3 multilinear tables), 2^{v} entries per GPU x "
            f"{n_gpus} GPU, field F_{p} (Fp64 Montgomery, 8 B/element)")


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


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------------------------ CPU legs (oracle)
def cpu_prove_sample(p, v, threads, steps=1):
    """Times the C restatement of the reference's prover (oracle/oracle.c::orc_product_prove: Prover::new + v rounds,
    separate fold and message passes, table copies included) on a 2^v sample.  Returns (Melem/s, seconds/step)."""
    import numpy as np
    from oracle.coracle import CField

    cf = CField(p)
    tabs = [cf.synth(0xB200 + k, 0, 1 << v) for k in range(K_TABLES)]
    ch = cf.synth(0xC4A1, 0, max(v - 1, 1))
    best = None
    for _ in range(steps):
        t0 = time.perf_counter()
        cf.product_prove(tabs, ch, K_TABLES + 1, threads=threads)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return (1 << v) / best / 1e6, best


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path.  The Rust crates cannot be built in this
    image (no cargo/rustc; arkworks is not vendored), so this times oracle/oracle.c -- a C restatement in the
    reference's structure -- with all host threads, on a bounded 2^cpu_vars sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle.coracle import max_threads

    cores = max_threads()
    v = args.cpu_vars
    for _ in range(max(0, min(args.warmup, 1))):
        cpu_prove_sample(args.modulus, v, cores)
    times = []
    for _ in range(max(1, args.steps)):
        _, dt = cpu_prove_sample(args.modulus, v, cores)
        times.append(dt)
    dt = sum(times) / len(times)
    val = (1 << v) / dt / 1e6
    sample = f"2^{v}-entry tables (bounded sample of the 2^{args.vars} workload), {cores} OpenMP threads over pair ranges"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": workload_name(args.vars, args.gpus, args.modulus), "sample": sample},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ------------------------------------------------------------------------------------------------ GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist

    import thaler_study_b200 as T
    from thaler_study_b200.distributed import CudaProductEngine, Peers, prove_sharded, prove_sharded_p2p

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
    n_gpus = world
    v, p, K = args.vars, args.modulus, K_TABLES
    F = T.Field(p)
    E = 8 * F.n
    # rank g owns entries [g*2^v, (g+1)*2^v) of every table: the synthetic stream is indexed globally
    tabs = [T.DenseMultilinearExtension.synthetic(F, v, 0xB200 + k, start=rank << v) for k in range(K)]
    g = T.ProductMLE.new(tabs)
    T.synchronize()

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

    def prove(poly):
        if world == 1:
            return T.generate_transcript(T.Prover(poly))
        if peers is not None:
            return prove_sharded_p2p(poly, peers)[1]
        return prove_sharded(CudaProductEngine(poly.clone()))[1]

    def step():
        return prove(g)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        transcript = step()
    assert len(transcript) == v + (world.bit_length() - 1)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    T.launch_count(reset=True)
    T.lib.scb_resident_stats_reset()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    barrier()
    launches = T.launch_count()
    clocks = sampler.stop() if rank == 0 else None  # samples cover exactly the timed region
    res_stats = resident_stats(T)  # CUDA-event time of the resident kernels launched inside the timed region
    ms = ev0.elapsed_time(ev1) / args.steps
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    total_entries = (1 << v) * n_gpus
    value = total_entries / (ms * 1e-3) / 1e6

    # ---- roofline of the dominant kernel
    # Small-prime fields (the default workload): two rounds per pass (csrc/pairs.cuh).  A proof of 2^28-entry tables is
    # three launches: k_grid_sp_pf (Prover::new: the (K+1)^2 grid sums that yield c_1, g_1 and g_2), k_pair_pass_sp (the
    # pair pass over the caller's 8-byte tables: folds two variables, accumulates the next grid) and the resident
    # k_persist_pairs_sp (the 12 remaining passes over packed tables).  The pair pass over the full tables is the
    # dominant kernel; the library times it with CUDA events around the launch, on its stream, for every launch inside
    # the timed region above (scb_pair_pass_stats); likewise the resident kernel (scb_resident_stats).
    # Other fields: one round per pass; the dominant launch is the fused fold + message kernel of round 1, timed alone.
    roof = None
    pass_stats = pair_pass_stats(T)
    first_alone = pass_stats["launches"] == args.steps  # the pass over the 8-byte tables ran as its own launch
    n_res_passes = (v - 1) // 2 - (1 if first_alone else 0)
    if rank == 0 and F.policy == 0 and world == 1 and os.environ.get("SCB_PAIRS", "1") != "0" \
            and res_stats["launches"] == args.steps and res_stats["last_rounds"] == n_res_passes:
        peak, peak_src = hbm_peak()
        pbytes, m, in_b = [], v, E
        while m >= 3:  # pass: read K tables of 2^m entries, write 2^(m-2) packed entries (nothing after the last fold)
            pbytes.append(K * ((1 << m) * in_b + ((1 << (m - 2)) * 4 if m >= 4 else 0)))
            m, in_b = m - 2, 4
        res_ms = res_stats["total_ms"] / res_stats["launches"]
        res_bytes = sum(pbytes[1:]) if first_alone else sum(pbytes)
        times = []
        for i in range(3 + max(args.steps, 5)):  # Prover::new's grid kernel, timed alone (call includes one sync + 128 B D2H)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            g.grid_evals()
            b.record()
            torch.cuda.synchronize()
            if i >= 3:
                times.append(a.elapsed_time(b))
        gms = sum(times) / len(times)
        grid_bytes = K * (1 << v) * E
        std = v == 28 and K == 3 and p == MODULUS
        proof_bytes = grid_bytes + sum(pbytes)
        grid = {"kernel": "k_grid_sp_pf<3> (Prover::new; u64 input, register double buffer), 2^%d-entry tables" % v, "kernel_ms": gms,
                "achieved": grid_bytes / (gms * 1e-3) / 1e9, "frac": grid_bytes / (gms * 1e-3) / 1e9 / peak,
                "algorithmic_bytes_per_launch": grid_bytes, "traffic": 6.457139e9 if std else None, "share_of_step": gms / ms,
                "timing": "CUDA events around the call, timed alone after the timed region (call includes one sync + 128 B D2H)"}
        resident = {"kernel": "k_persist_pairs_sp<3> (%d resident pair passes: the remaining rounds, one cooperative launch)" % n_res_passes,
                    "kernel_ms": res_ms, "algorithmic_bytes_per_launch": res_bytes, "achieved": res_bytes / (res_ms * 1e-3) / 1e9,
                    "frac": res_bytes / (res_ms * 1e-3) / 1e9 / peak, "share_of_step": res_ms / ms,
                    "latency_us": {"host_turnaround_total": sum(res_stats["turn_us"]), "device_passes_total": sum(res_stats["work_us"])},
                    "traffic": None, "traffic_note": "ncu serialises kernel and host, so a resident kernel cannot run under it"}
        proof = {"proof_bytes_moved": proof_bytes, "proof_frac_of_hbm_roofline": (proof_bytes / (ms * 1e-3) / 1e9) / peak,
                 "proof_survey_bytes": 4.0 * K * (1 << v) * E,
                 "proof_frac_vs_survey_bytes": (4.0 * K * (1 << v) * E / (ms * 1e-3) / 1e9) / peak}
        if first_alone:
            # dominant kernel of the step: the pair pass over the caller's 8-byte tables, an ordinary launch timed with CUDA
            # events around it on its stream, every launch of the timed region (scb_pair_pass_stats)
            pms = pass_stats["total_ms"] / pass_stats["launches"]
            roof = {"bound": "hbm",
                    "kernel": "k_pair_pass_sp<3,in=u64,out=u32> (rounds 3-4 of the proof: folds two variables of the 2^%d-entry tables and "
                              "accumulates the 16 grid sums of the next two messages)" % v,
                    "achieved": pbytes[0] / (pms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s", "frac": pbytes[0] / (pms * 1e-3) / 1e9 / peak,
                    "traffic": 7.241953e9 if std else None, "kernel_ms": pms, "algorithmic_bytes_per_launch": pbytes[0], "peak_source": peak_src,
                    "timing": "CUDA events around the launch on its stream, all %d launches of the timed region" % pass_stats["launches"],
                    "share_of_step": pms / ms, "grid_kernel_alone": grid, "resident_kernel": resident}
        else:
            w0 = res_stats["work_us"][0]
            roof = {"bound": "hbm", "kernel": resident["kernel"], "achieved": resident["achieved"], "peak": peak, "unit": "GB/s",
                    "frac": resident["frac"], "traffic": None, "traffic_note": resident["traffic_note"], "kernel_ms": res_ms,
                    "algorithmic_bytes_per_launch": res_bytes, "peak_source": peak_src,
                    "timing": "CUDA events around the launch on its stream, all %d launches of the timed region" % res_stats["launches"],
                    "share_of_step": res_ms / ms,
                    "pass0_phase": {"us": w0, "achieved": pbytes[0] / (w0 * 1e-6) / 1e9, "frac": pbytes[0] / (w0 * 1e-6) / 1e9 / peak,
                                    "bytes": pbytes[0], "source": "%globaltimer stamps inside the kernel, last launch"},
                    "pass0_alone_traffic": 7.241953e9 if std else None, "latency_us": resident["latency_us"], "grid_kernel_alone": grid}
        roof.update(proof)
    elif rank == 0:
        # Timed alone, on the same kernel variant the proof runs in round 1: with the small-prime policy the prover's
        # folded tables are packed uint32 (packed.cuh), so the launch reads 2^v ark elements (E bytes) per table and
        # writes 2^(v-1) 4-byte entries; otherwise it writes 2^(v-1) E-byte elements (SURVEY 8d's 1.5*K*2^v*E).
        packed = F.policy == 0 and os.environ.get("SCB_PACKED", "1") != "0"
        gk = g.clone().allow_packed(packed)
        d_out = torch.empty([K + 1, F.n], dtype=torch.int64, device="cuda")
        r = 123456 % p
        times = []
        for i in range(3 + max(args.steps, 5)):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            nxt = gk.fix_and_round_evals_device(r, d_out.data_ptr())
            b.record()
            torch.cuda.synchronize()
            if i >= 3:
                times.append(a.elapsed_time(b))
            del nxt
        kms = sum(times) / len(times)
        survey_bytes = 1.5 * K * (1 << v) * E
        alg_bytes = K * (1 << v) * (E + (2 if packed else E / 2))
        proof_bytes = K * (1 << v) * (3.0 * E if packed else 4.0 * E)  # packed: 8+8+2+3*(1+1/2+..)=24 B per entry-column
        peak, peak_src = hbm_peak()
        achieved = alg_bytes / (kms * 1e-3) / 1e9
        kname = "k_fold_round_sp<3,in=u64,out=u32>" if packed else "k_fold_round<%s,3>" % {0: "PolSP", 1: "PolG1", 4: "PolGN<4>"}[F.policy]
        # dram__bytes_read.sum + dram__bytes_write.sum of this kernel at this size from the committed ncu --set full
        # capture (profiles/r01_ncu_summary_final.md: 6.442495 GB + 1.613603 GB per launch); null for other shapes
        traffic = 8.056098e9 if (packed and v == 28 and K == 3 and p == MODULUS) else None
        alone = {"kernel": kname + " (fused fold + round message), 2^%d-entry tables, timed alone" % v,
                 "achieved": achieved, "frac": achieved / peak, "kernel_ms": kms, "algorithmic_bytes_per_launch": alg_bytes,
                 "traffic": traffic, "survey_bytes_per_launch_unpacked": survey_bytes,
                 "frac_vs_survey_bytes": survey_bytes / (kms * 1e-3) / 1e9 / peak}
        proof = {"proof_bytes_moved": proof_bytes, "proof_frac_of_hbm_roofline": (proof_bytes / (ms * 1e-3) / 1e9) / peak,
                 "proof_survey_bytes": 4.0 * K * (1 << v) * E,
                 "proof_frac_vs_survey_bytes": (4.0 * K * (1 << v) * E / (ms * 1e-3) / 1e9) / peak}
        resident = world == 1 and res_stats["launches"] == args.steps and res_stats["last_rounds"] == v - 1
        if resident:
            # algorithmic bytes of rounds 1..v-1: round j reads K tables of 2^(v-j+1) entries and writes 2^(v-j)
            in_b, out_b = E, (4 if packed else E)
            rbytes = []
            for j in range(1, v):
                rbytes.append(K * ((1 << (v - j + 1)) * in_b + (1 << (v - j)) * out_b))
                in_b = out_b
            res_ms = res_stats["total_ms"] / res_stats["launches"]
            res_achieved = sum(rbytes) / (res_ms * 1e-3) / 1e9
            w0 = res_stats["work_us"][0]
            roof = {"bound": "hbm", "kernel": "k_persist_rounds<%s,3> (rounds 1..%d of the proof, fused fold + round message, one "
                                              "cooperative launch)" % ({0: "PolSP", 1: "PolG1", 4: "PolGN<4>"}[F.policy], v - 1),
                    "achieved": res_achieved, "peak": peak, "unit": "GB/s", "frac": res_achieved / peak, "traffic": None,
                    "traffic_note": "ncu serialises kernel and host, so the resident kernel cannot run under it; its passes are "
                                    "the per-round kernels' loop bodies, whose captured DRAM traffic equals the algorithmic bytes "
                                    "(round1_kernel_alone.traffic, profiles/)",
                    "kernel_ms": res_ms, "algorithmic_bytes_per_launch": sum(rbytes), "peak_source": peak_src,
                    "timing": "CUDA events around the launch on its stream, all %d launches of the timed region" % res_stats["launches"],
                    "share_of_step": res_ms / ms,
                    "round1_phase": {"us": w0, "achieved": rbytes[0] / (w0 * 1e-6) / 1e9, "frac": rbytes[0] / (w0 * 1e-6) / 1e9 / peak,
                                     "source": "%globaltimer stamps inside the kernel, last launch"},
                    "latency_us": {"host_turnaround_total": sum(res_stats["turn_us"]), "rounds_after_1_total": sum(res_stats["work_us"][1:])},
                    "round1_kernel_alone": alone}
        else:
            roof = {"bound": "hbm", "kernel": alone["kernel"], "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": traffic, "kernel_ms": kms, "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peak_src,
                    "survey_bytes_per_launch_unpacked": survey_bytes, "frac_vs_survey_bytes": alone["frac_vs_survey_bytes"]}
        roof.update(proof)

    # ---- e2e: the same proof through the C ABI with HOST tables (pinned), H2D inside the timed region
    e2e = None
    if not args.no_e2e:
        import numpy as np

        host = []
        for t in tabs:
            ht = torch.empty([1 << v, F.n], dtype=torch.int64, pin_memory=True)
            # device -> pinned host staging (outside the timed region)
            dt_ = torch.empty([1 << v, F.n], dtype=torch.int64, device="cuda")
            T._lib.check(T.lib.scb_mle_copy_to_device(t._h, dt_.data_ptr()))
            ht.copy_(dt_)
            del dt_
            host.append(ht.numpy().view(np.uint64))
        torch.cuda.synchronize()

        # One GPU: scb_poly_product_from_host -- for the small-prime field the tables cross PCIe as packed uint32 where
        # the host cores keep up and as 8-byte entries where not (narrowed on the device).  Sharded runs keep one plain
        # cudaMemcpy per table: the ranks of a box share its host memory system, which is what bounds the narrowing
        # upload (2 GPUs measured with it: 117.8 ms per step against 119.0 plain, profiles/r01_bench_2gpu_packed_upload.json;
        # SCB_BENCH_E2E_UPLOAD=1 selects it anyway).
        one_call = (world == 1 or os.environ.get("SCB_BENCH_E2E_UPLOAD", "0") == "1") and os.environ.get("SCB_BENCH_E2E_PLAIN", "0") == "0"

        def e2e_step():
            if one_call:
                gg = T.ProductMLE.from_host_tables(F, v, host)
            else:
                hs = [T.DenseMultilinearExtension.from_evaluations_vec(F, v, h) for h in host]  # cudaMemcpy H2D
                gg = T.ProductMLE.new(hs)
            return prove(gg)  # messages come back device -> host every round

        e2e_step()
        barrier()
        t0 = time.perf_counter()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(args.e2e_steps):
            e2e_step()
        b.record()
        barrier()
        wall = (time.perf_counter() - t0) / args.e2e_steps
        ems = max(a.elapsed_time(b) / args.e2e_steps, wall * 1e3)
        if world > 1:
            t = torch.tensor([ems], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ems = float(t.item())
        rounds = v + (world.bit_length() - 1)
        h2d = K * (1 << v) * E * n_gpus
        note = "scb_mle_from_host x%d (pinned host tables, H2D) + Prover::new + all rounds + Fiat-Shamir, per step" % K
        upload = None
        if one_call:
            import ctypes

            pc, rc_, hb = ctypes.c_uint64(), ctypes.c_uint64(), ctypes.c_uint64()
            T._lib.check(T.lib.scb_host_pack_stats(ctypes.byref(pc), ctypes.byref(rc_), ctypes.byref(hb)))
            note = "scb_poly_product_from_host (pinned host tables, H2D) + Prover::new + all rounds + Fiat-Shamir, per step"
            if F.policy == 0 and pc.value + rc_.value > 0 and os.environ.get("SCB_HOST_PACK", "1") != "0":
                h2d = hb.value  # bytes of the copies the last step queued, counted by the library as it queued them
                if world > 1:
                    tb = torch.tensor([h2d], device="cuda", dtype=torch.int64)
                    dist.all_reduce(tb, op=dist.ReduceOp.SUM)
                    h2d = int(tb.item())
                wire21 = p < (1 << 21) and os.environ.get("SCB_HOST_PACK_WIRE", "21") == "21"
                upload = {"chunk_entries": K * (1 << v) // (pc.value + rc_.value), "chunks_narrowed_on_host": pc.value,
                          "chunks_narrowed_on_device": rc_.value,
                          "host_lane_wire_format": "three 21-bit entries per 64-bit word" if wire21 else "uint32",
                          "host_pack_threads_per_rank": min(32, int(os.environ.get("SCB_HOST_PACK_THREADS", max(1, (os.cpu_count() or 1) // world))))}
        e2e = {"value": total_entries / (ems * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": ems,
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": rounds * (K + 1) * E * n_gpus, "note": note}
        if upload:
            e2e["upload"] = upload
        del host

    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        from oracle.coracle import max_threads

        cv = args.cpu_vars
        v1, t1 = cpu_prove_sample(p, cv, 1)
        cores = max_threads()
        vN, tN = cpu_prove_sample(p, cv, cores)
        cpu = {"value": vN, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"oracle/oracle.c orc_product_prove on 2^{cv}-entry tables (bounded sample), {cores} OpenMP threads; "
                         f"single-thread (reference-faithful) = {v1:.2f} {UNIT}",
               "single_thread_value": v1, "seconds": tN}

    if rank == 0:
        sys.stdout.flush()
        os.write(json_fd, (json.dumps({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": workload_name(v, n_gpus, p), "tables": K, "vars_per_gpu": v, "total_vars": v + (world.bit_length() - 1),
                       "field_modulus": p, "bytes_per_element": E, "arith_policy": {0: "small-prime 32-bit", 1: "generic 64-bit", 4: "4-limb"}[F.policy],
                       "l2": "inputs (%.1f GB per GPU) larger than L2; no flush needed" % (K * (1 << v) * E / 1e9),
                       "parallelism": f"tables sharded by top variables over {n_gpus} GPU(s)"
                                      + ("" if world == 1 else f"; per-round exchange: {exchange}")},
            "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
        }) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
