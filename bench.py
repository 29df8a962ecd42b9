#!/usr/bin/env python
"""bench.py — BASELINE.json's metric on BASELINE.json's configs, on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Headline (`metric`/`value`): genomes sketched per second, config[1] of BASELINE.json —
1,000 synthetic 5 Mbp genomes, k=21 scaled=1500 D=4096 per GPU (weak scaling: every rank
sketches its own 1,000 genomes, no collective on the data path).  A "step" is one pass of
the sketch hot path (k-mer hash -> set -> HV encode -> norm -> quantise -> bit-pack) over the
whole batch with the sequence bytes already resident in HBM.  `e2e` is the same pass through
the host-pointer C-ABI call (pinned host buffers, H2D of the FASTA bytes and D2H of the
sketches inside the timed region).  The `dist` object carries the second half of the
metric: ANI pairs/s for the all-vs-all over 10,000 sketches (config[2]), ref rows sharded
over the ranks, queries broadcast and hits gathered with NCCL.

`--impl reference` times the reference's CPU path restated in C (oracle/hg_oracle.c — the
Rust crate cannot be built here: no cargo/rustc) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

K, SCALED, SEED, HV_D = 21, 1500, 123, 4096
GENOME_LEN = 5_000_000
# algorithmic bytes of the k-mer hash kernel per genome: 1 B per base read + 8 B per sampled
# hash written — n ~= L / scaled
ALG_BYTES_PER_GENOME = GENOME_LEN + 8 * (GENOME_LEN // SCALED)
# integer instructions the k-mer kernel executes per k-mer (ncu smsp__inst_executed /
# k-mers, profiles/): used only for the auxiliary INT32 roofline
KMER_INST_PER_KMER = float(os.environ.get("HG_KMER_INST_PER_KMER", "127"))


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm_gbs=float(d["hbm_gbs"]), bf16_tflops=float(d["bf16_tflops"]),
                    bf16_tflops_sustained=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), source="measured")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback")


class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region (NVML, 50 ms period)."""

    def __init__(self, index: int):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop = threading.Event()
        self._t = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self._nv = None
            return self
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def _run(self):
        nv = self._nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40,
                 "hw_power_brake_slowdown": 0x80, "sync_boost": 0x10, "applications_clocks_setting": 0x2}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.05)

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join()
        med = statistics.median(self.samples) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def cpu_sketch_baseline(seq_host: np.ndarray, n_avail: int, seconds_hint: float = 10.0):
    """The restated reference CPU path (oracle, `kind: port`) on a bounded sample, all cores."""
    import oracle as O  # the checker / baseline: never on the product path
    cores = os.cpu_count() or 1
    O.set_threads(cores)
    # ~0.15 s per 5 Mbp genome per core: size the sample for roughly seconds_hint of wall time
    n = int(min(n_avail, max(cores, cores * seconds_hint / 0.15)))
    off = np.arange(n + 1, dtype=np.uint64) * np.uint64(GENOME_LEN)
    t0 = time.perf_counter()
    O.sketch_batch(seq_host[: n * GENOME_LEN], off, k=K, scaled=SCALED, seed=SEED, hv_d=HV_D, want_hv=False)
    dt = time.perf_counter() - t0
    O.set_threads(1)
    t1 = time.perf_counter()
    O.sketch_batch(seq_host[: 4 * GENOME_LEN], off[:5], k=K, scaled=SCALED, seed=SEED, hv_d=HV_D, want_hv=False)
    per_core = 4 / (time.perf_counter() - t1)
    O.set_threads(cores)
    return dict(value=n / dt, unit="genomes/s", cores=cores, kind="port", value_1_thread=per_core,
                sample="%d of the step's 5 Mbp genomes, oracle/hg_oracle.c (C restatement of src/sketch.rs:35-52), %d threads, %.1f s"
                % (n, cores, dt))


def parity_sample(seq_host, packed, bits, norm, n):
    """Checker leg: the e2e results of a few genomes of the timed batch against the oracle."""
    import oracle as O
    pick = sorted(set([0, 1, n // 2, n - 1]))
    bad = 0
    for g in pick:
        s = seq_host[g * GENOME_LEN:(g + 1) * GENOME_LEN]
        w = O.sketch_batch(s, np.array([0, GENOME_LEN], np.uint64), k=K, scaled=SCALED, seed=SEED, hv_d=HV_D, want_hv=False)
        nb = int(w["quant_bits"][0]) * HV_D // 8
        ok = (int(w["quant_bits"][0]) == int(bits[g]) and int(w["norm2"][0]) == int(norm[g])
              and np.array_equal(w["packed"][0, :nb], packed[g, :nb]))
        bad += 0 if ok else 1
    if bad:
        raise RuntimeError("parity check failed on %d of %d sampled genomes" % (bad, len(pick)))
    return "ok: %d sampled genomes of the timed batch bit-identical to the oracle (packed HV, bits, norm)" % len(pick)


def run_reference(args, rank):
    """--impl reference: the reference's CPU algorithm on the host cores (rank 0 only)."""
    if rank != 0:
        return
    import torch
    from hypergen_b200 import synth
    cores = os.cpu_count() or 1
    import oracle as O
    O.set_threads(cores)
    n = int(min(args.genomes, max(cores, 2 * cores)))
    seq, off = synth.family_batch(n, GENOME_LEN, device="cpu")
    seq = seq.numpy()
    times = []
    for step in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        O.sketch_batch(seq, off, k=K, scaled=SCALED, seed=SEED, hv_d=HV_D, want_hv=False)
        if step >= args.warmup:
            times.append(time.perf_counter() - t0)
    total = sum(times)
    value = n * len(times) / total
    line = {
        "impl": "reference", "metric": "genomes sketched/sec (5 Mbp, k=21, D=4096)", "value": value, "unit": "genomes/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": "sketch %d synthetic 5 Mbp genomes per step (bounded sample of config[1]), k=21 scaled=1500 D=4096, CPU" % n},
        "cpu_baseline": {"value": value, "unit": "genomes/s", "cores": cores, "kind": "port",
                         "sample": "%d genomes per step; oracle/hg_oracle.c restates src/sketch.rs:35-52 (Rust crate not buildable here: no cargo)" % n},
        "e2e": {"value": value, "unit": "genomes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _emit(line)


_JSON_OUT = None


def _claim_stdout():
    """Keep stdout for the one JSON line: everything else that writes to fd 1 from here on (NCCL's
    version banner, library chatter) goes to stderr."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def _emit(line: dict):
    _JSON_OUT.write(json.dumps(line) + "\n")
    _JSON_OUT.flush()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--genomes", type=int, default=1000, help="genomes per GPU per step (config[1]: 1000)")
    ap.add_argument("--dist-n", type=int, default=10000, help="sketches in the all-vs-all (config[2]: 10000)")
    ap.add_argument("--no-dist", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-fasta", action="store_true", help="skip the raw-FASTA end-to-end leg")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus > 1 and world == 1 and "RANK" not in os.environ:
        # convenience: relaunch under torchrun exactly as the driver does
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", "29577", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))

    _claim_stdout()
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    import hypergen_b200 as hg
    from hypergen_b200 import multigpu, synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = multigpu.bind_to_gpu_numa(local_rank) if os.environ.get("HG_NUMA_BIND", "1") != "0" else {"bound": False}
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    peaks = load_peaks()
    ctx = hg.Context(local_rank)
    ext = torch.cuda.ExternalStream(ctx.stream, device=dev)
    params = hg.make_params(k=K, scaled=SCALED, seed=SEED, canonical=True, hv_d=HV_D)
    n = args.genomes

    # ---------------- workload: this rank's genomes, generated on the GPU ----------------
    seq_dev, seg_off = synth.family_batch(n, GENOME_LEN, device=dev, first=rank * n)
    d_packed = torch.empty((n, 2 * HV_D), dtype=torch.uint8, device=dev)
    d_bits = torch.empty(n, dtype=torch.uint8, device=dev)
    d_norm = torch.empty(n, dtype=torch.int32, device=dev)
    d_nh = torch.empty(n, dtype=torch.int32, device=dev)
    d_hv = torch.empty((n, HV_D), dtype=torch.int16, device=dev)
    torch.cuda.synchronize()

    def sketch_step():
        ctx.sketch_batch_dev(seq_dev.data_ptr(), seg_off, params, d_hv.data_ptr(), d_packed.data_ptr(),
                             d_bits.data_ptr(), d_norm.data_ptr(), d_nh.data_ptr())

    # ---------------- device-resident timing (`value`) ----------------
    ctx.set_profiling(True)
    for _ in range(args.warmup):
        sketch_step()
    ctx.sync()
    ctx.sketch_status()
    sampler = ClockSampler(local_rank).start()
    barrier()
    l0 = ctx.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stage = []
    with torch.cuda.stream(ext):
        e0.record()
        for _ in range(args.steps):
            sketch_step()
            stage.append(ctx.stage_ms())  # syncs the stream: < 20 us against a ~30 ms step
        e1.record()
    e1.synchronize()
    barrier()
    clocks = sampler.stop()
    launches = ctx.launches - l0
    dev_ms = max_over_ranks(e0.elapsed_time(e1))
    ctx.sketch_status()
    value = world * n * args.steps / (dev_ms * 1e-3)
    kmer_ms = statistics.mean(s[1] for s in stage)
    enc_ms = statistics.mean(s[2] for s in stage)
    stg_ms = statistics.mean(s[0] for s in stage)
    achieved_gbs = n * ALG_BYTES_PER_GENOME / (kmer_ms * 1e-3) / 1e9
    kmers_per_s = n * (GENOME_LEN - K + 1) / (kmer_ms * 1e-3)

    # ---------------- end to end through the host-pointer C ABI ----------------
    seq_host = torch.empty(n * GENOME_LEN, dtype=torch.uint8, pin_memory=True)
    seq_host.copy_(seq_dev)
    h_packed = torch.empty((n, 2 * HV_D), dtype=torch.uint8, pin_memory=True)
    h_bits = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    h_norm = torch.empty(n, dtype=torch.int32, pin_memory=True)
    h_nh = torch.empty(n, dtype=torch.int32, pin_memory=True)
    torch.cuda.synchronize()
    lib = hg.ffi.load()
    import ctypes as C

    def e2e_step():
        rc = lib.hg_sketch_batch(ctx._h, seq_host.data_ptr(), seg_off.ctypes.data, n, C.byref(params), None,
                                 h_packed.data_ptr(), h_bits.data_ptr(), h_norm.data_ptr(), h_nh.data_ptr())
        if rc != 0:
            raise RuntimeError(lib.hg_last_error().decode())

    # H2D of the FASTA bytes alone (north star: reported separately)
    h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    scratch_dev = torch.empty_like(seq_dev)
    scratch_dev.copy_(seq_host, non_blocking=True)
    torch.cuda.synchronize()
    h0.record()
    for _ in range(3):
        scratch_dev.copy_(seq_host, non_blocking=True)
    h1.record()
    h1.synchronize()
    h2d_only_gbs = 3 * n * GENOME_LEN / (h0.elapsed_time(h1) * 1e-3) / 1e9
    del scratch_dev

    e2e_steps = max(3, args.steps // 2)
    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()  # synchronous: returns after the D2H of the sketches
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    barrier()
    e2e_value = world * n * e2e_steps / e2e_s
    h2d = int(n * GENOME_LEN + (n + 1) * 8 + n * 40)
    d2h = int(n * 2 * HV_D + n * 9)
    # sanity: the e2e results equal the device-resident ones
    assert torch.equal(h_norm, d_norm.cpu()) and torch.equal(h_bits, d_bits.cpu())

    # ---------------- the same, from raw FASTA file bytes (hg_sketch_fasta_batch) ----------------
    # files = ">g\n" + 80-column lines; the library merges the records on the GPU (fastx_reader::read_merge_seq)
    fasta = None
    if not args.no_fasta:
        hdr = torch.tensor(list(b">g\n"), dtype=torch.uint8, device=dev).expand(n, 3)
        body = torch.cat([seq_dev.view(n, GENOME_LEN // 80, 80),
                          torch.full((n, GENOME_LEN // 80, 1), 10, dtype=torch.uint8, device=dev)], dim=2).view(n, -1)
        raw_dev = torch.cat([hdr, body], dim=1).contiguous()
        file_len = raw_dev.shape[1]
        raw_host = torch.empty(n * file_len, dtype=torch.uint8, pin_memory=True)
        raw_host.copy_(raw_dev.view(-1))
        del raw_dev, body, hdr
        file_off = (np.arange(n + 1, dtype=np.uint64) * np.uint64(file_len))

        def fasta_step():
            rc = lib.hg_sketch_fasta_batch(ctx._h, raw_host.data_ptr(), file_off.ctypes.data, n, C.byref(params), None,
                                           h_packed.data_ptr(), h_bits.data_ptr(), h_norm.data_ptr(), h_nh.data_ptr())
            if rc != 0:
                raise RuntimeError(lib.hg_last_error().decode())

        for _ in range(2):
            fasta_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            fasta_step()
        torch.cuda.synchronize()
        fa_s = max_over_ranks(time.perf_counter() - t0)
        barrier()
        assert torch.equal(h_norm, d_norm.cpu()) and torch.equal(h_bits, d_bits.cpu())
        fasta = {"value": world * n * e2e_steps / fa_s, "unit": "genomes/s", "ms_per_step": 1e3 * fa_s / e2e_steps,
                 "h2d_bytes_per_step": int(n * file_len + (n + 1) * 8 + n * 40), "d2h_bytes_per_step": d2h,
                 "h2d_gbs": world * n * file_len * e2e_steps / fa_s / 1e9,
                 "input": "%d FASTA files of %d bytes (one record, 80-column lines) in pinned host memory; "
                          "records merged on the GPU" % (n, file_len)}
        del raw_host

    # DRAM bytes of one kmer_hash launch, from the committed `ncu --set full` capture (bytes, per launch like `achieved`)
    traffic, traffic_src = None, None
    try:
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "kmer_traffic.json")) as f:
            tj = json.load(f)
        traffic = tj["dram_bytes_per_launch"] * (n / tj["genomes_per_launch"])
        traffic_src = tj["source"] + ("" if n == tj["genomes_per_launch"] else " (scaled per genome)")
    except (OSError, KeyError, ValueError):
        pass

    line = {
        "metric": "genomes sketched/sec (5 Mbp, k=21, D=4096)", "value": value, "unit": "genomes/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": "sketch %d synthetic 5 Mbp genomes per GPU per step (BASELINE configs[1]), k=21 scaled=1500 D=4096"
                               % n, "genomes_per_gpu": n, "genome_bp": GENOME_LEN, "k": K, "scaled": SCALED, "hv_d": HV_D,
                   "l2": "inputs (%.1f GB per step) larger than L2" % (n * GENOME_LEN / 1e9), "parallelism": "genomes sharded, no collective"},
        "e2e": {"value": e2e_value, "unit": "genomes/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps, "ms_per_step": 1e3 * e2e_s / e2e_steps,
                "h2d_gbs": world * h2d * e2e_steps / e2e_s / 1e9, "h2d_only_gbs_per_gpu": h2d_only_gbs,
                "h2d_only_genomes_per_s_per_gpu": h2d_only_gbs * 1e9 / GENOME_LEN, "from_fasta_files": fasta},
        "numa": numa,
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": "kmer_hash_kernel<21,true>", "achieved": achieved_gbs, "peak": peaks["hbm_gbs"],
                     "unit": "GB/s", "frac": achieved_gbs / peaks["hbm_gbs"], "traffic": traffic, "traffic_source": traffic_src,
                     "peak_source": peaks["source"],
                     "alg_bytes_per_launch": n * ALG_BYTES_PER_GENOME, "ms_per_launch": kmer_ms,
                     "note": "bit-exact t1ha2 makes this kernel INT32-issue bound, not HBM bound (SURVEY.md 8d): see roofline_int32"},
        "stages_ms": {"staging_memset": stg_ms, "kmer_hash": kmer_ms, "encode_pack": enc_ms},
    }

    # auxiliary INT32 roofline (denominator measured live with the library's probe kernels)
    try:
        int_peak = ctx.int_peak(2)
        line["roofline_int32"] = {"bound": "int32_issue", "achieved": kmers_per_s * KMER_INST_PER_KMER * 1.0,
                                  "peak": int_peak, "unit": "lane-instr/s", "frac": kmers_per_s * KMER_INST_PER_KMER / int_peak,
                                  "kmers_per_s": kmers_per_s, "inst_per_kmer": KMER_INST_PER_KMER,
                                  "peak_imad_only": ctx.int_peak(0), "peak_alu_only": ctx.int_peak(1)}
    except Exception as ex:  # pragma: no cover
        line["roofline_int32"] = {"error": str(ex)}

    # ---------------- dist: all-vs-all ANI over dist_n sketches ----------------
    if not args.no_dist:
        line["dist"] = bench_dist(args, ctx, ext, hg, multigpu, synth, dev, rank, world, barrier, max_over_ranks, peaks)

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_sketch_baseline(seq_host.numpy(), n)
        line["parity_check"] = parity_sample(seq_host.numpy(), h_packed.numpy(), h_bits.numpy(), h_norm.numpy(), n)
    elif rank == 0:
        line["cpu_baseline"] = None

    if rank == 0:
        _emit(line)
    torch.cuda.synchronize()
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def bench_dist(args, ctx, ext, hg, multigpu, synth, dev, rank, world, barrier, max_over_ranks, peaks):
    import torch
    import torch.distributed as dist
    nq = args.dist_n
    D = HV_D
    # sketches with controlled Jaccard, encoded on the GPU from hash sets (rank 0), then broadcast
    if rank == 0:
        sets = synth.hash_sets_family(nq)
        off = np.zeros(nq + 1, np.uint64)
        off[1:] = np.cumsum([len(s) for s in sets])
        hashes = torch.from_numpy(np.concatenate(sets).view(np.int64)).to(dev)
        hv = torch.empty((nq, D), dtype=torch.int16, device=dev)
        bits = torch.empty(nq, dtype=torch.uint8, device=dev)
        norm = torch.empty(nq, dtype=torch.int32, device=dev)
        ctx.encode_sets_dev(hashes.data_ptr(), off, D, hv.data_ptr(), None, bits.data_ptr(), norm.data_ptr())
        ctx.sync()
        del hashes
    else:
        hv = norm = None
    cap = 4_000_000
    # counter and hit array live in one buffer laid out as the multi-GPU gather sends it (no staging copy)
    hit_blk, d_cnt, d_hits = multigpu.hit_block(cap, dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    hits_pin = torch.empty(cap * 16, dtype=torch.uint8, pin_memory=True)
    bounds = multigpu.triangle_rows(nq, world, align=128)
    a, b = bounds[rank], bounds[rank + 1]
    n_pairs = nq * (nq - 1) // 2

    # multi-GPU: one fused broadcast buffer (HVs + norms) and one fixed-size gather per step
    if world > 1:
        qbuf = torch.empty(nq * D * 2 + nq * 4, dtype=torch.uint8, device=dev)
        if rank == 0:
            qbuf[: nq * D * 2] = hv.view(torch.uint8).view(-1)
            qbuf[nq * D * 2:] = norm.view(torch.uint8).view(-1)
        gather_cap = 1 << 18  # hits per rank carried by the one-shot gather (4 MiB per rank)
        gather_pin = torch.empty(cap * 16, dtype=torch.uint8, pin_memory=True)
        gather_recv = torch.empty(world * (16 + gather_cap * 16), dtype=torch.uint8, device=dev)

    def step():
        """broadcast queries -> local shard -> hits gathered on rank 0; returns device ms"""
        nonlocal hv, norm
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        with torch.cuda.stream(ext):  # collectives and kernels ordered on the context's stream
            e0.record()
            dbg = os.environ.get("HG_BENCH_DEBUG") and rank == 0
            if dbg:
                torch.cuda.synchronize(); tA = time.perf_counter()
            if world > 1:
                hv, norm = multigpu.broadcast_queries_fused(qbuf, nq, D)
            if dbg:
                torch.cuda.synchronize(); tB = time.perf_counter()
            ctx.dist_dev(hv[a:b].data_ptr(), norm[a:b].data_ptr(), b - a, a, hv.data_ptr(), norm.data_ptr(), nq, 0, D, K,
                         85.0, True, path_sel[0], d_hits.data_ptr(), cap, d_cnt.data_ptr())
            if dbg:
                torch.cuda.synchronize(); ctx.sync(); tC = time.perf_counter()
            if world > 1:
                allh, overflow = multigpu.gather_hits_fixed(None, None, gather_cap, host_buf=gather_pin, block=hit_blk,
                                                            recv=gather_recv)
                if overflow:  # some shard produced more hits than the one-shot block carries
                    cnt = int(d_cnt.item())
                    allh = multigpu.gather_hits(d_hits, dev, count=min(cnt, cap))
            else:
                cnt = int(d_cnt.item())  # D2H of the hit count
                if cnt > cap:
                    raise RuntimeError("hit buffer too small: %d > %d" % (cnt, cap))
                hits_pin[: cnt * 16].copy_(d_hits[: cnt * 16], non_blocking=True)  # D2H of the hit list
                torch.cuda.current_stream().synchronize()
                allh = hits_pin[: cnt * 16].numpy().view(hg.ffi.HIT_DTYPE)
            if dbg:
                torch.cuda.synchronize(); tD = time.perf_counter()
                print("[dist step] bcast %.3f ms, dist_dev %.3f ms, gather %.3f ms" % ((tB - tA) * 1e3, (tC - tB) * 1e3, (tD - tC) * 1e3), file=sys.stderr)
            e1.record()
        e1.synchronize()
        return e0.elapsed_time(e1), ctx.stage_ms()[3], (allh.size if allh is not None else 0)

    ctx.set_profiling(True)
    # first call: path 0 = auto (single-plane tensor kernel if the rows are narrow, else two-limb tensor
    # kernel, else SIMT; records why); the timed steps pass the chosen path directly
    path_sel = [0]
    step()
    path_sel[0], path_reason = ctx.dist_last_path, ctx.dist_last_reason
    for _ in range(3):
        step()
    steps = max(5, args.steps)
    tot_ms, kern_ms, n_hits = 0.0, [], 0
    barrier()
    for _ in range(steps):
        flush.fill_(1)  # inputs (82 MB) fit in L2: flush it between timed iterations
        barrier()
        ms, kms, n_hits = step()
        tot_ms += max_over_ranks(ms)
        kern_ms.append(max_over_ranks(kms))
    kms = statistics.mean(kern_ms)
    out = {
        "metric": "ANI pairs/sec (all-vs-all, D=4096, ani_th=85)", "unit": "pairs/s",
        "value": n_pairs * steps / (tot_ms * 1e-3), "kernel_value": n_pairs / (kms * 1e-3),
        "ms_per_step": tot_ms / steps, "kernel_ms": kms, "steps": steps, "hits": int(n_hits), "pairs": n_pairs,
        "config": {"workload": "all-vs-all dist over %d synthetic sketches (BASELINE configs[2]), D=4096, ani_th=85" % nq,
                   "rows_per_rank": [bounds[r + 1] - bounds[r] for r in range(world)], "l2": "flushed between iterations"},
        "path": path_sel[0], "path_reason": path_reason,
    }
    alg_ops = 2.0 * D * n_pairs
    int8_peak = 2.0 * peaks["bf16_tflops"] * world  # kind::i8 runs at twice the bf16 MMA rate; all ranks' tensor pipes
    mac_mult = 1.0 if path_sel[0] == 3 else 4.0  # MMAs executed per algorithmic MAC
    note = ("algorithmic 2*D ops per pair; peak = n_gpus x 2 x measured bf16 (kind::i8 runs at twice the bf16 MMA rate); "
            + ("single s8 plane (x = 2a + s): executed MACs = algorithmic MACs; kernel_ms includes the i16 -> s8 pre-pass and its host sync"
               if path_sel[0] == 3 else "the two-limb split executes 4x these MACs, so frac tops out at 0.25"))
    out["roofline"] = {"bound": "tensor", "achieved": alg_ops / (kms * 1e-3) / 1e12, "peak": int8_peak, "unit": "TOP/s",
                       "frac": alg_ops / (kms * 1e-3) / 1e12 / int8_peak, "note": note,
                       "executed_frac": mac_mult * alg_ops / (kms * 1e-3) / 1e12 / int8_peak,
                       # DRAM bytes of one config-3 step from the ncu captures (profiles/r1_dist_narrow_full.md,
                       # r1_narrow_prep_full.md): pre-pass 82 MB read + 41 MB written, kernel 41.3 MB read + 0.6 MB written
                       "traffic": (165.0e6 if (path_sel[0] == 3 and nq == 10000 and world == 1) else None)}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # CPU port of dist::compute_hv_ani (dist.rs:231-294) on a bounded sample of the same sketches
        import oracle as O
        cores = os.cpu_count() or 1
        O.set_threads(cores)
        m = min(nq, 10000)
        hv_s, norm_s = hv[:m].cpu().numpy(), norm[:m].cpu().numpy()
        t0 = time.perf_counter()
        ani_s, _ = O.dist_all(hv_s, norm_s, hv_s, norm_s, k=K, symmetric=True, want_dot=False)
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": (m * (m - 1) // 2) / dt, "unit": "pairs/s", "cores": cores, "kind": "port",
                               "sample": "all-vs-all over the first %d of the %d sketches (%d pairs), oracle/hg_oracle.c, %d threads, %.1f s"
                               % (m, nq, m * (m - 1) // 2, cores, dt)}
    # e2e through the host-pointer C-ABI call (N=1 only): pinned host matrices in, hits out
    if world == 1:
        import ctypes as C
        lib = hg.ffi.load()
        hv_h = torch.empty((nq, D), dtype=torch.int16, pin_memory=True)
        norm_h = torch.empty(nq, dtype=torch.int32, pin_memory=True)
        hv_h.copy_(hv)
        norm_h.copy_(norm)
        hits_h = torch.empty(cap * 16, dtype=torch.uint8, pin_memory=True)
        torch.cuda.synchronize()
        n_hits_c = C.c_uint64(0)

        def e2e():
            rc = lib.hg_dist(ctx._h, hv_h.data_ptr(), norm_h.data_ptr(), nq, hv_h.data_ptr(), norm_h.data_ptr(), nq, D, K,
                             85.0, 1, path_sel[0], hits_h.data_ptr(), cap, C.byref(n_hits_c))
            if rc != 0:
                raise RuntimeError(lib.hg_last_error().decode())

        e2e()
        t0 = time.perf_counter()
        reps = 5
        for _ in range(reps):
            e2e()
        dt = (time.perf_counter() - t0) / reps
        assert n_hits_c.value == n_hits
        out["e2e"] = {"value": n_pairs / dt, "unit": "pairs/s", "h2d_bytes_per_step": int(hv_h.numel() * 2 + norm_h.numel() * 4),
                      "d2h_bytes_per_step": int(n_hits_c.value * 16 + 8), "ms_per_step": dt * 1e3}
        # the same call with the output stage on the GPU (hits come back in the reference's TSV order)
        milli_h = torch.empty(cap, dtype=torch.int32, pin_memory=True)

        def e2e_sorted():
            rc = lib.hg_dist_sorted(ctx._h, hv_h.data_ptr(), norm_h.data_ptr(), nq, hv_h.data_ptr(), norm_h.data_ptr(), nq, D,
                                    K, 85.0, 1, path_sel[0], hits_h.data_ptr(), milli_h.data_ptr(), cap, C.byref(n_hits_c))
            if rc != 0:
                raise RuntimeError(lib.hg_last_error().decode())

        e2e_sorted()
        t0 = time.perf_counter()
        for _ in range(reps):
            e2e_sorted()
        dts = (time.perf_counter() - t0) / reps
        hs = hits_h[: n_hits_c.value * 16].numpy().view(hg.ffi.HIT_DTYPE)
        ok = bool(np.all((hs["ani"][:-1] > hs["ani"][1:]) | ((hs["ani"][:-1] == hs["ani"][1:]) & (
            (hs["i"][:-1] > hs["i"][1:]) | ((hs["i"][:-1] == hs["i"][1:]) & (hs["j"][:-1] > hs["j"][1:]))))))
        out["e2e_sorted"] = {"value": n_pairs / dts, "unit": "pairs/s", "ms_per_step": dts * 1e3, "sort_ms": (dts - dt) * 1e3,
                             "d2h_bytes_per_step": int(n_hits_c.value * 20 + 8), "order_verified": ok,
                             "note": "hg_dist_sorted: hits radix-sorted on the GPU into dump_ani_file's order (utils.rs:262-285)"}
        # the `hyper-gen dist` flow: sketch-file payload (bit-packed rows) in, sorted hits out (hg_dist_packed)
        bits_np = bits.cpu().numpy()
        width = int(bits_np.max()) * D // 8
        packed_d = torch.empty((nq, 2 * D), dtype=torch.uint8, device=dev)
        hv2 = torch.empty_like(hv)
        hashes2 = torch.from_numpy(np.concatenate(sets).view(np.int64)).to(dev)
        ctx.encode_sets_dev(hashes2.data_ptr(), off, D, hv2.data_ptr(), packed_d.data_ptr(), bits.data_ptr(), norm.data_ptr())
        ctx.sync()
        del hashes2, hv2
        packed_h = torch.empty((nq, 2 * D), dtype=torch.uint8, pin_memory=True)
        packed_h.copy_(packed_d)
        bits_h = torch.empty(nq, dtype=torch.uint8, pin_memory=True)
        bits_h.copy_(bits)
        torch.cuda.synchronize()

        def e2e_packed():
            rc = lib.hg_dist_packed(ctx._h, packed_h.data_ptr(), 2 * D, bits_h.data_ptr(), norm_h.data_ptr(), nq,
                                    packed_h.data_ptr(), 2 * D, bits_h.data_ptr(), norm_h.data_ptr(), nq, D, K, 85.0, 1, 0, 1,
                                    hits_h.data_ptr(), milli_h.data_ptr(), cap, C.byref(n_hits_c))
            if rc != 0:
                raise RuntimeError(lib.hg_last_error().decode())

        e2e_packed()
        t0 = time.perf_counter()
        for _ in range(reps):
            e2e_packed()
        dtp = (time.perf_counter() - t0) / reps
        hp = hits_h[: n_hits_c.value * 16].numpy().view(hg.ffi.HIT_DTYPE)
        out["e2e_packed_sorted"] = {"value": n_pairs / dtp, "unit": "pairs/s", "ms_per_step": dtp * 1e3,
                                    "h2d_bytes_per_step": int(nq * width + nq * 5), "d2h_bytes_per_step": int(n_hits_c.value * 20 + 8),
                                    "same_hits_as_sorted": bool(n_hits_c.value == hs.size and np.array_equal(hp, hs)),
                                    "note": "hg_dist_packed: %d-bit packed sketch rows over PCIe, decompress + dist + sort on the GPU" % int(bits_np.max())}
    return out


if __name__ == "__main__":
    main()
