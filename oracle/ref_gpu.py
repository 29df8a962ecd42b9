"""The reference's own GPU kernel on this box.  TEST / MEASUREMENT INFRASTRUCTURE ONLY (like the rest of oracle/).

`cuda_kmer_t1ha2` (/root/reference/src/cuda_kernel.cu:250-321) is built by `make -C oracle ref_gpu` from the source
where it lies - nothing is copied - into
    oracle/_ref/cuda_kmer_hash.ptx          build.rs:31-39's recipe with the newest feature the reference has
                                            (cuda-sketch-hopper: -ptx -arch=compute_90 -code=sm_90); JIT-compiled
                                            forward onto the B200 by the driver, as cudarc's load_ptx would do it
    oracle/_ref/cuda_kernel_sm100a.cubin    the same file compiled natively for sm_100a
and launched here exactly as src/sketch_cuda.rs:119-166 launches it: 512 k-mer starts per thread (:130),
n_threads = ceil(n_kmers / 512) (:131), n_hash_per_thread = max(512 / scaled * 4, 8) slots per thread (:136), a
zero-filled u64 slot array (:138), LaunchConfig::for_num_elems = 1024-thread blocks (:141, cudarc 0.10), the whole
slot array copied back (:156) and every non-zero slot inserted into a set on the host (:158-163).

Only bench.py's `reference_gpu_kernel` leg and tests call this; the product never does.
"""
from __future__ import annotations

import ctypes
import os
import time

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PTX = os.path.join(_HERE, "_ref", "cuda_kmer_hash.ptx")
CUBIN = os.path.join(_HERE, "_ref", "cuda_kernel_sm100a.cubin")
KMER_PER_THREAD = 512


def available() -> bool:
    return os.path.exists(PTX) or os.path.exists(CUBIN)


class RefKernel:
    """cuda_kmer_t1ha2 loaded into torch's CUDA context (driver API through cuda-python)."""

    def __init__(self, image: str = "cubin"):
        import torch
        from cuda.bindings import driver as drv
        torch.cuda.current_stream().synchronize()  # torch has created / bound the primary context
        self.drv = drv
        path = CUBIN if image == "cubin" else PTX
        with open(path, "rb") as f:
            data = f.read() + b"\0"
        err, self.mod = drv.cuModuleLoadData(data)
        if int(err) != 0:
            raise RuntimeError("cuModuleLoadData(%s) failed: %s" % (os.path.basename(path), err))
        err, self.fn = drv.cuModuleGetFunction(self.mod, b"cuda_kmer_t1ha2")
        if int(err) != 0:
            raise RuntimeError("cuModuleGetFunction failed: %s" % err)
        self.image = os.path.basename(path)

    @staticmethod
    def geometry(n_bps: int, k: int, scaled: int):
        n_kmers = n_bps - k + 1
        n_threads = (n_kmers + KMER_PER_THREAD - 1) // KMER_PER_THREAD
        n_hash_per_thread = max(KMER_PER_THREAD // scaled * 4, 8)
        return n_threads, n_hash_per_thread

    def launch(self, d_seq: int, n_bps: int, d_slots: int, k: int, scaled: int, seed: int, canonical: bool, stream: int):
        """one genome; d_slots must be zero-filled u64[n_threads * n_hash_per_thread]"""
        n_threads, nh = self.geometry(n_bps, k, scaled)
        args = ((d_seq, n_bps, KMER_PER_THREAD, nh, k, (2 ** 64 - 1) // scaled, seed, bool(canonical), d_slots),
                (ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_uint64,
                 ctypes.c_uint64, ctypes.c_bool, ctypes.c_void_p))
        err, = self.drv.cuLaunchKernel(self.fn, (n_threads + 1023) // 1024, 1, 1, 1024, 1, 1, 0, self.drv.CUstream(stream), args, 0)
        if int(err) != 0:
            raise RuntimeError("cuLaunchKernel(cuda_kmer_t1ha2) failed: %s" % err)


def time_reference_kernel(seq_dev, seq_host_pinned, genome_len: int, n_genomes: int, k: int, scaled: int, seed: int,
                          image: str = "cubin"):
    """The reference's GPU hashing of `n_genomes` genomes of the bench batch, one launch per genome on one stream
    (all rayon workers share cudarc's single stream).
      kernel_ms_per_genome   launches back to back on sequences already in HBM, slot arrays pre-zeroed
      e2e_ms_per_genome      per genome: H2D of the sequence from pinned memory, zero the slots, launch, D2H of the whole
                             slot array (src/sketch_cuda.rs:134-156) - without the host-side set build
    Returns (dict, list of per-genome sorted unique hash arrays of the first 4 genomes)."""
    import torch
    rk = RefKernel(image)
    dev = seq_dev.device
    n_threads, nh = rk.geometry(genome_len, k, scaled)
    slots = torch.zeros((n_genomes, n_threads * nh), dtype=torch.int64, device=dev)
    st = torch.cuda.current_stream()
    # warm-up (module load, JIT)
    rk.launch(seq_dev.data_ptr(), genome_len, slots[0].data_ptr(), k, scaled, seed, True, st.cuda_stream)
    st.synchronize()
    slots.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for g in range(n_genomes):
        rk.launch(seq_dev.data_ptr() + g * genome_len, genome_len, slots[g].data_ptr(), k, scaled, seed, True, st.cuda_stream)
    e1.record()
    e1.synchronize()
    kernel_ms = e0.elapsed_time(e1) / n_genomes
    sets = []
    for g in range(min(4, n_genomes)):
        h = slots[g].cpu().numpy().view(np.uint64)
        sets.append(np.unique(h[h != 0]))
    # end to end per file, as extract_kmer_t1ha2_cuda does it
    d_one = torch.empty(genome_len, dtype=torch.uint8, device=dev)
    h_slots = torch.empty(n_threads * nh, dtype=torch.int64, pin_memory=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for g in range(n_genomes):
        d_one.copy_(seq_host_pinned[g * genome_len:(g + 1) * genome_len], non_blocking=True)
        slots[0].zero_()
        rk.launch(d_one.data_ptr(), genome_len, slots[0].data_ptr(), k, scaled, seed, True, st.cuda_stream)
        h_slots.copy_(slots[0], non_blocking=True)
        st.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / n_genomes
    return dict(image=rk.image, genomes=n_genomes, kernel_ms_per_genome=kernel_ms, kernel_genomes_per_s=1e3 / kernel_ms,
                e2e_ms_per_genome=e2e_ms, e2e_genomes_per_s=1e3 / e2e_ms, threads=n_threads, slots_per_thread=nh,
                ctas=(n_threads + 1023) // 1024,
                launch="512 k-mers per thread, 1024-thread blocks, %d slots per thread (src/sketch_cuda.rs:130-141)" % nh), sets
