"""Synthetic workloads of the shapes BASELINE.json names.

Genomes are i.i.d. uniform over ACGT from a splitmix64 stream (genome seed ``0xB200 + a``,
base = "ACGT"[x >> 62]); families are an ancestor plus members with i.i.d. substitutions at
a fixed per-member rate, so that ANI ~ 1 - rate and `dist` has pairs on both sides of the
threshold.  Pure integer arithmetic on int64 tensors: the same function yields bit-identical
bytes on CPU and on the GPU (the bench generates on the GPU, the tests on the CPU).
"""
from __future__ import annotations

import numpy as np
import torch

_GAMMA = 0x9E3779B97F4A7C15
_M1 = 0xBF58476D1CE4E5B9
_M2 = 0x94D049BB133111EB
FAMILY_RATES = (0.0, 0.001, 0.01, 0.02, 0.05, 0.08, 0.12, 0.2, 0.3, 0.5)


def _s64(x: int) -> int:
    x &= (1 << 64) - 1
    return x - (1 << 64) if x >= (1 << 63) else x


def _lsr(x: torch.Tensor, s: int) -> torch.Tensor:
    return (x >> s) & ((1 << (64 - s)) - 1)


def splitmix64_stream(seed: int, n: int, device="cpu") -> torch.Tensor:
    """x_i = mix(seed + (i+1) * GAMMA), i = 0..n-1, as int64 (two's complement of the u64)."""
    i = torch.arange(1, n + 1, dtype=torch.int64, device=device)
    z = i * _s64(_GAMMA) + _s64(seed)
    z = (z ^ _lsr(z, 30)) * _s64(_M1)
    z = (z ^ _lsr(z, 27)) * _s64(_M2)
    return z ^ _lsr(z, 31)


_ACGT = (65, 67, 71, 84)


def genome(seed: int, length: int, device="cpu") -> torch.Tensor:
    """uint8 ASCII genome of `length` bases."""
    x = splitmix64_stream(seed, length, device)
    lut = torch.tensor(_ACGT, dtype=torch.uint8, device=device)
    return lut[_lsr(x, 62)]


def mutate(anc: torch.Tensor, seed: int, rate: float) -> torch.Tensor:
    """Substitute each base independently with probability `rate` by a different base."""
    if rate <= 0.0:
        return anc.clone()
    n = anc.numel()
    u = splitmix64_stream(seed, n, anc.device)
    hit = _lsr(u, 40) < int(rate * (1 << 24))
    # code of the ancestor base: A0 C1 G2 T3
    a = anc.to(torch.int64)
    code = ((a >> 1) ^ (a >> 2)) & 3
    new = (code + 1 + (u & 0xFFFF) % 3) & 3
    lut = torch.tensor(_ACGT, dtype=torch.uint8, device=anc.device)
    return torch.where(hit, lut[new], anc)


def family_member(g: int, length: int, members: int = 10, device="cpu", base_seed: int = 0xB200) -> torch.Tensor:
    """Genome g of a family-structured collection: ancestor g // members, member g % members."""
    a, m = divmod(g, members)
    anc = genome(base_seed + a, length, device)
    return mutate(anc, (base_seed << 16) + g, FAMILY_RATES[m % len(FAMILY_RATES)])


def family_batch(n_genomes: int, length: int, device="cpu", first: int = 0, members: int = 10,
                 base_seed: int = 0xB200):
    """(seq uint8[n*length] on `device`, seg_off uint64[n+1]) for genomes first..first+n-1."""
    seq = torch.empty(n_genomes * length, dtype=torch.uint8, device=device)
    anc_id, anc = None, None
    for t in range(n_genomes):
        g = first + t
        a, m = divmod(g, members)
        if a != anc_id:
            anc_id, anc = a, genome(base_seed + a, length, device)
        seq[t * length:(t + 1) * length] = mutate(anc, (base_seed << 16) + g, FAMILY_RATES[m % len(FAMILY_RATES)])
    seg_off = (np.arange(n_genomes + 1, dtype=np.uint64) * np.uint64(length))
    return seq, seg_off


def hash_sets_family(n_sketches: int, n_per: int = 3333, members: int = 10, scaled: int = 1500,
                     seed: int = 0xD157):
    """Controlled-Jaccard hash sets for dist-only workloads: every
    family shares a pool; a member keeps each pool element with probability p_m and tops up
    with private elements.  Returns a list of sorted unique uint64 arrays (all < MAX/scaled)."""
    thr = (2 ** 64 - 1) // scaled
    rng = np.random.default_rng(seed)
    keep = (1.0, 0.98, 0.9, 0.8, 0.65, 0.5, 0.35, 0.2, 0.1, 0.02)
    out = []
    pool = None
    for s in range(n_sketches):
        f, m = divmod(s, members)
        if m == 0 or pool is None:
            pool = rng.integers(0, thr, n_per, dtype=np.uint64)
        mask = rng.random(n_per) < keep[m % len(keep)]
        mine = pool[mask]
        priv = rng.integers(0, thr, n_per - mine.size, dtype=np.uint64)
        out.append(np.unique(np.concatenate([mine, priv])))
    return out


def _mix64(z: torch.Tensor) -> torch.Tensor:
    """splitmix64 finaliser on int64 tensors (two's complement of the u64)."""
    z = (z ^ _lsr(z, 30)) * _s64(_M1)
    z = (z ^ _lsr(z, 27)) * _s64(_M2)
    return z ^ _lsr(z, 31)


_KEEP = (1.0, 0.98, 0.9, 0.8, 0.65, 0.5, 0.35, 0.2, 0.1, 0.02)


def hash_sets_family_dev(n_sketches: int, n_per: int = 3333, members: int = 10, scaled: int = 1500, seed: int = 0xD157,
                         device="cpu", first: int = 0, chunk: int = 4096):
    """Controlled-Jaccard hash sets like hash_sets_family, generated with tensor arithmetic so that large
    collections (config 4: 100,000 sketches, config 5: 20,000 x 10,000 hashes) come straight from the GPU and any
    block of sketches [first, first + n) can be generated on its own (every value is a pure function of the sketch /
    family index).  Every set has exactly n_per distinct values < u64::MAX / scaled.
    Returns (hashes int64[n * n_per] as the bit pattern of the u64 values, hash_off uint64[n + 1])."""
    thr = (2 ** 64 - 1) // scaled
    keep = torch.tensor([int(k * (1 << 24)) for k in _KEEP], dtype=torch.int64, device=device)
    e = torch.arange(n_per, dtype=torch.int64, device=device)[None, :]
    out = torch.empty((n_sketches, n_per), dtype=torch.int64, device=device)
    for c0 in range(0, n_sketches, chunk):
        m = min(chunk, n_sketches - c0)
        sid = torch.arange(first + c0, first + c0 + m, dtype=torch.int64, device=device)[:, None]
        fam, mem = sid // members, sid % members
        pool = _mix64((fam * n_per + e) * _s64(_GAMMA) + _s64(seed))
        priv = _mix64((sid * n_per + e) * _s64(_GAMMA) + _s64(seed ^ 0x5EED5EED5EED))
        u = _mix64((sid * n_per + e) * _s64(_GAMMA) + _s64(seed ^ 0x0123456789AB))
        take = _lsr(u, 40) < keep[mem % len(_KEEP)]
        v = torch.where(take, pool, priv)
        v = _lsr(v, 1) % thr
        vs, _ = torch.sort(v, dim=1)
        if bool((vs[:, 1:] == vs[:, :-1]).any()):
            raise RuntimeError("hash_sets_family_dev: duplicate value inside a set (change the seed)")
        out[c0:c0 + m] = vs
    off = np.arange(n_sketches + 1, dtype=np.uint64) * np.uint64(n_per)
    return out.view(-1), off
