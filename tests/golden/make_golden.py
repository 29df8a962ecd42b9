"""Regenerates the committed golden fixtures.  Run HERE (the container that mounts
/root/reference): the t1ha2 / k-mer vectors come from the reference's own src/cuda_kernel.cu
compiled as host code (oracle/_ref, built by oracle/Makefile), the rest from the oracle.

    python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle as O  # noqa: E402
from hypergen_b200 import synth  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def dump(name, obj):
    with open(os.path.join(HERE, name), "w") as f:
        json.dump(obj, f, indent=1)
        f.write("\n")


def main():
    O.build(force=True)
    assert O.ref() is not None, "oracle/_ref missing: /root/reference not mounted?"
    rng = np.random.default_rng(20261017)

    # 1. t1ha2_atonce: upstream self-check constants + the reference's device function run on the host
    pat = bytes(range(64))
    upstream = [["", (1 << 64) - 1, 0x772C7311BE32FF42], [pat[:1].hex(), 1, 0x71F6DF5DA3B4F532],
                [pat[:2].hex(), 2, 0x555859635365F660], [pat[:3].hex(), 4, 0xE98808F1CD39C626],
                [pat[:4].hex(), 8, 0x2EB18FAF2163BB09], [pat[:5].hex(), 16, 0x7B9DD892C8019C87],
                [pat[:6].hex(), 32, 0xE2B1431C4DA4D15A], ["", 0, 0]]
    ref_vecs = []
    for L in range(0, 33):
        for _ in range(6):
            d = rng.integers(0, 256, L, dtype=np.uint8)
            s = int(rng.integers(0, 1 << 63)) * 2 + int(rng.integers(0, 2))
            ref_vecs.append([d.tobytes().hex(), s, O.ref_t1ha2_atonce(d, s)])
    ref_vecs.append([b"ACGTACGTACGTACGTACGTA".hex(), 123, O.ref_t1ha2_atonce(b"ACGTACGTACGTACGTACGTA", 123)])
    dump("t1ha2_vectors.json", dict(
        source="upstream: t1ha t1ha2_atonce self-check constants; reference: /root/reference/src/cuda_kernel.cu:196-246 compiled as host code (oracle/_ref)",
        upstream=upstream, reference=ref_vecs))

    # 2. k-mer hash sets from the reference kernel body (cuda_kernel.cu:250-321) run on the host
    g = synth.genome(0xB200 + 77, 60_000).numpy().copy()
    g[5000:5030] = ord("N")
    g[20000:26000] |= 0x20
    sets = {}
    for canonical in (True, False):
        for k, scaled in ((21, 100), (31, 50), (15, 200)):
            hs = O.ref_kmer_hash_set(g, k=k, scaled=scaled, canonical=canonical)
            sets["k%d_s%d_c%d" % (k, scaled, int(canonical))] = dict(
                k=k, scaled=scaled, canonical=canonical, n=int(hs.size),
                sha256=hashlib.sha256(hs.tobytes()).hexdigest(), head=[int(x) for x in hs[:4]])
    dump("kmer_ref_sets.json", dict(
        source="reference cuda_kmer_t1ha2 compiled as host code; genome = synth.genome(0xB200+77, 60000) with N at [5000,5030) and lower case at [20000,26000)",
        sets=sets))

    # 3. wyrng (wyhash-rs README vectors)
    dump("wyrng_vectors.json", dict(source="wyhash-rs README: WyRng::seed_from_u64(3).next_u64(); wyrng(&mut 0) x3",
                                    vectors=[[3, [0x03E99A772750DCBE]],
                                             [0, [0x111CB3A78F59A58E, 0xCEABD938FF4E856D, 0x61FB51318F47D2A4]]]))

    # 4. full-size synthetic probe (SURVEY.md 8c values + digests of the full vectors)
    g = synth.genome(0xB200, 5_000_000).numpy()
    cases = []
    for scaled, hv_d in ((1500, 4096), (500, 8192)):
        hs = O.kmer_hash_set(g, scaled=scaled)
        hv = O.encode_hd(hs, hv_d)
        assert np.array_equal(hv, O.encode_hd_avx2_intrinsics(hs, hv_d))
        b, p = O.compress_hd_sketch(hv)
        cases.append(dict(scaled=scaled, hv_d=hv_d, n_hashes=int(hs.size), quant_bits=b, norm2=O.hv_l2_norm_sq(hv),
                          hv_head=[int(x) for x in hv[:8]], hv_min=int(hv.min()), hv_max=int(hv.max()),
                          smallest_hashes=[int(x) for x in hs[:3]],
                          hashes_sha256=hashlib.sha256(hs.tobytes()).hexdigest(),
                          hv_sha256=hashlib.sha256(hv.tobytes()).hexdigest(),
                          packed_sha256=hashlib.sha256(p.tobytes()).hexdigest()))
    dump("synthetic_probe.json", dict(source="oracle on synth.genome(0xB200, 5_000_000), k=21 seed=123 canonical", cases=cases))

    # 5. a small sketch -> dist pipeline (family-structured), every intermediate kept
    seq, off = synth.family_batch(12, 120_000, first=4)
    sk = O.sketch_batch(seq.numpy(), off, scaled=300, hv_d=1024)
    ani, dot = O.dist_all(sk["hv"], sk["norm2"], sk["hv"], sk["norm2"], symmetric=True)
    order = O.ani_output_order(ani, 85.0)
    np.savez_compressed(os.path.join(HERE, "small_pipeline.npz"), hv=sk["hv"], packed=sk["packed"][:, :1024 * 2],
                        quant_bits=sk["quant_bits"], norm2=sk["norm2"], n_hashes=sk["n_hashes"], ani=ani, dot=dot,
                        order=order, params=np.array([21, 300, 123, 1, 1024, 12, 120_000, 4], np.int64))
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
