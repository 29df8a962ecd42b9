"""The run-summary algebra behind the GPU FASTA merge (csrc/fasta.cu), restated in Python and checked against the
oracle's read_merge_seq (reference src/fastx_reader.rs:6-29).  A run of bytes is summarised by
(bytes emitted if it starts outside a header line, bytes emitted if it starts inside one, contains a line start,
state handed on); the kernels rely on such summaries composing associatively, so that blocks can be summarised
independently and chained afterwards.  CPU only - the kernels themselves are checked in tests/test_gpu_fasta.py."""
import numpy as np


def classify(data: bytes):
    """per byte: (line start, header start, emitted when outside a header) - fasta.cu classify()"""
    n = len(data)
    out = []
    for i, c in enumerate(data):
        prev = data[i - 1] if i else 10
        nxt = data[i + 1] if i + 1 < n else 10
        ls = prev == 10
        hs = ls and c == ord(">")
        keep = c != 10 and not (c == 13 and nxt == 10)
        out.append((ls, hs, keep or hs))
    return out


def summarise(cls):
    """(c0, c1, has, hdr) of a run, by running it under both incoming states"""
    res = []
    for state in (False, True):
        h, cnt, has, last = state, 0, False, False
        for ls, hs, emit_in in cls:
            if ls:
                h, has, last = hs, True, hs
            cnt += 1 if (hs or (not h and emit_in)) else 0
        res.append(cnt)
    return res[0], res[1], has, last


def compose(a, b):
    """fasta.cu seg_compose: a, then b"""
    after0 = a[3] if a[2] else False
    after1 = a[3] if a[2] else True
    return (a[0] + (b[1] if after0 else b[0]), a[1] + (b[1] if after1 else b[0]), a[2] or b[2], b[3] if b[2] else a[3])


def _random_fasta(rng, n_rec):
    parts = []
    for r in range(n_rec):
        parts.append(b">rec%d some text > with a bracket\r\n" % r if r % 3 == 0 else b">r%d\n" % r)
        for _ in range(int(rng.integers(0, 6))):
            line = bytes(rng.choice(np.frombuffer(b"ACGTNacgt\r>", np.uint8), int(rng.integers(0, 40))))
            if line[:1] == b">":
                line = b"A" + line[1:]
            parts.append(line + (b"\r\n" if rng.random() < 0.3 else b"\n"))
    body = b"".join(parts)
    return body[:-1] if rng.random() < 0.5 else body  # with or without the final newline


def test_block_summaries_compose_to_the_whole_file(oracle):
    rng = np.random.default_rng(5)
    for trial in range(30):
        data = _random_fasta(rng, int(rng.integers(1, 8)))
        cls = classify(data)
        want = oracle.read_merge_seq(data)
        whole = summarise(cls)
        assert whole[0] == want.size                      # a file starts outside a header
        for block in (1, 3, 16, 64):                      # any block size chains to the same totals
            acc = (0, 0, False, False)                    # the identity run
            for b0 in range(0, len(cls), block):
                acc = compose(acc, summarise(cls[b0:b0 + block]))
            assert acc == whole, (trial, block)


def test_compose_is_associative():
    rng = np.random.default_rng(6)
    runs = [(int(a), int(b), bool(h), bool(d) and bool(h)) for a, b, h, d in rng.integers(0, 5, (200, 4)) % [5, 5, 2, 2]]
    for _ in range(300):
        x, y, z = (runs[i] for i in rng.integers(0, len(runs), 3))
        assert compose(compose(x, y), z) == compose(x, compose(y, z))
