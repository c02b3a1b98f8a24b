#!/usr/bin/env python3
"""Transcribe the reference's golden vectors into a numbers-only JSON fixture.

Run in the build container only (needs /root/reference, which does not exist
on the GPU box):

    python tests/golden/make_golden.py

Sources (reference file:line):
  * fft/_test_values.mojo:8-1107   65 real-input 1-D vectors -> complex spectra
  * fft/tests.mojo:274-371         the 56 (length, user bases) combinations
  * fft/tests.mojo:422-458         2-D 6x4 uint8 image + expected spectrum
  * fft/tests.mojo:613-905         3-D 6x4x8 uint8 volume + expected spectrum
Only numbers are written out; no Mojo source is copied into the repo.
"""
import json
import os
import re
import sys

REF = os.environ.get("B200FFT_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def _numbers(text):
    return [float(t) for t in re.findall(r"-?\d+(?:\.\d+)?(?:[eE][-+]?\d+)?", text)]


def _parse_complex_list(text, ctor):
    """Parse 'Ctor(re[, im]), Ctor(re[, im]) ...' into [[re, im], ...]."""
    out = []
    for m in re.finditer(re.escape(ctor) + r"\(([^()]*)\)", text):
        parts = [p.strip() for p in m.group(1).split(",") if p.strip()]
        re_ = float(parts[0])
        im_ = float(parts[1]) if len(parts) > 1 else 0.0
        out.append([re_, im_])
    return out


def parse_test_values(path):
    src = open(path).read()
    chunks = re.split(r"def _get_test_values_(\d+)\[", src)
    table = {}
    # chunks = [preamble, len, body, len, body, ...]
    for k in range(1, len(chunks), 2):
        length = int(chunks[k])
        body = chunks[k + 1]
        body = body[body.index("res = [") + len("res = ["):]
        vectors = []
        # every vector is `{ [ints...], [Complex(..), ...] }`
        for m in re.finditer(r"\{\s*\[([^\]]*)\]\s*,\s*\[(.*?)\]\s*,?\s*\}", body, re.S):
            xs = [int(v) for v in _numbers(m.group(1))]
            spec = _parse_complex_list(m.group(2), "Complex")
            assert len(xs) == length, (length, len(xs))
            assert len(spec) == length, (length, len(spec))
            vectors.append({"x": xs, "X": spec})
        assert vectors, length
        table[str(length)] = vectors
    return table


def parse_cases(path):
    """The (length, bases) combinations exercised by `_test_fft`."""
    src = open(path).read()
    body = src[src.index("def _test_fft["):src.index("comptime _test[")]
    cases = []
    for line in body.splitlines():
        s = line.strip()
        if s.startswith("#"):
            continue
        m = re.match(r"func\[\[([\d,\s]+)\],\s*values_(\d+)\]\(\)", s)
        if m:
            bases = [int(v) for v in m.group(1).split(",") if v.strip()]
            cases.append({"length": int(m.group(2)), "bases": bases})
    return cases


def _nested_block(src, name):
    start = src.index("comptime " + name)
    eq = src.index("=", src.index("]", start))  # skip the type annotation
    # the type annotation itself contains brackets; find the `] = [` marker
    marker = src.index("] = [", start) + 4
    depth = 0
    for pos in range(marker, len(src)):
        ch = src[pos]
        if ch == "[":
            depth += 1
        elif ch == "]":
            depth -= 1
            if depth == 0:
                return src[marker:pos + 1]
    raise ValueError(name)


def parse_nd(path):
    src = open(path).read()
    in2 = [int(v) for v in _numbers(_nested_block(src, "input_2d"))]
    ex2 = _parse_complex_list(_nested_block(src, "expected_2d"), "Co")
    in3 = [int(v) for v in _numbers(_nested_block(src, "input_3d"))]
    ex3 = _parse_complex_list(_nested_block(src, "expected_3d"), "Co")
    assert len(in2) == 24 and len(ex2) == 24, (len(in2), len(ex2))
    assert len(in3) == 192 and len(ex3) == 192, (len(in3), len(ex3))
    return {
        "2d": {"dims": [6, 4], "x": in2, "X": ex2},
        "3d": {"dims": [6, 4, 8], "x": in3, "X": ex3},
    }


def main():
    fft_dir = os.path.join(REF, "fft")
    if not os.path.isdir(fft_dir):
        sys.exit("reference not found at %s (this script only runs in the build container)" % REF)
    vectors = parse_test_values(os.path.join(fft_dir, "_test_values.mojo"))
    cases = parse_cases(os.path.join(fft_dir, "tests.mojo"))
    nd = parse_nd(os.path.join(fft_dir, "tests.mojo"))
    n_vec = sum(len(v) for v in vectors.values())
    out = {
        "source": "martinvuyk/hackathon-fft: fft/_test_values.mojo, fft/tests.mojo:274-371,422-458,613-905",
        "atol": 1e-2,
        "rtol": 1e-5,
        "vectors_1d": vectors,
        "cases_1d": cases,
        "nd": nd,
    }
    dst = os.path.join(HERE, "reference_vectors.json")
    with open(dst, "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print("wrote %s: %d 1-D vectors over %d lengths, %d (length,bases) cases, 2-D + 3-D" % (
        dst, n_vec, len(vectors), len(cases)))


if __name__ == "__main__":
    main()
