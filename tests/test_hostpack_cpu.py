"""CPU tier: the host-side ASCII -> 2-bit packer (hostpack.cpp, the host half of db200_sketch_batch's upload) against a
numpy restatement of alph::DNA4 (bonsai/include/bonsai/alphabet.h:128) — every byte value, ragged lengths, every ISA
variant boundary."""
import os
import subprocess
import sys

import numpy as np
import pytest


def restate(a):
    a = np.asarray(a, dtype=np.uint8)
    n = a.size
    ng = (n + 15) // 16
    pad = np.zeros(ng * 16, np.uint8)
    pad[:n] = a
    x = pad.astype(np.uint32)
    code = (((x >> 1) ^ (x >> 2)) & 3)
    up = x & 0xDF
    ok = ((up == 0x41) | (up == 0x43) | (up == 0x47) | (up == 0x54)) & (np.arange(ng * 16) < n)
    code = np.where(np.arange(ng * 16) < n, code, 0).reshape(ng, 16)
    codes = (code << (2 * np.arange(16, dtype=np.uint32))).sum(axis=1).astype(np.uint32)
    valid = (ok.reshape(ng, 16).astype(np.uint32) << np.arange(16, dtype=np.uint32)).sum(axis=1).astype(np.uint16)
    return codes, valid


def test_hostpack_matches_dna4_table(capi):
    isa = capi.lib.db200_hostpack_isa().decode()
    assert isa in ("avx512bw", "avx2", "scalar")
    rng = np.random.default_rng(0)
    # every byte value, in every position of a 64-byte vector
    allb = np.tile(np.arange(256, dtype=np.uint8), 64)
    allb = np.concatenate([np.roll(allb, s) for s in range(0, 64, 7)])
    for a in (allb, np.frombuffer(b"ACGTacgtNnUuRYKM-*\n>@", dtype=np.uint8)):
        for n in (0, 1, 15, 16, 17, 31, 32, 33, 63, 64, 65, 127, 128, 129, a.size):
            c, v = capi.hostpack(a[:n])
            wc, wv = restate(a[:n])
            assert np.array_equal(c, wc) and np.array_equal(v, wv), (isa, n)
    seq = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=1_000_003)]
    seq[rng.integers(0, seq.size, size=500)] = ord("N")
    seq[1000:2000] |= 0x20
    c, v = capi.hostpack(seq)
    wc, wv = restate(seq)
    assert np.array_equal(c, wc) and np.array_equal(v, wv)
    # decoding the codes gives the sequence back wherever it is valid
    dec = np.frombuffer(b"ACGT", dtype=np.uint8)[(c[:, None] >> (2 * np.arange(16, dtype=np.uint32))) & 3].reshape(-1)[: seq.size]
    ok = ((v[:, None] >> np.arange(16, dtype=np.uint16)) & 1).astype(bool).reshape(-1)[: seq.size]
    assert np.array_equal(dec[ok], seq[ok] & 0xDF) and ok.sum() == seq.size - np.isin(seq & 0xDF, [ord("N")]).sum()


@pytest.mark.parametrize("isa", ["avx2", "scalar"])
def test_hostpack_other_isa_variants(isa):
    """The variant is picked once per process: the narrower ones run in a child with DB200_HOSTPACK_ISA capping the choice."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r); import numpy as np; from dashing_b200 import capi; import test_hostpack_cpu as t; "
            "assert capi.lib.db200_hostpack_isa().decode() in (%r, 'scalar'); t.test_hostpack_matches_dna4_table(capi); print('ok')"
            % (root, os.path.join(root, "tests"), isa))
    r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, DB200_HOSTPACK_ISA=isa), capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr[-2000:]
