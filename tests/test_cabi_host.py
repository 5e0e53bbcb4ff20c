"""CPU: the C-ABI library loads and exports every symbol include/stf_b200.h declares, the ctypes
stub agrees with the header, and the HOST-side rANS codec (the part of the library that runs on the
CPU by design) is bit-exact against the oracle and the reference's recorded streams.  No CUDA
compute entry point is called here."""
import ctypes
import hashlib
import json
import os
import re

import numpy as np
import pytest

from oracle import entropy as OE
from stf_b200 import _C, ans

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "stf_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(stf_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    names = header_functions()
    assert len(names) >= 20
    L = ctypes.CDLL(_C.LIB_PATH)
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/stf_b200.h but not exported"
    assert sorted(_C.SIGNATURES) == names, "stf_b200/_C.py and include/stf_b200.h disagree"
    assert b"sm_100a" in _C.lib().stf_version()


def test_every_entry_point_is_mapped_to_the_reference_in_integration_md():
    """INTEGRATION.md names, for every exported function, the reference code it replaces."""
    doc = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "INTEGRATION.md")).read()
    missing = [n for n in header_functions() if n not in doc]
    assert not missing, missing


def test_library_contains_sm100a_tcgen05_code():
    """The shipped .so carries sm_100a SASS with tcgen05 (UTC*MMA) and bulk-TMA (UBLKCP) instructions."""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", _C.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    assert re.search(r"UTC\w*MMA", sass), "no tcgen05.mma in SASS"
    assert "UBLKCP" in sass and "LDTM" in sass


@pytest.fixture(scope="module")
def tables():
    cdf, lens, offs = OE.gaussian_tables()
    return cdf, lens, offs, ans.RansTable(cdf, lens, offs)


def test_rans_kat_bytes(tables, golden_dir):
    cdf, lens, offs, tab = tables
    r = json.load(open(os.path.join(golden_dir, "kat.json")))["rans"]
    b = ans.RansEncoder().encode_with_indexes(r["symbols"], r["indexes"], cdf.tolist(), lens.tolist(), offs.tolist())
    assert b.hex() == r["bytes_hex"]
    assert ans.RansDecoder().decode_with_indexes(b, r["indexes"], cdf.tolist(), lens.tolist(), offs.tolist()) == r["symbols"]
    enc = ans.BufferedRansEncoder()
    enc.encode_with_indexes(r["symbols"][:4], r["indexes"][:4], tab)
    enc.encode_with_indexes(r["symbols"][4:], r["indexes"][4:], tab)
    assert enc.flush().hex() == r["bytes_hex"]


def test_rans_reference_streams(tables, golden_dir):
    cdf, lens, offs, tab = tables
    kat = json.load(open(os.path.join(golden_dir, "kat.json")))
    rng = np.random.default_rng(kat["rans_streams_seed"])
    table = OE.scale_table().numpy()
    for rec in kat["rans_streams"]:
        n, spread = rec["n"], rec["spread"]
        ix = rng.integers(0, 64, size=n).astype(np.int32)
        sy = np.rint(rng.standard_normal(n) * table[ix] * spread).astype(np.int32)
        if n < 4:
            sy = np.array([70000], dtype=np.int32)[:n]
        b = ans.encode_array(tab, sy, ix)
        assert hashlib.sha256(b).hexdigest() == rec["sha256"], f"n={n}"
        d = ans.RansDecoder()
        d.set_stream(b)
        half = n // 2
        out = np.concatenate([d.decode_stream_array(ix[:half], tab), d.decode_stream_array(ix[half:], tab)])
        assert np.array_equal(out, sy)


@pytest.mark.parametrize("n", [0, 1, 2, 3, 17, 4096, 200000])
def test_rans_vs_oracle_random(tables, n):
    cdf, lens, offs, tab = tables
    rng = np.random.default_rng(n + 5)
    ix = rng.integers(0, 64, size=n).astype(np.int32)
    sy = np.rint(rng.standard_normal(n) * OE.scale_table().numpy()[ix] * 2.5).astype(np.int32)
    if n >= 17:
        sy[::7] = rng.integers(-5000, 5000, size=sy[::7].size)      # force escapes of several nibble counts
    b = ans.encode_array(tab, sy, ix)
    assert b == OE.rans_encode(sy, ix, cdf, lens, offs)
    assert len(b) % 4 == 0 and len(b) >= 8
    d = ans.RansDecoder()
    d.set_stream(b)
    assert np.array_equal(d.decode_stream_array(ix, tab), sy)


@pytest.mark.parametrize("lo,hi", [(0, 20), (45, 64), (0, 64), (20, 34)])
def test_rans_both_step_variants_match_oracle(tables, lo, hi):
    """The codec picks its step per run from a sample of the indexes: branchy + 256-bucket LUT for runs of narrow CDF rows,
    branch-free + 4096-bucket LUT for wide (high-entropy) ones.  Both must produce the oracle's bytes, single-stream and in
    lockstep groups, including tail symbols of frequency 1 and escapes."""
    cdf, lens, offs, tab = tables
    rng = np.random.default_rng(lo * 64 + hi)
    table = OE.scale_table().numpy()
    syms, idxs = [], []
    for i in range(5):
        n = 3000 + 501 * i
        ix = rng.integers(lo, hi, size=n).astype(np.int32)
        sy = np.rint(rng.standard_normal(n) * table[ix] * (1.0 + i)).astype(np.int32)      # i >= 1: tails, then escapes
        sy[::97] = rng.integers(-4000, 4000, size=sy[::97].size)
        syms.append(sy), idxs.append(ix)
    want = [OE.rans_encode(sy, ix, cdf, lens, offs) for sy, ix in zip(syms, idxs)]
    assert [ans.encode_array(tab, sy, ix) for sy, ix in zip(syms, idxs)] == want
    for threads in (1, 2):
        assert ans.encode_batch(tab, syms, idxs, threads=threads) == want
        decs = []
        for b in want:
            d = ans.RansDecoder()
            d.set_stream(b)
            decs.append(d)
        outs = ans.decode_batch(decs, tab, idxs, threads=threads)
        assert all(np.array_equal(o, s_) for o, s_ in zip(outs, syms))
    d = ans.RansDecoder()
    d.set_stream(want[0])
    assert np.array_equal(d.decode_stream_array(idxs[0], tab), syms[0])


def test_rans_narrow_transfer_format_same_bytes(tables):
    """stf_rans_encode_batch_narrow (int16 symbols, uint8 indexes) and stf_rans_decode_batch_u8 give the bytes / symbols of
    the int32 entry points (narrow and wide runs, escapes within int16, ragged lockstep groups)."""
    cdf, lens, offs, tab = tables
    rng = np.random.default_rng(21)
    table = OE.scale_table().numpy()
    B, n = 7, 2500
    for lo, hi in ((0, 20), (30, 64)):
        ix = rng.integers(lo, hi, size=(B, n)).astype(np.int32)
        sy = np.rint(rng.standard_normal((B, n)) * table[ix] * 2.0).astype(np.int32)
        sy[:, ::53] = rng.integers(-30000, 30000, size=sy[:, ::53].shape)
        want = [ans.encode_array(tab, sy[b], ix[b]) for b in range(B)]
        for threads in (1, 2, 8):
            assert ans.encode_rows(tab, sy, ix, threads=threads) == want
            assert ans.encode_rows(tab, sy.astype(np.int16), ix.astype(np.uint8), threads=threads) == want
            decs = []
            for s_ in want:
                d = ans.RansDecoder()
                d.set_stream(s_)
                decs.append(d)
            out = np.empty((B, n), dtype=np.int32)
            half = n // 2
            ix8 = ix.astype(np.uint8)
            ans.DecodePlan(decs, tab, np.ascontiguousarray(ix8[:, :half]), out[:, :half].copy(), threads=threads).run()
            plan = ans.DecodePlan(decs, tab, np.ascontiguousarray(ix8[:, half:]), np.empty((B, n - half), np.int32), threads=threads)
            plan.run()
            assert np.array_equal(plan._keep[3], sy[:, half:])
    with pytest.raises(TypeError):
        ans.encode_rows(tab, sy.astype(np.int16), ix)


def test_rans_batch_threads_identical(tables):
    cdf, lens, offs, tab = tables
    rng = np.random.default_rng(3)
    syms, idxs = [], []
    for i in range(6):
        n = 1000 + 977 * i
        ix = rng.integers(0, 64, size=n).astype(np.int32)
        idxs.append(ix)
        syms.append(np.rint(rng.standard_normal(n) * OE.scale_table().numpy()[ix]).astype(np.int32))
    single = [ans.encode_array(tab, s, i) for s, i in zip(syms, idxs)]
    assert ans.encode_batch(tab, syms, idxs, threads=4) == single
    decs = []
    for b in single:
        d = ans.RansDecoder()
        d.set_stream(b)
        decs.append(d)
    outs = ans.decode_batch(decs, tab, idxs, threads=3)
    assert all(np.array_equal(o, s) for o, s in zip(outs, syms))
    # more streams than threads: tasks code two streams in lockstep (odd count, unequal lengths, escape-heavy tails)
    for s in syms:
        s[-50:] *= 40                                   # far outside the modelled range -> bypass nibbles
    single = [ans.encode_array(tab, s, i) for s, i in zip(syms[:5], idxs[:5])]
    for threads in (1, 2):
        assert ans.encode_batch(tab, syms[:5], idxs[:5], threads=threads) == single
        decs = []
        for b in single:
            d = ans.RansDecoder()
            d.set_stream(b)
            decs.append(d)
        outs = ans.decode_batch(decs, tab, idxs[:5], threads=threads)
        assert all(np.array_equal(o, s) for o, s in zip(outs, syms[:5]))
        # a second run of the same decoders continues where the first stopped: nothing left -> stream error
        with pytest.raises(ValueError):
            ans.decode_batch(decs, tab, [np.zeros(100000, np.int32) + 63] * 5, threads=threads)


def test_rans_error_paths(tables):
    cdf, lens, offs, tab = tables
    with pytest.raises(ValueError):
        ans.RansTable(np.zeros((2, 5), np.int32), [5, 5], [0, 0])          # rows do not end at 65536
    with pytest.raises(ValueError):
        ans.encode_array(tab, [0, 1], [0])
    with pytest.raises(ValueError):
        ans.encode_array(tab, [0], [64])                                      # index out of range
    d = ans.RansDecoder()
    with pytest.raises(ValueError):
        d.set_stream(b"\0\0\0")
    d.set_stream(ans.encode_array(tab, [1, 2, 3], [3, 3, 3]))
    with pytest.raises(ValueError):
        d.decode_stream_array(np.zeros(100000, np.int32) + 63, tab)          # runs off the end of the stream


def test_pmf_to_quantized_cdf_matches_oracle():
    rng = np.random.default_rng(0)
    for n in (1, 2, 5, 23, 400):
        p = rng.random(n).astype(np.float32) ** 4
        p /= p.sum()
        assert ans.pmf_to_quantized_cdf(p) == OE.pmf_to_quantized_cdf(p).tolist()
    spike = np.zeros(300, np.float32)
    spike[150] = 1.0                                                          # many zero-mass symbols -> stealing
    assert ans.pmf_to_quantized_cdf(spike) == OE.pmf_to_quantized_cdf(spike).tolist()


def test_rans_batch_grouping_matches_single_streams():
    """encode_batch / decode_batch split the images evenly over the threads and code each thread's range in lockstep groups
    of 4 / 3 / 2 / 1: whatever the (count, threads) combination, every string equals the single-stream encode and decodes
    back (ragged lengths included)."""
    from stf_b200 import ans
    cdf, lens, offs = OE.gaussian_tables()
    tab = ans.RansTable(cdf, lens, offs)
    rng = np.random.default_rng(7)
    table = OE.scale_table().numpy()
    imgs = []
    for i in range(11):
        n = int(rng.integers(200, 3000))
        ix = rng.integers(0, 64, size=n).astype(np.int32)
        sy = np.rint(rng.standard_normal(n) * table[ix] * (3.0 if i % 3 == 0 else 1.0)).astype(np.int32)   # incl. escapes
        imgs.append((sy, ix))
    singles = [ans.encode_array(tab, sy, ix) for sy, ix in imgs]
    for count in (1, 2, 3, 5, 7, 11):
        for threads in (1, 2, 3, 4):
            got = ans.encode_batch(tab, [s for s, _ in imgs[:count]], [i for _, i in imgs[:count]], threads=threads)
            assert got == singles[:count], (count, threads)
            decs = []
            for s in got:
                d = ans.RansDecoder()
                d.set_stream(s)
                decs.append(d)
            outs = [np.empty(len(sy), dtype=np.int32) for sy, _ in imgs[:count]]
            ans.decode_batch(decs, tab, [i for _, i in imgs[:count]], outs=outs, threads=threads)
            assert all(np.array_equal(o, sy) for o, (sy, _) in zip(outs, imgs[:count])), (count, threads)
