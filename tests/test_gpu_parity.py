"""Parity of the CUDA path against the oracle — every call goes through the C ABI of librrx.so
(include/rrx.h) via ctypes.  Run on the B200 box: pytest -m gpu."""
import random
import re

import numpy as np
import pytest

from conftest import cpu_corpus
from roaringregex_b200.configs import CONFIGS

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch


def _check(rrx, oracle_api, pat, text, cache={}):
    key = pat
    if key not in cache:
        cache[key] = (rrx.RRegex(pat), rrx.RRegex(pat, rrx.COMPILE_NO_REDUCE), oracle_api.Oracle(pat))
    r1, r2, o = cache[key]
    want = o.lazy_matches(text)
    for r in (r1, r2):
        got = r.scan_host(text)
        assert got.shape == want.shape, (pat, got.shape, want.shape)
        assert np.array_equal(got, want), pat


def test_survey_vectors(rrx, oracle_api, torch_cuda):
    for pat, txt, want in [("(ab|cd)*e", "xxababe cde e", [(6, 7), (10, 11), (12, 13)]),
                           ("[a-z]+@[a-z]+\\.(com|org|net)", "foo@bar.com x a@b.org c@d.comet@e.net", [(2, 11), (14, 21), (22, 29), (30, 37)]),
                           ("a{1,100}b", "aaab b ab aab", [(2, 4), (7, 9), (11, 13)]),
                           ("(a|b){1,300}c", "abbac c bc", [(3, 5), (8, 10)]),
                           ("(x|y)*z.*q", "xyzaaqbbq q zq", [(2, 6), (2, 9), (2, 11), (12, 14)]),
                           ("ab|bc", "abc", [(0, 2), (1, 3)]), ("a+", "aaa", [(0, 1), (1, 2), (2, 3)])]:
        got = rrx.RRegex(pat).scan_host(txt.encode())
        assert [tuple(x) for x in got.tolist()] == want, pat
        _check(rrx, oracle_api, pat, txt.encode())


@pytest.mark.parametrize("n", [0, 1, 15, 16, 17, 127, 128, 129, 4095, 4096, 4097, 32767, 32768, 32769, 65536 + 77, 1 << 20])
def test_sizes_around_row_and_tile_edges(rrx, oracle_api, torch_cuda, n):
    rnd = random.Random(n)
    for pat, alpha in [("(ab|cd)*e", b"abcde "), ("(x|y)*z.*q", b"xyzq a"), ("a{1,100}b", b"aaaaab "),
                       ("[a-z]+@[a-z]+\\.(com|org|net)", b"acomnetrg@. ")]:
        text = bytes(rnd.choices(alpha, k=n))
        _check(rrx, oracle_api, pat, text)


@pytest.mark.parametrize("cfg_id,n", [(1, 1 << 20), (2, 16 << 20), ("2v", 16 << 20), (3, 16 << 20), (4, 4 << 20), (5, 16 << 20)])
def test_config_corpora(rrx, oracle_api, torch_cuda, cfg_id, n):
    cfg = CONFIGS[cfg_id]
    text = cpu_corpus(cfg, 0, n)
    _check(rrx, oracle_api, cfg.pattern, text)


def test_match_straddling_every_boundary(rrx, oracle_api, torch_cuda):
    pat = "[a-z]+@[a-z]+\\.(com|org|net)"
    addr = b"someone@example.org"
    for edge in (128, 4096, 32768):
        for shift in range(len(addr) + 2):
            text = bytearray(b" " * (edge + 64))
            text[edge - shift:edge - shift + len(addr)] = addr
            _check(rrx, oracle_api, pat, bytes(text))


def test_unbounded_carry_and_zero_matches(rrx, oracle_api, torch_cuda):
    n = 40 * 32768 + 500
    text = bytearray(b"a" * n)
    text[0] = ord("z")
    text[-1] = ord("q")
    text[70000] = ord("\n")
    _check(rrx, oracle_api, "(x|y)*z.*q", bytes(text))
    text[0] = ord("a")
    got = rrx.RRegex("(x|y)*z.*q").scan_host(bytes(text))
    assert got.shape == (0, 2)
    # a match ending at the final byte, a text shorter than one row
    _check(rrx, oracle_api, "(ab|cd)*e", b"abcde")


def test_hard_rows(rrx, oracle_api, torch_cuda):
    rnd = random.Random(5)
    for run in (128, 299, 300, 301, 1000):
        text = bytearray()
        while len(text) < 200000:
            text += bytes(rnd.choices(b"ab", k=run)) + b"c" + bytes(rnd.choices(b"ab cd", k=rnd.randint(0, 40)))
        _check(rrx, oracle_api, "(a|b){1,300}c", bytes(text))
    _check(rrx, oracle_api, "(a|b){1,300}c", b"ab" * 100000 + b"c")
    _check(rrx, oracle_api, "(a|b)*a(a|b){3}c", bytes(rnd.choices(b"ab", k=300000)) + b"c")


def test_random_regexes(rrx, oracle_api, torch_cuda):
    from test_emulator import _random_regex
    rnd = random.Random(77)
    done = 0
    while done < 60:
        pat = _random_regex(rnd)
        try:
            o = oracle_api.Oracle(pat)
        except Exception:
            continue
        if o.n_raw > 400:
            continue
        done += 1
        text = bytes(rnd.choices(b"abc\n", k=rnd.randint(0, 70000)))
        _check(rrx, oracle_api, pat, text)


def test_device_corpus_equals_cpu_corpus(rrx, oracle_api, torch_cuda):
    torch = torch_cuda
    for cid in (1, 2, "2v", 3, 4, 5):
        cfg = CONFIGS[cid]
        for off, n in ((0, 1 << 20), (12345, 100003), (1 << 33, 4096 + 5)):
            d = torch.empty(n + 64, dtype=torch.uint8, device="cuda")
            rrx.corpus_device(cfg, d.data_ptr(), off, n)
            torch.cuda.synchronize()
            assert np.array_equal(d[:n].cpu().numpy(), cpu_corpus(cfg, off, n)), (cid, off, n)


def test_config2_full_size_properties(rrx, oracle_api, torch_cuda):
    """1 GiB of log records on the device: match count equals the number of valid records (a
    closed-form property of the generator), matches are sorted, a 64 MiB prefix is bit-exact."""
    torch = torch_cuda
    cfg = CONFIGS[2]
    n = cfg.length
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    rrx.corpus_records(d.data_ptr(), 0, n, cfg.seed)
    torch.cuda.synchronize()
    r = rrx.RRegex(cfg.pattern)
    got = r.scan_device(d.data_ptr(), n)
    # kinds of all records straight from the hash: kind = h(seed, 8*rec+2) & 3 ; valid iff kind < 2
    def sm64(z):
        z = (z + np.uint64(0x9E3779B97F4A7C15))
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))
    with np.errstate(over="ignore"):
        s = sm64(np.uint64(cfg.seed))
        rec = np.arange(n // 128, dtype=np.uint64)
        kind = sm64(s + np.uint64(8) * rec + np.uint64(2)) & np.uint64(3)
    assert len(got) == int((kind < 2).sum())
    assert np.all(np.diff(got[:, 1]) > 0) and np.all(got[:, 0] < got[:, 1])
    assert np.array_equal(got[:, 1] >> 7, np.nonzero(kind < 2)[0])          # exactly one match per valid record
    pre = 64 << 20
    want = oracle_api.Oracle(cfg.pattern).lazy_matches(cpu_corpus(cfg, 0, pre))
    assert np.array_equal(got[:len(want)], want)
    # every reported span is a member of the language (size-independent check on a sample)
    rx = re.compile(cfg.pattern.encode(), re.S)
    idx = np.linspace(0, len(got) - 1, 500).astype(int)
    host = d.cpu().numpy()
    for s_, e_ in got[idx]:
        assert rx.fullmatch(host[s_:e_].tobytes())


@pytest.mark.parametrize("cfg_id,n", [(1, 1 << 20), (2, 8 << 20), (3, 8 << 20), (4, 4 << 20), (5, 8 << 20)])
def test_runtime_specialised_kernel(rrx, oracle_api, torch_cuda, cfg_id, n):
    """The NVRTC-specialised kernel (regex compiled to straight-line code) against the oracle and
    against the pre-compiled generic kernel."""
    torch = torch_cuda
    cfg = CONFIGS[cfg_id]
    text = cpu_corpus(cfg, 0, n)
    d = torch.from_numpy(text).cuda()
    want = oracle_api.Oracle(cfg.pattern).lazy_matches(text)
    r = rrx.RRegex(cfg.pattern)
    for jit in (True, False):
        sc = rrx.Scanner(r, n, len(want) + 16, jit=jit)
        assert sc.is_jit == jit
        sc.scan(d.data_ptr(), n)
        res = sc.result()
        assert res["error_bits"] == 0 and res["n_matches"] == len(want)
        assert np.array_equal(sc.matches(), want), (cfg_id, jit)


def test_runtime_specialised_random_patterns(rrx, oracle_api, torch_cuda):
    from test_emulator import _random_regex
    torch = torch_cuda
    rnd = random.Random(4242)
    done = 0
    while done < 16:
        pat = _random_regex(rnd)
        try:
            o = oracle_api.Oracle(pat)
            r = rrx.RRegex(pat)
        except Exception:
            continue
        if r.info["uses_rows"]:
            continue
        done += 1
        text = np.frombuffer(bytes(rnd.choices(b"abc\n", k=200000)), dtype=np.uint8)
        d = torch.from_numpy(text.copy()).cuda()
        want = o.lazy_matches(text)
        sc = rrx.Scanner(r, len(text), len(want) + 16, jit=True)
        sc.scan(d.data_ptr(), len(text))
        assert sc.result()["n_matches"] == len(want), pat
        assert np.array_equal(sc.matches(), want), pat


def _run_world_gpu(rrx, torch, pattern, text, world, halo):
    """`world` shards of one device buffer, scanned by `world` threads through the product's
    DeviceBackend (librrx), exchanging summaries through an in-process allgather."""
    import threading
    from roaringregex_b200 import sharded
    from test_sharded import ThreadComm
    n = len(text)
    cuts = [((n * r // world) // 128) * 128 for r in range(world)] + [n]
    d = torch.from_numpy(np.ascontiguousarray(text)).cuda()
    comm = ThreadComm(world)
    out, errs = [None] * world, []
    regex = rrx.RRegex(pattern)

    def work(rank):
        try:
            torch.cuda.set_device(0)
            lo, hi = cuts[rank], cuts[rank + 1]
            be = sharded.DeviceBackend(regex, d.data_ptr() + lo, hi - lo, lo, min(halo, lo), (hi - lo) // 2 + 4096)
            res = sharded.scan_sharded(regex, be, comm.view(rank), rank, world, fetch=True)
            out[rank] = res.to_host(be)
        except Exception as e:      # pragma: no cover
            errs.append(e)
            comm.bar.abort()
    th = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    [t.start() for t in th]
    [t.join() for t in th]
    if errs:
        raise errs[0]
    return np.concatenate(out)


@pytest.mark.parametrize("cfg_id", [2, 5])
def test_sharded_on_device(rrx, oracle_api, torch_cuda, cfg_id):
    cfg = CONFIGS[cfg_id]
    text = cpu_corpus(cfg, 0, (8 << 20) + 4096 * 3)
    want = oracle_api.Oracle(cfg.pattern).lazy_matches(text)
    for world in (2, 4):
        got = _run_world_gpu(rrx, torch_cuda, cfg.pattern, text, world, halo=1 << 16)
        assert np.array_equal(got, want), (cfg_id, world)


@pytest.mark.parametrize("jit", [False, True])
def test_shard_ending_on_a_row_boundary_inside_a_tile(rrx, oracle_api, torch_cuda, monkeypatch, jit):
    """A run alive at the last byte of a shard whose last row is full but not its tile's last: the state at the end of
    the text is the shard's exit (the straddle kernel must not follow the run into a row that does not exist).
    jit=False: the pre-compiled kernels (built with the straddle code); jit=True: the straddle variant, forced."""
    torch = torch_cuda
    import random
    if jit:
        monkeypatch.setenv("RRX_JIT_VARIANT", "1")
    rnd = random.Random(5)
    n = 9000
    for pat, alpha, wts in [("a.", b"abc\xe9\x80", [200, 1, 1, 1, 200]), ("y[a-z]+q", b"abcxyzq ", [40, 1, 5, 200, 40, 1, 200, 1]),
                            ("[a-z]+z[a-z]+", b"abcxyzq ", [1, 5, 5, 200, 1, 5, 200, 5]), ("c[a-z]+", b"ab@.com ", [1, 5, 1, 200, 200, 200, 1, 1])]:
        o = oracle_api.Oracle(pat)
        r = rrx.RRegex(pat)
        sc = rrx.Scanner(r, n, n + 64, jit=jit)
        nw = r.info["words"]
        for cut in (128, 1920, 4480):
            for _ in range(3):
                text = np.frombuffer(bytes(rnd.choices(alpha, weights=wts, k=n)), dtype=np.uint8)
                want = o.lazy_matches(text)
                d = torch.zeros(n + 64, dtype=torch.uint8, device="cuda")
                d[:n] = torch.from_numpy(text.copy()).cuda()
                sc.scan(d.data_ptr(), cut)
                assert sc.result()["error_bits"] == 0
                left = sc.matches()
                comp = sc.shard_summary()
                assert comp[0] & 0xff in (0, 1)
                sc.scan(d.data_ptr() + cut, n - cut, base_offset=cut, left_context=cut, entry=comp[1:1 + nw].copy())
                res = sc.result()
                assert res["error_bits"] == 0 and res["n_pending"] == 0
                got = np.concatenate([left, sc.matches()])
                assert np.array_equal(got, want), (pat, cut, jit)


def test_sharded_start_scan_travels_through_all_shards(rrx, oracle_api, torch_cuda, monkeypatch):
    from roaringregex_b200 import sharded
    monkeypatch.setattr(sharded, "PREFIX_WINDOW", 8192)
    n = 4 * 65536
    text = np.full(n, ord("a"), dtype=np.uint8)
    text[0] = ord("z")
    text[-5] = ord("q")
    got = _run_world_gpu(rrx, torch_cuda, "(x|y)*z.*q", text, 4, halo=256)
    assert got.tolist() == [[0, n - 4]]


def test_no_overlap_view(rrx, oracle_api, torch_cuda):
    """RRX_MODE_NO_OVERLAP on the device against the oracle's filter: sparse and
    heavy overlap (every match sharing one start), empty lists, and the config corpora."""
    rnd = random.Random(3)
    cases = [("(x|y)*z.*q", b"xyzaaqbbq q zq"), ("a+", b"aaa b aa"), ("ab|bc", b"abc abcbc"), ("(ab|cd)*e", b""), ("a+", b"bbbb")]
    text = bytearray(b"a" * 200000)
    text[0] = ord("z")
    for i in range(50, len(text), 97):
        text[i] = ord("q")
    cases.append(("(x|y)*z.*q", bytes(text)))                       # all matches start at offset 0
    cases.append(("(a|b)*a(a|b){3}c", bytes(rnd.choices(b"abc", k=100000))))
    cases.append(("[ab]+c", bytes(rnd.choices(b"abc ", k=100000))))
    for cid, n in ((1, 1 << 20), (2, 4 << 20), (3, 2 << 20), (5, 4 << 20)):
        cases.append((CONFIGS[cid].pattern, cpu_corpus(CONFIGS[cid], 0, n)))
    for pat, txt in cases:
        want = oracle_api.Oracle(pat).no_overlap_matches(txt)
        got = rrx.RRegex(pat).scan_host(txt, 0, rrx.MODE_NO_OVERLAP)
        assert got.shape == want.shape and np.array_equal(got, want), (pat, got.shape, want.shape)
        if len(txt):
            t = torch_cuda.tensor(np.frombuffer(bytes(txt), dtype=np.uint8), device="cuda")
            got_d = rrx.RRegex(pat).scan_device(t.data_ptr(), len(txt), 0, rrx.MODE_NO_OVERLAP)
            assert np.array_equal(got_d, want), pat


def test_host_scans_reuse_buffers_and_stay_exact(rrx, oracle_api, torch_cuda):
    """rrx_scan_host keeps its device text buffer, scanner and pinned match buffers between calls:
    growing, shrinking and overflowing (many matches) inputs on ONE program stay exact, and results of
    earlier calls stay valid while later calls run."""
    pat = "(ab|cd)*e"
    r, o = rrx.RRegex(pat), oracle_api.Oracle(pat)
    rnd = random.Random(8)
    kept = []
    for n in (1000, 3 << 20, 17, 1 << 20, 0, 5 << 20):
        text = bytes(rnd.choices(b"abcdeee ", k=n))             # dense in matches: exercises the capacity retry
        got = r.scan_host(text)
        want = o.lazy_matches(text)
        assert np.array_equal(got, want), n
        kept.append((got, want))
    for got, want in kept:
        assert np.array_equal(got, want)


@pytest.mark.parametrize("cfg_id,n", [(3, 1 << 30), (4, 1 << 30), (5, 2 << 30)])
def test_large_corpora_properties(rrx, oracle_api, torch_cuda, cfg_id, n):
    """Configs 3-5 at GiB scale, through size-independent properties: ends strictly ascending and one
    match per end; every end is where the pattern's last byte sits and no such byte is missed
    (closed forms of these regexes); each sampled span is in the language and no later start works
    (laziness); scanning in two pieces with the carried state gives the same list as scanning whole;
    a 32 MiB prefix is bit-exact against the oracle."""
    torch = torch_cuda
    cfg = CONFIGS[cfg_id]
    d = torch.empty(n + 256, dtype=torch.uint8, device="cuda")
    rrx.corpus_iid(d.data_ptr(), 0, n, cfg.seed, cfg.alpha)
    torch.cuda.synchronize()
    r = rrx.RRegex(cfg.pattern)
    got = r.scan_device(d.data_ptr(), n)
    assert np.all(np.diff(got[:, 1]) > 0) and np.all(got[:, 0] < got[:, 1]) and got[-1, 1] <= n
    host = d[:n].cpu().numpy()
    last = {3: ord("b"), 4: ord("c"), 5: ord("q")}[cfg_id]
    prev_ok = {3: (ord("a"),), 4: (ord("a"), ord("b")), 5: None}[cfg_id]
    # closed form of the end set: the last byte of the pattern preceded by a letter of the counted class
    # (configs 3, 4) / preceded anywhere by a 'z' (config 5)
    pos = np.nonzero(host == last)[0]
    if prev_ok is not None:
        pos = pos[pos > 0]
        pos = pos[np.isin(host[pos - 1], prev_ok)]
    else:
        first_z = int(np.argmax(host == ord("z")))
        pos = pos[pos > first_z]
    assert np.array_equal(got[:, 1], pos + 1)
    rx = re.compile(cfg.pattern.encode(), re.S)
    for s_, e_ in got[np.linspace(0, len(got) - 1, 300).astype(int)]:
        span = host[s_:e_].tobytes()
        assert rx.fullmatch(span)
        assert not any(rx.fullmatch(span[k:]) for k in range(1, min(len(span), 400)))     # lazy: no later start
    # two pieces with the carried state == whole
    cut = (n // 3) & ~4095
    cap = len(got) + 1024
    sc = rrx.Scanner(r, n, cap)
    sc.scan(d.data_ptr(), cut)
    left, summary = sc.matches(), sc.shard_summary()
    entry = np.zeros(r.info["words"], dtype=np.uint32)
    from roaringregex_b200 import sharded
    entry = sharded.apply_summary(r, summary, entry)
    sc.scan(d.data_ptr() + cut, n - cut, cut, cut, entry)
    res = sc.result()
    assert res["error_bits"] == 0 and res["n_pending"] == 0
    both = np.concatenate([left, sc.matches()])
    assert np.array_equal(both, got)
    pre = 32 << 20
    want = oracle_api.Oracle(cfg.pattern).lazy_matches(cpu_corpus(cfg, 0, pre))
    assert np.array_equal(got[:len(want)], want)


def test_device_digest_equals_host_digest(rrx, oracle_api, torch_cuda):
    """rrx_digest_device / rrx_scanner_digest (csrc/rrx_hash.cu) against oracle.api.digest (numpy) on the
    same lists: full, with an end limit, empty, and a list with order violations."""
    torch = torch_cuda
    cfg = CONFIGS[5]
    text = cpu_corpus(cfg, 0, 4 << 20)
    want = oracle_api.Oracle(cfg.pattern).lazy_matches(text)
    r = rrx.RRegex(cfg.pattern)
    sc = rrx.Scanner(r, len(text), len(want) + 16)
    sc.scan(torch.from_numpy(text).cuda().data_ptr(), len(text))
    for lim in (1 << 62, 1 << 20, 12345, 0):
        assert sc.digest(lim) == oracle_api.digest(want, lim), lim
    import ctypes as C
    bad = want[::-1].copy()
    d = torch.from_numpy(bad).cuda()
    out = np.zeros(4, dtype=np.uint64)
    assert rrx.lib().rrx_digest_device(d.data_ptr(), len(bad), 1 << 62, out.ctypes.data, None) == 0
    assert tuple(int(x) for x in out) == oracle_api.digest(bad) and out[3] == len(bad) - 1
    assert rrx.lib().rrx_digest_device(None, 0, 1 << 62, out.ctypes.data, None) == 0 and not out.any()


def test_offsets_beyond_4gib(rrx, oracle_api, torch_cuda):
    """A 5 GiB shard whose global offsets start above 2^32 (config 3's corpus from offset 2^32 + 12345*128 on):
    tile counts, staging slots and every offset the kernels form exceed 32 bits.  Bit-exact against the
    oracle on windows at the start, around local offset 4 GiB, and at the end; match count against the
    closed form ('ab' occurrences); digest consistent with the count."""
    torch = torch_cuda
    cfg = CONFIGS[3]
    n = 5 << 30
    off = (1 << 32) + 12345 * 128
    d = torch.empty(n + 256, dtype=torch.uint8, device="cuda")
    rrx.corpus_device(cfg, d.data_ptr(), off, n)
    torch.cuda.synchronize()
    r = rrx.RRegex(cfg.pattern)
    sc = rrx.Scanner(r, n, n // 160)
    sc.scan(d.data_ptr(), n, off)
    res = sc.result()
    assert res["error_bits"] == 0 and res["n_matches"] <= res["capacity"]
    import bench
    assert res["n_matches"] == bench.closed_form_count(torch, cfg, d[:n], n)
    dg = sc.digest()
    assert dg[0] == res["n_matches"] and dg[3] == 0
    got = sc.matches()
    assert got[0, 1] > off and got[-1, 1] <= off + n
    o = oracle_api.Oracle(cfg.pattern)
    W = 1 << 20
    for lo in (0, (1 << 32) - W // 2, (4 << 30) - 4096 * 3 - 77, n - W):
        ov = min(lo, 4096)
        want = o.lazy_matches(cpu_corpus(cfg, off + lo - ov, W + ov), base=off + lo - ov)
        want = want[want[:, 1] > off + lo]
        sel = got[(got[:, 1] > off + lo) & (got[:, 1] <= off + lo + W)]
        assert np.array_equal(sel, want), lo


def test_streaming_host_scan(rrx, oracle_api, torch_cuda, monkeypatch):
    """rrx_scan_host on texts of several chunks (chunked H2D overlapped with scanning, state carried from chunk to
    chunk; pageable input through pinned bounce buffers): identical to the oracle's one-piece list, for every config,
    pinned and pageable host memory, chunk size forced down to 1 MiB so that small texts have many chunks."""
    torch = torch_cuda
    monkeypatch.setenv("RRX_HOST_CHUNK", str(1 << 20))
    for cid, n in ((1, 5 << 20), (2, 9 << 20), ("2v", 9 << 20), (3, 7 << 20), (4, 4 << 20), (5, 9 << 20)):
        cfg = CONFIGS[cid]
        text = cpu_corpus(cfg, 0, n + 12345)
        want = oracle_api.Oracle(cfg.pattern).lazy_matches(text)
        r = rrx.RRegex(cfg.pattern)
        got = r.scan_host(text)                                   # pageable numpy buffer
        assert np.array_equal(got, want), (cid, "pageable")
        pinned = torch.from_numpy(text).pin_memory()
        got = r.scan_host(pinned.numpy())
        assert np.array_equal(got, want), (cid, "pinned")
        old = rrx.RRegex(cfg.pattern, rrx.COMPILE_NO_DONLY)
        assert np.array_equal(old.scan_host(text), want), (cid, "row-bounded scan")
    # the unbounded carry across every chunk, and a match whose start lies many chunks back
    text = bytearray(b"a" * (6 << 20))
    text[5] = ord("z")
    text[-7] = ord("q")
    want = oracle_api.Oracle("(x|y)*z.*q").lazy_matches(bytes(text))
    assert want.tolist() == [[5, len(text) - 6]]
    for flags in (0, rrx.COMPILE_NO_DONLY):
        assert np.array_equal(rrx.RRegex("(x|y)*z.*q", flags).scan_host(bytes(text)), want), flags
    # default chunk size (64 MiB) on a 300 MiB text
    monkeypatch.delenv("RRX_HOST_CHUNK")
    cfg = CONFIGS[2]
    n = 300 << 20
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    rrx.corpus_device(cfg, d.data_ptr(), 0, n)
    host = d.cpu().numpy()
    r = rrx.RRegex(cfg.pattern)
    whole = r.scan_device(d.data_ptr(), n)
    assert np.array_equal(r.scan_host(host), whole)


def test_no_overlap_view_chunked(rrx, oracle_api, torch_cuda):
    """rrx_scanner_no_overlap: the non-overlapping view computed piece by piece, the 8-byte carry (end of the last
    selected match) passed from piece to piece, equals the oracle's filter over the whole text."""
    torch = torch_cuda
    rnd = random.Random(9)
    text = bytearray(b"a" * 300000)
    text[3] = ord("z")
    for i in range(50, len(text), 211):
        text[i] = ord("q")
    cases = [("(x|y)*z.*q", bytes(text)), ("(a|b)*a(a|b){3}c", bytes(rnd.choices(b"abc", k=200000))), ("[ab]+c", bytes(rnd.choices(b"abc ", k=150000)))]
    for cid, n in ((2, 3 << 20), (5, 2 << 20), (3, 2 << 20)):
        cases.append((CONFIGS[cid].pattern, cpu_corpus(CONFIGS[cid], 0, n)))
    for pat, txt in cases:
        txt = np.frombuffer(bytes(txt), dtype=np.uint8)
        want = oracle_api.Oracle(pat).no_overlap_matches(txt)
        r = rrx.RRegex(pat)
        d = torch.from_numpy(txt.copy()).cuda()
        n = len(txt)
        cuts = [0, (n // 3) & ~127, (2 * n // 3) & ~127, n]
        sc = rrx.Scanner(r, n, n // 2 + 1024)
        entry = np.zeros(r.info["words"], dtype=np.uint32)
        from roaringregex_b200 import sharded
        carry, parts = 0, []
        for lo, hi in zip(cuts[:-1], cuts[1:]):
            sc.scan(d.data_ptr() + lo, hi - lo, lo, lo, entry)
            assert sc.result()["error_bits"] == 0
            entry = sharded.apply_summary(r, sc.shard_summary(), entry)
            part, carry = sc.no_overlap(carry)
            parts.append(part)
        got = np.concatenate(parts)
        assert got.shape == want.shape and np.array_equal(got, want), pat


def test_anchors(rrx, oracle_api, torch_cuda):
    """`^` and `$`: the one-shot scan calls frame the text with one virtual 0x00 on each side that
    only the anchor positions consume; the oracle does the same on the reference's tables (lazy_matches_anchored).
    Multi-line texts: the anchors hold at the text's ends only, as in POSIX ERE without REG_NEWLINE."""
    torch = torch_cuda
    rnd = random.Random(12)
    pats = ["^ab", "ab$", "^ab$", "(^|x)ab", "(x|^)ab", "ab($|c)", "^a+b|cd$", "^.*b", "[^x]b$", "^(ab|cd)*e", "(ab|cd)*e$",
            "^[a-z]+@[a-z]+\\.(com|org|net)", "[a-z]+@[a-z]+\\.(com|org|net)$", "^a{1,100}b", "a{1,100}b$", "a*$", "^", "$", "^$", "x^a", "a$b"]
    texts = [b"", b"a", b"ab", b"ab xab ab", b"abc ab", b"aab cd aab cd", b"ab\nab\nab", b"abab" * 50 + b"e",
             b"bob@example.com and ann@site.org", b"zed@q.net", b"a" * 300 + b"b", b"cdabe cd"]
    for n in (127, 128, 129, 4096, 5000, 70000):
        texts.append(bytes(rnd.choices(b"abcde x\n", k=n)))
        texts.append(b"ab" + bytes(rnd.choices(b"abcd \n", k=n)) + b"ab")
    for pat in pats:
        o = oracle_api.Oracle(pat)
        for flags in (0, rrx.COMPILE_NO_DONLY):
            r = rrx.RRegex(pat, flags)
            for txt in texts:
                want = o.lazy_matches_anchored(txt)
                got = r.scan_host(txt)
                assert got.shape == want.shape and np.array_equal(got, want), (pat, flags, txt[:40], got[:4].tolist(), want[:4].tolist())
                if len(txt):
                    d = torch.tensor(np.frombuffer(txt, dtype=np.uint8), device="cuda")
                    assert np.array_equal(r.scan_device(d.data_ptr(), len(txt)), want), (pat, flags, "device")
    # a long text: the anchored thread of `^.*b` is carried through every tile
    txt = b"a" * 200000 + b"b" + b"a" * 1000 + b"b"
    for pat in ("^.*b", "a.*b$", "^a*b"):
        want = oracle_api.Oracle(pat).lazy_matches_anchored(txt)
        assert np.array_equal(rrx.RRegex(pat).scan_host(txt), want), pat
    # more start walks reaching the text start than the default pending list holds (65,536): the one-shot calls rescan
    # with a list sized by the count the kernel reported
    txt = b"b" * 70000
    want = oracle_api.Oracle("^.*b").lazy_matches_anchored(txt)
    assert len(want) == 70000 and want[:, 0].max() == 0
    for flags in (0, rrx.COMPILE_NO_DONLY):
        r = rrx.RRegex("^.*b", flags)
        assert np.array_equal(r.scan_host(txt), want), flags
        assert np.array_equal(r.scan_host(txt[:100]), want[:100]), flags           # the grown scanner is reused
        d = torch.tensor(np.frombuffer(txt, dtype=np.uint8), device="cuda")
        assert np.array_equal(r.scan_device(d.data_ptr(), len(txt)), want), (flags, "device")


def test_scanner_picks_its_kernel_variant(rrx, oracle_api, torch_cuda):
    """A thread-owned program has a lean and a straddle kernel; the scanner starts lean and switches by what the previous
    scan saw: aligned records stay lean, variable-length lines (matches straddling rows) move to straddle — and back.
    Both give the oracle's list."""
    torch = torch_cuda
    n = 40 << 20
    r = rrx.RRegex(CONFIGS[2].pattern)
    sc = rrx.Scanner(r, n, n // 48, jit=True)
    o = oracle_api.Oracle(CONFIGS[2].pattern)
    seen = []
    for key in (2, "2v", "2v", 2, 2):
        cfg = CONFIGS[key]
        d = torch.empty(n + 64, dtype=torch.uint8, device="cuda")
        rrx.corpus_device(cfg, d.data_ptr(), 0, n)
        sc.scan(d.data_ptr(), n)
        res = sc.result()
        assert res["error_bits"] == 0
        seen.append(sc.kernel_variant)
        want = o.lazy_matches(cpu_corpus(cfg, 0, 4 << 20))
        assert np.array_equal(sc.matches()[:len(want)], want), key
    assert seen == [0, 1, 1, 0, 0], seen
