"""The C++ drop-in front end (include/rrx_regex.hpp): builds on CPU, the reference's own test
program compiles against it unchanged, and on a GPU it produces the oracle's matches."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "roaringregex_b200")


def _build(src, out, incs):
    cmd = ["g++", "-std=c++17", "-Wall"] + [f"-I{i}" for i in incs] + [src, f"-L{PKG}", "-lrrx", f"-Wl,-rpath,{PKG}", "-o", out]
    subprocess.run(cmd, check=True)


@pytest.fixture(scope="module")
def driver(tmp_path_factory, rrx):
    out = str(tmp_path_factory.mktemp("cpp") / "test_regex")
    _build(os.path.join(ROOT, "examples", "test_regex.cpp"), out, [os.path.join(ROOT, "include")])
    return out


def test_example_builds_and_fails_loudly_without_gpu(driver):
    import torch
    r = subprocess.run([driver], input=b"abc\n*a\n", capture_output=True)
    assert r.returncode == 2 and b"invalid expression" in r.stdout          # the reference's parser error
    if not torch.cuda.is_available():
        r = subprocess.run([driver], input=b"xxababe cde e\n(ab|cd)*e\n", capture_output=True)
        assert r.returncode == 2 and b"no CPU fallback" in r.stdout


def test_reference_test_program_compiles_unchanged(tmp_path, rrx):
    src = "/root/reference/src/test/main.cpp"
    if not os.path.exists(src):
        pytest.skip("reference tree not present on this machine")
    _build(src, str(tmp_path / "ref_main"), [os.path.join(ROOT, "include", "compat")])


@pytest.mark.gpu
def test_example_on_gpu(driver):
    r = subprocess.run([driver], input=b"xxababe cde e\n(ab|cd)*e\n", capture_output=True, check=True)
    out = r.stdout.decode()
    assert "is match? 0" in out
    assert [l for l in out.splitlines() if l.startswith("match")] == ["match [6,7): e", "match [10,11): e", "match [12,13): e"]
    r = subprocess.run([driver], input=b"ababcde\n(ab|cd)*e\n", capture_output=True, check=True)
    assert "is match? 1" in r.stdout.decode()


@pytest.mark.gpu
def test_whole_string_acceptance_matches_golden(rrx):
    import json
    acc = json.load(open(os.path.join(ROOT, "tests", "golden", "acceptance.json")))
    cache = {}
    for pat, text, want in acc[::7]:
        r = cache.get(pat) or cache.setdefault(pat, rrx.RRegex(pat))
        assert r.accepts(text.encode("latin-1")) == bool(want), (pat, text)
