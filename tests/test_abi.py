"""CPU-side checks of the drop-in boundary: libbbenv.so loads and exports every symbol include/bbenv.h declares,
struct layouts agree, the host-side ideal grammar works, and the product path does not reach into oracle/."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_exported():
    from deepgroebner_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        pytest.skip("libbbenv.so not built (run __graft_entry__.build())")
    hdr = open(os.path.join(ROOT, "include", "bbenv.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(bb_[a-z_0-9]+)\s*\(", hdr))
    assert len(declared) >= 20
    lib = C.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), "symbol %s declared in bbenv.h is not exported" % name
    assert declared == set(_lib.EXPORTS)
    assert lib.bb_abi_version() == _lib.BB_ABI_VERSION


def test_struct_layouts_and_hash():
    from deepgroebner_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        pytest.skip("libbbenv.so not built")
    lib = _lib.load()
    assert C.sizeof(_lib.BBEpisodeStats) == 72 == np.dtype(_lib.STATS_DTYPE).itemsize
    assert C.sizeof(_lib.BBConfig) == 16 * 4 and C.sizeof(_lib.BBCounters) == 12 * 8
    from hashing import hash_item
    for x, pos in ((0, 0), (123456789, 7), ((1 << 63) + 5, 1 << 40)):
        assert int(lib.bb_hash_item(x, pos)) == int(hash_item(x, pos))


def test_create_fails_loudly_without_gpu():
    import torch
    from deepgroebner_b200 import _lib
    if torch.cuda.is_available() or not os.path.exists(_lib.LIB_PATH):
        pytest.skip("only meaningful on a CPU-only box with the library built")
    from deepgroebner_b200.buchberger import BuchbergerEngine
    with pytest.raises(_lib.BBError):
        BuchbergerEngine("3-20-10-weighted")
    lib = _lib.load()
    cfg = _lib.BBConfig(abi_version=_lib.BB_ABI_VERSION, device=0, nvars=3, k=1, prime=32003, elimination=0, rewards=0, sort_input=0,
                        sort_reducers=1, num_envs=1, max_basis=64, max_pairs=64, max_terms=128, max_poly_terms=16,
                        max_gens=10, max_gen_terms=20)
    h = C.c_void_p()
    assert lib.bb_create(C.byref(cfg), C.byref(h)) < 0
    assert b"no CUDA device" in lib.bb_last_error(None)
    cfg.prime = 32004
    assert lib.bb_create(C.byref(cfg), C.byref(h)) < 0 and b"prime" in lib.bb_last_error(None)


def test_parse_ideal_dist_grammar():
    from deepgroebner_b200.ideals import BinomialSpec, FixedIdealGenerator, cyclic, parse_ideal_dist
    s = parse_ideal_dist("3-20-10-weighted")
    assert s == BinomialSpec(3, 20, 10, "weighted", False, False, False)
    s = parse_ideal_dist("5-5-10-uniform-consts-homog-pure")
    assert (s.n, s.d, s.s, s.dist, s.constants, s.homogeneous, s.pure) == (5, 5, 10, "uniform", True, True, True)
    g = parse_ideal_dist("cyclic-4")
    assert isinstance(g, FixedIdealGenerator) and g.nvars() == 4 and len(g.F) == 4
    q = parse_ideal_dist("3-20-10-0.5-uniform-homog")   # RandomIdealGenerator, ideals.cpp:131-141
    assert (q.n, q.d, q.s, q.lam, q.dist, q.constants, q.homogeneous) == (3, 20, 10, 0.5, "uniform", False, True)
    assert q.max_gen_terms() >= 10 * (2 + 8)
    with pytest.raises(ValueError):
        parse_ideal_dist("banana")


def test_cyclic_matches_oracle(port):
    from deepgroebner_b200.ideals import cyclic
    for n in (3, 4, 6):
        mine = [sorted((c, tuple(e) + (0,) * (8 - n)) for c, e in f) for f in cyclic(n)]
        assert mine == [sorted(f) for f in port.cyclic(n)]


def test_product_path_never_touches_oracle():
    """The oracle is test infrastructure: nothing under deepgroebner_b200/ may import, link or execute it."""
    pkg = os.path.join(ROOT, "deepgroebner_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", "Makefile")):
                text = open(os.path.join(dirpath, f)).read()
                code = "\n".join(l for l in text.splitlines() if "oracle" in l and not l.strip().startswith(("#", "//", '"', "*", "there", "The", "test")))
                assert "import oracle" not in text and "from oracle" not in text and "bb_oracle" not in text \
                    and "libdgref" not in text, os.path.join(dirpath, f)
