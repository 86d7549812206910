"""SURVEY 8(f) row 4 on the GPU: deepgroebner_b200.strat.make_strat writes the same CSV bytes as the reference's
scripts/make_strat.cpp binary (oracle/_ref/make_strat, compiled from the unmodified sources) on the same ideal file."""
import os
import subprocess

import numpy as np
import pytest

from helpers import best_oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch


def reference_csv(orc, root, dist, strategy, seed):
    """The reference binary's output file (it reads/writes data/stats/... relative to its working directory), or -- when
    the binary did not travel -- the same rows from the oracle's buchberger()."""
    from oracle import oracle as O
    from deepgroebner_b200.strat import HEADER, read_ideal_file
    name = "%s_%s%s.csv" % (dist, strategy, "_%d" % seed if (seed is not None and strategy == "random") else "")
    if os.path.exists(O.REF_MAKE_STRAT):
        ref_root = os.path.join(root, "ref")
        os.makedirs(os.path.join(ref_root, "data", "stats", dist), exist_ok=True)
        src = os.path.join(root, "data", "stats", dist, dist + ".csv")
        dst = os.path.join(ref_root, "data", "stats", dist, dist + ".csv")
        if not os.path.exists(dst):
            with open(src) as a, open(dst, "w") as b:
                b.write(a.read())
        subprocess.run([O.REF_MAKE_STRAT, dist, strategy] + ([str(seed)] if seed is not None else []), cwd=ref_root, check=True)
        return open(os.path.join(ref_root, "data", "stats", dist, name), "rb").read()
    rows = [HEADER]
    for F in read_ideal_file(os.path.join(root, "data", "stats", dist, dist + ".csv")):
        _, st = orc.buchberger(F, selection=strategy, gamma=0.99, seed=seed or 0)
        rows.append("%d,%d,%d" % (st["zero_reductions"], st["nonzero_reductions"], st["polynomial_additions"]))
    return ("\n".join(rows) + "\n").encode()


@pytest.mark.parametrize("dist,count", [("3-20-10-weighted", 400), ("5-5-10-uniform", 200), ("3-6-5-0.5-uniform", 100)])
def test_make_strat_csv_is_byte_identical(torch_cuda, tmp_path, dist, count):
    from deepgroebner_b200.strat import make_strat, write_ideal_file
    orc = best_oracle()
    gen = orc.generator(dist)
    gen.seed(2024)
    ideals = [gen.next() for _ in range(count)]
    root = str(tmp_path)
    stats_root = os.path.join(root, "data", "stats")
    write_ideal_file(os.path.join(stats_root, dist, dist + ".csv"), ideals)
    for strategy, seed in (("degree", None), ("normal", None), ("sugar", None), ("first", None), ("random", 7)):
        code, out = make_strat(dist, strategy, seed, root=stats_root)
        assert code == 0, out
        assert open(out, "rb").read() == reference_csv(orc, root, dist, strategy, seed), (dist, strategy)
    # the reference refuses to overwrite (make_strat.cpp:44-48) and reports a missing distribution (:35-38)
    assert make_strat(dist, "degree", None, root=stats_root)[0] == 3
    assert make_strat("no-such-dist", "degree", None, root=stats_root)[0] == 2
