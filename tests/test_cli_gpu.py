"""GPU test of the predict command line (satyr.py) end to end: JSON / DIMACS input -> batches -> solver -> output
lines, against the reference's own output on the same file (tests/golden/cli_expected.json, made by
oracle/make_golden.py cli with the reference's trainer.predict on CPU) and against the C oracle's CNF check."""
import json
import os

import numpy as np
import pytest

from tests.conftest import GOLDEN, ROOT

pytestmark = pytest.mark.gpu


def _run_cli(tmp_path, extra, test_path=None, config="sp.yaml", iters="1000"):
    from pdp_solver_b200 import satyr
    out = str(tmp_path / "out.json")
    argv = [os.path.join(ROOT, "config", "Predict", config), test_path or os.path.join(GOLDEN, "cli_small.json"), iters,
            "-o", out] + extra
    assert satyr.main(argv) == 0
    return [l for l in open(out).read().split("\n") if l]


def _problems():
    rows = {}
    with open(os.path.join(GOLDEN, "cli_small.json")) as f:
        for line in f:
            d = json.loads(line)
            rows[d[4][0]] = d
    return rows


def _check_consistent(lines, rows):
    "every line's verdict equals the C oracle's CNF check of the solution printed in that line"
    from oracle import pdp_oracle
    from pdp_solver_b200.factorgraph.dataset import collate_segment, parse_row
    for l in lines:
        o = json.loads(l)
        d = rows[o["ID"]]
        assert o["label"] == int(d[3]) and len(o["solution"]) == d[0][0]
        gm, bvm, bfm, ef, _, _, _ = collate_segment([parse_row(json.dumps(d))])
        orc = pdp_oracle.Oracle(gm.numpy(), bvm.numpy(), bfm.numpy(), ef.numpy())
        solved, n_unsat = orc.cnf_eval(np.array(o["solution"], dtype=np.float32))
        assert int(np.asarray(solved).reshape(-1)[0]) == o["solved"]
        assert int(np.asarray(n_unsat).reshape(-1)[0]) == o["unsat_clauses"]


def test_cli_json_matches_reference_output(tmp_path):
    exp = json.load(open(os.path.join(GOLDEN, "cli_expected.json")))
    ref1 = [l for l in exp["outputs"]["seed1"].split("\n") if l]
    ref2 = [l for l in exp["outputs"]["seed2"].split("\n") if l]
    lines = _run_cli(tmp_path, ["-l", str(exp["test_batch_limit"]), "-w", "0", "-s", "1"], iters=str(exp["iterations"]))
    assert len(lines) == len(ref1)
    # same problems in the same (segment) order, same keys in the same order
    for l, r in zip(lines, ref1):
        a, b = json.loads(l), json.loads(r)
        assert list(a) == list(b) == ["ID", "label", "solved", "unsat_clauses", "solution"]
        assert a["ID"] == b["ID"] and a["label"] == b["label"] and len(a["solution"]) == len(b["solution"])
    # problems the reference decides without any random draw (identical under two seeds) must come out identical,
    # character for character
    det = [i for i, (x, y) in enumerate(zip(ref1, ref2)) if x == y]
    assert det
    for i in det:
        assert lines[i] == ref1[i]
    _check_consistent(lines, _problems())


def test_cli_walksat_and_replication(tmp_path):
    rows = _problems()
    base = _run_cli(tmp_path, ["-w", "0", "-s", "3"], iters="300")
    ws = _run_cli(tmp_path, ["-w", "400", "-e", "0.4", "-s", "3"], iters="300")
    rep = _run_cli(tmp_path, ["-w", "400", "-e", "0.4", "-s", "3", "-b", "4"], iters="300")
    only = _run_cli(tmp_path, ["-s", "3"], config="walksat.yaml", iters="400")
    for lines in (base, ws, rep, only):
        assert len(lines) == len(rows)
        _check_consistent(lines, rows)
    solved = lambda ls: sum(json.loads(l)["solved"] for l in ls)   # noqa: E731
    assert solved(ws) >= solved(base) and solved(ws) >= 3
    assert solved(rep) >= solved(base)
    assert solved(only) >= 3


def test_cli_dimacs_input(tmp_path):
    from oracle import pdp_oracle
    from pdp_solver_b200 import dimacs2json
    from pdp_solver_b200.factorgraph.dataset import collate_segment
    ddir = os.path.join(GOLDEN, "dimacs")
    lines = _run_cli(tmp_path, ["-d", "-w", "200", "-s", "1"], test_path=ddir, iters="100")
    files = dimacs2json.dimacs_files(ddir)
    assert [json.loads(l)["ID"] for l in lines] == [os.path.basename(p) for p in files]
    for l, p in zip(lines, files):
        o = json.loads(l)
        row = dimacs2json.convert_one(p)
        gm, bvm, bfm, ef, _, _, _ = collate_segment([row])
        solved, n_unsat = pdp_oracle.Oracle(gm.numpy(), bvm.numpy(), bfm.numpy(), ef.numpy()).cnf_eval(
            np.array(o["solution"], dtype=np.float32))
        assert int(np.asarray(solved).reshape(-1)[0]) == o["solved"] == 1      # all three are easy and satisfiable
        assert o["label"] == int(row[5])
    one = _run_cli(tmp_path, ["-d", "-w", "200", "-s", "1"], test_path=files[0], iters="100")
    assert len(one) == 1 and json.loads(one[0])["solved"] == 1


def test_cli_refuses_cpu_mode(tmp_path):
    with pytest.raises(RuntimeError):
        _run_cli(tmp_path, ["-c"])


def _many_rows_file(tmp_path, count, seed):
    "a compact-JSON file of `count` small random 3-SAT problems (ids p0..), written like the reference reads them"
    from pdp_solver_b200 import cnfgen
    rng = np.random.Generator(np.random.PCG64(seed))
    path = str(tmp_path / "many.json")
    with open(path, "w") as f:
        for j in range(count):
            n = int(rng.integers(30, 90))
            var, sgn = cnfgen.random_ksat(n, 3, 3.9, rng)          # [m, 3] each, clause-major
            m = var.shape[0]
            lits = [int((v + 1) * s) for v, s in zip(var.reshape(-1).tolist(), sgn.reshape(-1).tolist())]
            cls = [c + 1 for c in range(m) for _ in range(3)]
            f.write(json.dumps([[n, m], lits, cls, 0, ["p%d" % j]]) + "\n")
    return path


def test_cli_segments_are_device_independent(tmp_path):
    """The output does not depend on how the DynamicBatchDivider segments are spread over GPUs: with a small memory limit
    the input falls into many segments; `-g 1` and (when the box has them) `-g 2` give the same bytes, WalkSAT's random
    draws included, and so does a second run with `-g 1`."""
    import torch
    path = _many_rows_file(tmp_path, 60, 5)
    args = ["-l", "6000", "-w", "50", "-e", "0.5", "-s", "11"]
    one = _run_cli(tmp_path, args + ["-g", "1"], test_path=path, iters="200")
    again = _run_cli(tmp_path, args + ["-g", "1"], test_path=path, iters="200")
    assert len(one) == 60 and one == again
    assert sorted(json.loads(l)["ID"] for l in one) == sorted("p%d" % j for j in range(60))   # (the divider sorts by size)
    if torch.cuda.device_count() >= 2:
        two = _run_cli(tmp_path, args + ["-g", "2"], test_path=path, iters="200")
        assert two == one
