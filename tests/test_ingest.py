"""CPU tests of the predict path's input/output side: the row scanner, the
vectorised collate and the dynamic batch divider against arrays produced by the reference's own
FactorGraphDataset.dag_collate_fn, the streaming DIMACS converter against the reference's CompactDimacs
(fixtures: oracle/make_golden.py cli), and the output formatter."""
import ctypes
import json
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def test_parse_ints_scanner():
    from pdp_solver_b200 import _lib
    lib = _lib.load()
    out = np.zeros(8, np.int32)
    txt = b"[3, -14,  0,-7 ,2147483647]"
    assert lib.pdp_host_parse_ints(txt, len(txt), ctypes.c_void_p(out.ctypes.data), 8) == 5
    assert out[:5].tolist() == [3, -14, 0, -7, 2147483647]
    assert lib.pdp_host_parse_ints(txt, len(txt), None, 0) == 5                 # sizing call
    assert lib.pdp_host_parse_ints(b"", 0, None, 0) == 0
    big = b"2147483648"
    assert lib.pdp_host_parse_ints(big, len(big), ctypes.c_void_p(out.ctypes.data), 8) == -2


def test_parse_row_matches_json_loads():
    from pdp_solver_b200.factorgraph.dataset import parse_row
    with open(os.path.join(GOLD, "cli_small.json")) as f:
        for line in f:
            d = json.loads(line)
            n, m, gm, ef, gf, label, misc = parse_row(line)
            assert (n, m) == tuple(d[0]) and gf is None and label == float(d[3]) and misc == d[4]
            assert gm.dtype == np.int32 and ef.dtype == np.float32
            assert np.array_equal(gm[0], np.abs(np.array(d[1])) - 1)
            assert np.array_equal(gm[1], np.array(d[2]) - 1)
            assert np.array_equal(ef, np.sign(np.array(d[1])).astype(np.float32))
    # no id list, integer label, odd spacing
    n, m, gm, ef, _, label, misc = parse_row('[[3,2],[1,-2,3,-1],[1,1,2,2],1]')
    assert (n, m, label, misc) == (3, 2, 1.0, []) and gm.tolist() == [[0, 1, 2, 0], [0, 0, 1, 1]]
    with pytest.raises(ValueError):
        parse_row('[[3,2],[1,-2,3],[1,1],1,[]]')


@pytest.mark.parametrize("tag", ["one", "split"])
def test_collate_matches_reference(tag):
    from pdp_solver_b200.factorgraph.dataset import FactorGraphDataset
    g = np.load(os.path.join(GOLD, "cli_collate.npz"))
    ds = FactorGraphDataset(os.path.join(GOLD, "cli_small.json"), limit=int(g[tag + "_limit"]), hidden_dim=3)
    batches = list(ds.batches(5000))
    assert len(batches) == 1
    gm, bvm, bfm, ef, gf, lab, misc = batches[0]
    assert len(gm) == int(g[tag + "_segments"])
    assert (tag == "one") == (len(gm) == 1)
    for s in range(len(gm)):
        assert np.array_equal(gm[s].numpy(), g["%s_%d_gm" % (tag, s)])
        assert np.array_equal(bvm[s].numpy(), g["%s_%d_bvm" % (tag, s)])
        assert np.array_equal(bfm[s].numpy(), g["%s_%d_bfm" % (tag, s)])
        assert np.array_equal(ef[s].numpy(), g["%s_%d_ef" % (tag, s)])
        assert np.array_equal(lab[s].numpy(), g["%s_%d_label" % (tag, s)])
        assert [m[0] for m in misc[s]] == g["%s_%d_ids" % (tag, s)].tolist()
        assert gf[s] is None
        assert gm[s].dtype.is_signed and str(gm[s].dtype) == "torch.int32" and tuple(ef[s].shape) == (gm[s].shape[1], 1)
    # rows are not modified by collation (the reference shifts its cached arrays in place)
    again = list(ds.batches(5000))[0]
    assert all(np.array_equal(a.numpy(), b.numpy()) for a, b in zip(gm, again[0]))


def test_batch_divider_rules():
    from pdp_solver_b200.factorgraph.dataset import DynamicBatchDivider
    d = DynamicBatchDivider(limit=3000, hidden_dim=3)
    assert d.divide_indices([100, 200, 50]) == [[0, 1, 2]]                    # 3000 // 600 = 5 >= 3
    d = DynamicBatchDivider(limit=1200, hidden_dim=3)
    # sorted descending (stable): 200(1), 100(0), 100(3), 50(2); allowed 2, then 1200//300 = 4
    assert d.divide_indices([100, 200, 50, 100]) == [[1, 0], [3, 2]]
    with pytest.raises(ValueError):                                             # the reference loops forever here
        DynamicBatchDivider(limit=100, hidden_dim=3).divide_indices([100, 10])
    assert d.divide_indices([]) == []


def test_dimacs_matches_reference():
    from pdp_solver_b200 import dimacs2json
    exp = json.load(open(os.path.join(GOLD, "dimacs_expected.json")))
    for name, e in exp.items():
        row = dimacs2json.convert_one(os.path.join(GOLD, "dimacs", name))
        got = json.loads(dimacs2json.row_to_json(row))
        assert got[0] == e[0] and got[1] == e[1] and got[2] == e[2] and got[3] == e[3] and got[4] == e[4], name
    files = dimacs2json.dimacs_files(os.path.join(GOLD, "dimacs"))
    assert sorted(os.path.basename(p) for p in files) == sorted(exp)


def test_dimacs_edge_cases(tmp_path):
    from pdp_solver_b200 import dimacs2json
    # clause split over two lines, missing final terminator, empty clause, tabs, no header
    p = tmp_path / "x.cnf"
    p.write_text("c nothing\n1 -3\n 4 0\n0\n\t-1\t2 0\n3 -4")
    nv, nc, gm, ef, _, label, misc = dimacs2json.convert_one(str(p))
    assert (nv, nc) == (4, 3) and misc == ["x.cnf"] and label == -1
    assert gm.tolist() == [[0, 2, 3, 0, 1, 2, 3], [0, 0, 0, 1, 1, 2, 2]]
    assert ef.tolist() == [1, -1, 1, -1, 1, 1, -1]
    e = tmp_path / "e.cnf"
    e.write_text("p cnf 3 0\n")
    assert dimacs2json.convert_one(str(e))[:2] == (0, 0)


def test_solution_formatter():
    from pdp_solver_b200.trainer import _bits_to_json_list
    for bits in ([], [1], [0, 1, 1, 0, 1]):
        assert _bits_to_json_list(np.array(bits, dtype=np.uint8)) == str(bits)


def test_cli_parser_and_config_merge():
    import yaml
    from pdp_solver_b200 import satyr
    a = vars(satyr.build_parser().parse_args([os.path.join(ROOT, "config", "Predict", "walksat.yaml"), "in.json", "77",
                                              "-b", "4", "-e", "0.3", "-s", "5", "-o", "out.json"]))
    cfg = satyr.make_config(yaml.safe_load(open(a["model_config"])), a)
    assert cfg["model_type"] == "walk-sat" and cfg["local_search_iteration"] == 77 and cfg["hidden_dim"] == 3
    assert cfg["batch_replication"] == 4 and cfg["epsilon"] == 0.3 and cfg["model_path"] is None
    assert cfg["batch_size"] == 5000 and cfg["test_batch_limit"] == 40000000 and cfg["dropout"] == 0


def test_parse_file_equals_parse_row(tmp_path):
    "the one-pass file scanner (pdp_host_parse_rows) against the per-row path on the fixture file and on odd rows"
    from pdp_solver_b200.factorgraph.dataset import parse_file, parse_row
    path = os.path.join(GOLD, "cli_small.json")
    rows = parse_file(path)
    lines = [l for l in open(path) if l.strip()]
    assert len(rows) == len(lines)
    for r, line in zip(rows, lines):
        q = parse_row(line)
        assert r[0] == q[0] and r[1] == q[1] and r[5] == q[5] and r[6] == q[6] and r[4] is None
        assert np.array_equal(r[2], q[2]) and np.array_equal(r[3], q[3])
        assert r[2].dtype == np.int32 and r[3].dtype == np.float32
    odd = tmp_path / "odd.json"
    odd.write_text('[[3,2],[1,-2,3,-1],[1,1,2,2],1]\n\n  [[2, 0], [], [], -1, ["empty", 7]]\n[[1,1],[ -1 ],[1],0.0,[]]')
    rows = parse_file(str(odd))
    assert [(r[0], r[1], r[5], r[6]) for r in rows] == [(3, 2, 1.0, []), (2, 0, -1.0, ["empty", 7]), (1, 1, 0.0, [])]
    assert rows[0][2].tolist() == [[0, 1, 2, 0], [0, 0, 1, 1]] and rows[0][3].tolist() == [1, -1, 1, -1]
    assert rows[1][2].shape == (2, 0) and rows[2][2].tolist() == [[0], [0]] and rows[2][3].tolist() == [-1]
    bad = tmp_path / "bad.json"
    bad.write_text('[[3,2],[1,-2,3,-1],[1,1,2,2],1]\n[[3,2],[1,-2,3],[1,1],1]\n')
    with pytest.raises(ValueError, match="line 2"):
        parse_file(str(bad))


def test_dimacs_directory_to_json_roundtrip(tmp_path):
    "convert_directory writes rows that the file scanner reads back identically to the direct DIMACS conversion"
    from pdp_solver_b200 import dimacs2json
    from pdp_solver_b200.factorgraph.dataset import parse_file
    out = tmp_path / "rows.json"
    ddir = os.path.join(GOLD, "dimacs")
    dimacs2json.convert_directory(ddir, str(out))
    rows = parse_file(str(out))
    direct = [dimacs2json.convert_one(p) for p in dimacs2json.dimacs_files(ddir)]
    assert len(rows) == len(direct) == 3
    for a, b in zip(rows, direct):
        assert a[0] == b[0] and a[1] == b[1] and a[5] == b[5] and a[6] == b[6]
        assert np.array_equal(a[2], b[2]) and np.array_equal(a[3], b[3])
