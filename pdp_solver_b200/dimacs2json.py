#!/usr/bin/env python3
"""DIMACS CNF -> the compact row format of the predict path (reference src/dimacs2json.py).

The reference builds a dense [clauses, variables] int32 matrix per file (dimacs2json.py:24-50), which is
O(n*m) memory (4 * 4.2e12 bytes at n = 1 M).  Here the file is scanned once by the library's streaming scanner
(`pdp_host_parse_dimacs`) and the reference's normalisation is applied to the sparse literal list:

* within a clause a variable appears once, with the sign of its LAST occurrence (the dense assignment
  `mat[j, |v|-1] = sign(v)` of dimacs2json.py:42-44);
* empty clauses and variables that occur in no clause are dropped and the survivors renumbered in ascending order
  (dimacs2json.py:46-50);
* edges are emitted clause-major with ascending variable inside a clause (`np.nonzero` order, dimacs2json.py:93-96).

Clauses are delimited by their `0` terminator (standard DIMACS); the reference takes one clause per line and drops
the line's last token, which is the same thing on well-formed files.  `-s/--simplify` (subsumption by dense
matrix products, dimacs2json.py:55-84) is not offered here.
"""
import argparse
import ctypes
import sys
from os import listdir
from os.path import isfile, join, split, splitext

import numpy as np

from . import _lib


def parse_dimacs_bytes(data):
    """-> (variable_num, clause_num, lit int32[E] = +-(variable+1), cls int32[E] = clause+1)"""
    lib = _lib.load()
    cap = len(data) // 2 + 2
    lits = np.empty(cap, dtype=np.int32)
    info = (ctypes.c_int64 * 4)()
    _lib.check(lib.pdp_host_parse_dimacs(data, len(data), ctypes.c_void_p(lits.ctypes.data), cap, ctypes.byref(info)),
               "pdp_host_parse_dimacs")
    lits = lits[:info[2]]
    is_end = lits == 0
    clause = np.cumsum(is_end) - is_end            # clause index of every entry
    keep = ~is_end
    lits, clause = lits[keep], clause[keep].astype(np.int64)
    if lits.size == 0:
        return 0, 0, np.zeros(0, np.int32), np.zeros(0, np.int32)
    var = np.abs(lits).astype(np.int64) - 1
    nv = int(var.max()) + 1
    key = clause * nv + var
    # last occurrence of every (clause, variable) pair, ascending key = clause-major, variable ascending
    _, first_rev = np.unique(key[::-1], return_index=True)
    sel = key.size - 1 - first_rev
    lits, clause, var = lits[sel], clause[sel], var[sel]
    _, cid = np.unique(clause, return_inverse=True)  # drops empty clauses
    vuniq, vid = np.unique(var, return_inverse=True)  # drops unused variables
    out_lit = (np.sign(lits) * (vid + 1)).astype(np.int32)
    return int(vuniq.size), int(cid.max()) + 1, out_lit, (cid + 1).astype(np.int32)


def file_label(path, single_file):
    "the label rules of dimacs2json.py:111-112 (directory) and :125-129 (single file)"
    if single_file:
        if len(path) < 8:
            return -1
        c = path[-8]
    else:
        c = splitext(path)[0][-1]
    return float(c) if c.isdigit() else -1


def convert_one(path, single_file=False):
    "one DIMACS file -> a parsed row (variable_num, function_num, graph_map, edge_feature, None, label, [file name])"
    with open(path, "rb") as fh:
        nv, nc, lit, cls = parse_dimacs_bytes(fh.read())
    graph_map = np.stack((np.abs(lit) - 1, cls - 1)).astype(np.int32)
    return (nv, nc, graph_map, np.sign(lit).astype(np.float32), None, float(file_label(path, single_file)),
            [split(path)[1]])


def dimacs_files(dimacs_dir):
    "the files convert_directory (dimacs2json.py:99-118) would convert, in its order"
    out = []
    for f in listdir(dimacs_dir):
        p = join(dimacs_dir, f)
        if isfile(p) and splitext(p)[1].lower() in (".dimacs", ".cnf"):
            out.append(p)
    return out


def row_to_json(row):
    "the row as one line of the compact JSON format (numpy-2 safe: plain ints)"
    nv, nc, gm, ef, _, label, misc = row
    lit = ((gm[0].astype(np.int64) + 1) * ef.astype(np.int64)).tolist()
    cls = (gm[1].astype(np.int64) + 1).tolist()
    return str([[int(nv), int(nc)], lit, cls, label, list(misc)]).replace("'", '"')


def convert_directory(dimacs_dir, output_file, propagate=False, only_positive=False):
    if propagate:
        raise NotImplementedError("--simplify (dense subsumption) is not offered; see module docstring")
    with open(output_file, "w") as f:
        for p in dimacs_files(dimacs_dir):
            row = convert_one(p)
            if only_positive and row[5] == 0:
                continue
            f.write(row_to_json(row) + "\n")


def convert_file(file_name, output_file, propagate=False):
    if propagate:
        raise NotImplementedError("--simplify (dense subsumption) is not offered; see module docstring")
    with open(output_file, "w") as f:
        f.write(row_to_json(convert_one(file_name, single_file=True)) + "\n")


if __name__ == "__main__":
    parser = argparse.ArgumentParser()
    parser.add_argument("in_dir", action="store", type=str)
    parser.add_argument("out_file", action="store", type=str)
    parser.add_argument("-s", "--simplify", help="Propagate binary constraints", action="store_true", default=False)
    parser.add_argument("-p", "--positive", help="Output only positive examples", action="store_true", default=False)
    a = vars(parser.parse_args())
    convert_directory(a["in_dir"], a["out_file"], a["simplify"], a["positive"])
    sys.exit(0)
