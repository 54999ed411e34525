"""Vectorised synthetic CNF batches in the reference's batch-tensor layout.

Distribution follows the reference's UniformCNFGenerator (reference src/pdp/generator.py:107-125):
m = int(n * alpha) clauses, each clause = k DISTINCT variables uniform over n, each literal sign
i.i.d. +-1 with probability 1/2, edges emitted clause-major.  The reference generator costs O(n) per
clause and is unusable at n = 1M, so this is a same-distribution numpy.Generator implementation, not
a stream-identical one; the same tensors are fed to every implementation being compared.

Batch layout (reference src/pdp/factorgraph/dataset.py:138-187, dag_collate_fn):
  graph_map          int32 [2, E]   row 0 = batch-global variable index, row 1 = batch-global clause index
  batch_variable_map int32 [V]      problem id of each variable (non-decreasing)
  batch_function_map int32 [F]      problem id of each clause
  edge_feature       float32 [E, 1] literal sign, +1 / -1
"""
import numpy as np


def random_ksat(n, k, alpha, rng, m=None):
    """One uniform random k-SAT instance -> (var[m,k] int32, sign[m,k] float32)."""
    if m is None:
        m = int(n * alpha)
    var = rng.integers(0, n, size=(m, k), dtype=np.int64)
    if k > 1:
        # redraw rows holding a repeated variable until every clause has k distinct variables
        while True:
            s = np.sort(var, axis=1)
            bad = np.nonzero((s[:, 1:] == s[:, :-1]).any(axis=1))[0]
            if bad.size == 0:
                break
            var[bad] = rng.integers(0, n, size=(bad.size, k), dtype=np.int64)
    sign = (rng.integers(0, 2, size=(m, k), dtype=np.int64) * 2 - 1).astype(np.float32)
    return var.astype(np.int32), sign


def collate(problems):
    """problems: list of (n, var[m,k_i...] , sign) with clause-major flattened edges, or of
    (n, clause_lists) -- see `from_clauses`.  Returns the four numpy batch tensors."""
    gm0, gm1, ef, bvm, bfm = [], [], [], [], []
    voff = foff = 0
    for b, (n, var, sign) in enumerate(problems):
        m, k = var.shape
        gm0.append((var.reshape(-1) + voff).astype(np.int32))
        gm1.append((np.repeat(np.arange(m, dtype=np.int32), k) + foff).astype(np.int32))
        ef.append(sign.reshape(-1).astype(np.float32))
        bvm.append(np.full(n, b, dtype=np.int32))
        bfm.append(np.full(m, b, dtype=np.int32))
        voff += n
        foff += m
    graph_map = np.stack([np.concatenate(gm0), np.concatenate(gm1)]).astype(np.int32)
    return (graph_map, np.concatenate(bvm), np.concatenate(bfm),
            np.concatenate(ef).reshape(-1, 1).astype(np.float32))


def from_clauses(problem_list):
    """problem_list: list of (n, clauses) where clauses is a list of lists of signed 1-based
    literals (DIMACS style, ragged clause lengths allowed, empty problems allowed).
    Returns the four numpy batch tensors."""
    gm0, gm1, ef, bvm, bfm = [], [], [], [], []
    voff = foff = 0
    for b, (n, clauses) in enumerate(problem_list):
        for a, cl in enumerate(clauses):
            for lit in cl:
                gm0.append(abs(lit) - 1 + voff)
                gm1.append(a + foff)
                ef.append(1.0 if lit > 0 else -1.0)
        bvm += [b] * n
        bfm += [b] * len(clauses)
        voff += n
        foff += len(clauses)
    graph_map = np.array([gm0, gm1], dtype=np.int32).reshape(2, -1)
    return (graph_map, np.array(bvm, dtype=np.int32), np.array(bfm, dtype=np.int32),
            np.array(ef, dtype=np.float32).reshape(-1, 1))


def random_batch(batch, n, k, alpha, seed):
    """`batch` uniform random k-SAT problems of n variables -> the four numpy batch tensors."""
    rng = np.random.Generator(np.random.PCG64(seed))
    return collate([(n,) + random_ksat(n, k, alpha, rng) for _ in range(batch)])


def mixed_batch(specs, seed):
    """specs: list of (n, k, alpha) -> one batch with heterogeneous problems."""
    rng = np.random.Generator(np.random.PCG64(seed))
    return collate([(n,) + random_ksat(n, k, alpha, rng) for (n, k, alpha) in specs])


def to_json_line(n, var, sign, label=0, pid=None):
    """One problem in the reference's JSON-lines format (dataset.py:120-136 reads it):
    [[n, m], [signed 1-based variable per edge], [1-based clause per edge], label, [id]]."""
    m, k = var.shape
    lits = ((var.reshape(-1).astype(np.int64) + 1) * sign.reshape(-1).astype(np.int64)).tolist()
    cls = (np.repeat(np.arange(m, dtype=np.int64), k) + 1).tolist()
    return str([[int(n), int(m)], [int(x) for x in lits], [int(x) for x in cls], int(label),
                [pid] if pid is not None else []]).replace("'", '"')
