"""Sharding of a batch of CNFs across the GPUs of one box.

Problems of a batch never interact (the reference concatenates them into one block-diagonal factor graph,
reference src/pdp/factorgraph/dataset.py:138-187), so the unit of distribution is the problem: every rank
solves its own sub-batch with the very same kernels and NO data-path collective; the only communication
is the final gather of the assignments and verdicts (torch.distributed, NCCL on the box, gloo in the CPU
tests).  A problem's result does not depend on its batch mates (tests/test_gpu_parity.py::
test_batch_composition_invariance), so sharded == unsharded bit for bit.

The reference's only multi-GPU notion is nn.DataParallel (reference src/pdp/factorgraph/base.py:96-97),
which would split graph_map[2,E] along dim 0 and is functionally single-GPU.
"""
import numpy as np


def problem_sizes(graph_map, batch_variable_map, batch_function_map, batch_size=None):
    """(variables, clauses, edges) per problem of a batch in the reference's layout (numpy int64 arrays)."""
    bvm = np.asarray(batch_variable_map).astype(np.int64)
    bfm = np.asarray(batch_function_map).astype(np.int64)
    B = int(batch_size) if batch_size is not None else (int(bvm.max()) + 1 if bvm.size else 0)
    nv = np.bincount(bvm, minlength=B)
    nf = np.bincount(bfm, minlength=B)
    ne = np.bincount(bvm[np.asarray(graph_map)[0].astype(np.int64)], minlength=B) if np.asarray(graph_map).size else np.zeros(B, np.int64)
    return nv, nf, ne


def lpt_assign(costs, world_size):
    """Longest-processing-time-first assignment of problems to ranks by cost (edge count).  Returns one
    ascending list of problem ids per rank; deterministic (ties by problem id)."""
    costs = np.asarray(costs, dtype=np.int64)
    order = sorted(range(costs.size), key=lambda j: (-int(costs[j]), j))
    load = [0] * world_size
    parts = [[] for _ in range(world_size)]
    for j in order:
        r = min(range(world_size), key=lambda q: (load[q], q))
        parts[r].append(j)
        load[r] += int(costs[j]) + 1
    return [sorted(p) for p in parts]


def extract_problems(batch, problem_ids, batch_size=None):
    """Sub-batch holding `problem_ids` (ascending) of `batch` = (graph_map[2,E], batch_variable_map[V],
    batch_function_map[F], edge_feature[E,1]) with variables, clauses and problems re-numbered from 0.
    Returns (sub_batch, variable_index) where variable_index[i] is the batch-global index of the sub-batch's
    variable i (to scatter assignments back)."""
    gm, bvm, bfm, ef = [np.asarray(x) for x in batch]
    B = int(batch_size) if batch_size is not None else (int(bvm.max()) + 1 if bvm.size else 0)
    ids = np.asarray(problem_ids, dtype=np.int64)
    new_id = np.full(B + 1, -1, dtype=np.int64)
    new_id[ids] = np.arange(ids.size)
    vsel = np.nonzero(new_id[bvm] >= 0)[0]
    fsel = np.nonzero(new_id[bfm] >= 0)[0]
    vmap = np.full(bvm.size + 1, -1, dtype=np.int64)
    vmap[vsel] = np.arange(vsel.size)
    fmap = np.full(bfm.size + 1, -1, dtype=np.int64)
    fmap[fsel] = np.arange(fsel.size)
    esel = np.nonzero(vmap[gm[0]] >= 0)[0] if gm.size else np.zeros(0, np.int64)
    sub_gm = np.stack([vmap[gm[0][esel]], fmap[gm[1][esel]]]).astype(np.int32) if gm.size else np.zeros((2, 0), np.int32)
    sub = (sub_gm, new_id[bvm[vsel]].astype(np.int32), new_id[bfm[fsel]].astype(np.int32),
           np.ascontiguousarray(np.asarray(ef).reshape(-1, 1)[esel]).astype(np.float32))
    return sub, vsel


def solve_sharded(batch, solve_fn, rank, world_size, dist=None, batch_size=None):
    """Every rank solves its LPT share of `batch` with `solve_fn(sub_batch) -> (prediction[V_sub] float32,
    solved[B_sub] float32)`; the full-batch (prediction[V], solved[B]) is assembled on every rank with one
    all_gather_object at the end (a few bytes per variable).  `dist` = torch.distributed (initialised) or
    None for a single process."""
    gm, bvm, bfm, ef = [np.asarray(x) for x in batch]
    B = int(batch_size) if batch_size is not None else (int(bvm.max()) + 1 if bvm.size else 0)
    _, _, ne = problem_sizes(gm, bvm, bfm, B)
    parts = lpt_assign(ne, world_size)
    mine = parts[rank]
    sub, vsel = extract_problems((gm, bvm, bfm, ef), mine, B)
    if len(mine):
        pred, solved = solve_fn(sub)
        pred = np.asarray(pred, dtype=np.float32).reshape(-1)
        solved = np.asarray(solved, dtype=np.float32).reshape(-1)
    else:
        pred, solved = np.zeros(0, np.float32), np.zeros(0, np.float32)
    piece = (np.asarray(mine, dtype=np.int64), vsel, pred, solved)
    if dist is not None and world_size > 1:
        pieces = [None] * world_size
        dist.all_gather_object(pieces, piece)
    else:
        pieces = [piece]
    full_pred = np.zeros(bvm.size, dtype=np.float32)
    full_solved = np.zeros(B, dtype=np.float32)
    for ids, vs, p, s in pieces:
        full_pred[vs] = p
        full_solved[ids] = s
    return full_pred, full_solved
