"""Input pipeline of the predict path: compact-JSON rows -> batched edge lists (reference
src/pdp/factorgraph/dataset.py).  Same classes and batching rules as the reference; what changed is how the
work is done:

* a row is scanned by the library's integer scanner (`pdp_host_parse_ints`) straight into int32 arrays instead of
  `json.loads` -> Python lists -> `np.array` (dataset.py:120-136);
* a segment is collated with one concatenate + `np.repeat` of the per-problem offsets instead of the per-problem
  `np.concatenate` growth loop (dataset.py:166-185, O(B^2) copies), and the cached rows are not modified (the
  reference adds the offsets into its cached arrays in place, dataset.py:172-173);
* the tensors come out in pinned memory, ready for an asynchronous host-to-device copy.

Row format (dataset.py:120-136): `[[n, m], [+-(variable+1) per edge], [clause+1 per edge], label, [id]]`.
"""
import ctypes
import json

import numpy as np
import torch

from .. import _lib


class DynamicBatchDivider(object):
    "The dynamic batching rule of the reference (dataset.py:17-79): segments of at most limit // (E_max * hidden_dim) problems."

    def __init__(self, limit, hidden_dim):
        self.limit = limit
        self.hidden_dim = hidden_dim

    def divide_indices(self, edge_num):
        """index lists of the segments, in the reference's order: one segment in input order when everything fits,
        otherwise descending edge count (stable), cut greedily"""
        batch_size = len(edge_num)
        if batch_size == 0:
            return []
        if (self.limit // (max(edge_num) * self.hidden_dim)) >= batch_size:
            return [list(range(batch_size))]
        order = sorted(range(batch_size), reverse=True, key=lambda k: edge_num[k])
        segments, i = [], 0
        while i < batch_size:
            allowed = self.limit // (edge_num[order[i]] * self.hidden_dim)
            if allowed <= 0:
                raise ValueError("test_batch_limit too small for a problem with %d edges" % edge_num[order[i]])
            segments.append(order[i:min(i + allowed, batch_size)])
            i += allowed
        return segments

    def divide(self, variable_num, function_num, graph_map, edge_feature, graph_feature, label, misc_data):
        "reference signature (dataset.py:24-79): lists of per-segment lists"
        cols = (variable_num, function_num, graph_map, edge_feature, label, misc_data)
        segs = self.divide_indices([len(e) for e in edge_feature])
        out = [[[c[j] for j in ind] for ind in segs] for c in cols]
        gf = [[None] if graph_feature[0] is None else [graph_feature[j] for j in ind] for ind in segs]
        return out[0], out[1], out[2], out[3], gf, out[4], out[5]


def parse_row(text):
    """One compact-JSON row -> (variable_num, function_num, graph_map int32[2,E], edge_feature f32[E], None, label,
    misc) exactly as the reference's `_convert_line` (dataset.py:120-136)."""
    if isinstance(text, str):
        text = text.encode()
    lib = _lib.load()
    # the two integer lists are the 2nd and 3rd bracketed groups: "[[n, m], [..], [..], label, [id]]"
    h1 = text.index(b"]")
    a0 = text.index(b"[", h1)
    a1 = text.index(b"]", a0)
    b0 = text.index(b"[", a1)
    b1 = text.index(b"]", b0)
    head = json.loads(text[text.index(b"[") + 1:h1 + 1].decode())
    variable_num, function_num = int(head[0]), int(head[1])
    tail = json.loads("[" + text[b1 + 1:].decode().lstrip().lstrip(","))   # label, [id]] -> [label, [id]]
    label = float(tail[0])
    misc = tail[1] if len(tail) > 1 else []

    def ints(lo, hi):
        n_max = (hi - lo) // 2 + 1
        out = np.empty(n_max, dtype=np.int32)
        seg = text[lo:hi]
        n = lib.pdp_host_parse_ints(seg, len(seg), ctypes.c_void_p(out.ctypes.data), n_max)
        if n < 0 or n > n_max:
            raise ValueError("malformed integer list in input row")
        return out[:n]

    lit = ints(a0 + 1, a1)
    cls = ints(b0 + 1, b1)
    if lit.shape[0] != cls.shape[0]:
        raise ValueError("input row: %d literals but %d clause indices" % (lit.shape[0], cls.shape[0]))
    graph_map = np.stack((np.abs(lit) - 1, np.abs(cls) - 1))
    edge_feature = np.sign(lit).astype(np.float32)
    return (variable_num, function_num, graph_map, edge_feature, None, label, misc)


def parse_file(path):
    """Every row of a compact-JSON file in one pass of the library's scanner (`pdp_host_parse_rows`): a list of the tuples
    `parse_row` returns.  The per-row arrays are views into four file-wide arrays (no per-row allocation)."""
    lib = _lib.load()
    with open(path, "rb") as fh:
        data = fh.read()
    int_cap = len(data) // 2 + 1
    row_cap = len(data) // 16 + 2          # the shortest row has 19 bytes
    lits = np.empty(int_cap, dtype=np.int32)
    cls = np.empty(int_cap, dtype=np.int32)
    row_ptr = np.empty(row_cap + 1, dtype=np.int64)
    nm = np.empty(2 * row_cap, dtype=np.int32)
    label = np.empty(row_cap, dtype=np.float64)
    tail = np.empty(2 * row_cap, dtype=np.int64)
    P = lambda a: ctypes.c_void_p(a.ctypes.data)      # noqa: E731
    n = lib.pdp_host_parse_rows(data, len(data), P(lits), P(cls), int_cap, P(row_ptr), P(nm), P(label), P(tail), row_cap)
    if n < 0:
        raise ValueError("%s: malformed row at line %d" % (path, -n) if n > -(1 << 61) else "%s: parser capacity" % path)
    total = int(row_ptr[n])
    lits, cls = lits[:total], cls[:total]
    gm = np.empty((2, total), dtype=np.int32)
    np.abs(lits, out=gm[0])
    gm[0] -= 1
    np.abs(cls, out=gm[1])
    gm[1] -= 1
    ef = np.sign(lits).astype(np.float32)
    rows = []
    for r in range(n):
        a, b = int(row_ptr[r]), int(row_ptr[r + 1])
        t = data[int(tail[2 * r]):int(tail[2 * r + 1])].strip()
        t = t[:-1].strip() if t.endswith(b"]") else t              # the row's own closing bracket
        misc = json.loads(t.decode()) if t else []
        rows.append((int(nm[2 * r]), int(nm[2 * r + 1]), gm[:, a:b], ef[a:b], None, float(label[r]), misc))
    return rows


def collate_segment(rows, pin=False):
    """One segment -> (graph_map int32[2,E], batch_variable_map int32[V], batch_function_map int32[F],
    edge_feature f32[E,1], None, label f32[B,1], misc list) with the reference's numbering
    (dataset.py:166-185): problem j's variables / clauses are shifted by the totals of problems 0..j-1."""
    B = len(rows)
    vn = np.fromiter((r[0] for r in rows), dtype=np.int64, count=B)
    fn = np.fromiter((r[1] for r in rows), dtype=np.int64, count=B)
    en = np.fromiter((r[2].shape[1] for r in rows), dtype=np.int64, count=B)
    E, V, F = int(en.sum()), int(vn.sum()), int(fn.sum())
    if max(E, V, F) >= 2 ** 31:
        raise ValueError("segment exceeds int32 indexing")
    voff = np.concatenate(([0], np.cumsum(vn)[:-1])).astype(np.int32)
    foff = np.concatenate(([0], np.cumsum(fn)[:-1])).astype(np.int32)

    def buf(shape, dtype):
        t = torch.empty(shape, dtype=dtype)
        return t.pin_memory() if pin and torch.cuda.is_available() else t

    gm = buf((2, E), torch.int32)
    ef = buf((E, 1), torch.float32)
    g, f = gm.numpy(), ef.numpy()
    if B:
        np.concatenate([r[2] for r in rows], axis=1, out=g)
        g[0] += np.repeat(voff, en)
        g[1] += np.repeat(foff, en)
        np.concatenate([r[3] for r in rows], out=f[:, 0])
    bvm = buf((V,), torch.int32)
    bfm = buf((F,), torch.int32)
    pid = np.arange(B, dtype=np.int32)
    bvm.numpy()[:] = np.repeat(pid, vn)
    bfm.numpy()[:] = np.repeat(pid, fn)
    label = torch.from_numpy(np.array([[r[5]] for r in rows], dtype=np.float32).reshape(B, 1))
    return gm, bvm, bfm, ef, None, label, [r[6] for r in rows]


class FactorGraphDataset(object):
    "Reads CNFs in the compact JSON format (reference dataset.py:85-187), one row per line."

    def __init__(self, input_file, limit, hidden_dim, max_cache_size=100000, generator=None, epoch_size=0,
                 batch_replication=1, rows=None):
        self._input_file = input_file
        # already-parsed rows (DIMACS input), or the whole file scanned once (the reference reads the whole file just to
        # count its rows, dataset.py:96-97, then parses each row with json.loads)
        self._rows = rows if rows is not None else parse_file(input_file)
        self._offsets = None
        self.batch_divider = DynamicBatchDivider(limit // batch_replication, hidden_dim)

    def __len__(self):
        return len(self._rows) if self._rows is not None else len(self._offsets)

    def __getitem__(self, idx):
        if self._rows is not None:
            return self._rows[idx]
        pos, n = self._offsets[idx]
        with open(self._input_file, "rb") as fh:
            fh.seek(pos)
            return parse_row(fh.read(n))

    _convert_line = staticmethod(parse_row)

    def dag_collate_fn(self, input_data, pin=False):
        """reference dataset.py:138-187: a list of rows -> per-segment lists
        (graph_map, batch_variable_map, batch_function_map, edge_feature, graph_feat, label, misc_data)"""
        segs = self.batch_divider.divide_indices([r[2].shape[1] for r in input_data])
        cols = [collate_segment([input_data[j] for j in ind], pin) for ind in segs]
        return tuple([c[k] for c in cols] for k in range(7))

    def segments(self, batch_size, pin=False):
        """The same segments as `batches`, one at a time and collated only when asked for: the multi-GPU predict path
        hands them to the devices while later ones are still being built."""
        n = len(self)
        fh = None if self._rows is not None else open(self._input_file, "rb")
        try:
            for lo in range(0, n, batch_size):
                if self._rows is not None:
                    rows = self._rows[lo:lo + batch_size]
                else:
                    rows = []
                    for pos, size in self._offsets[lo:lo + batch_size]:
                        fh.seek(pos)
                        rows.append(parse_row(fh.read(size)))
                for ind in self.batch_divider.divide_indices([r[2].shape[1] for r in rows]):
                    yield collate_segment([rows[j] for j in ind], pin)
        finally:
            if fh is not None:
                fh.close()

    def batches(self, batch_size, pin=False):
        "what the reference's DataLoader(batch_size, shuffle=False, collate_fn=dag_collate_fn) yields"
        n = len(self)
        if self._rows is not None:
            for lo in range(0, n, batch_size):
                yield self.dag_collate_fn(self._rows[lo:lo + batch_size], pin)
            return
        with open(self._input_file, "rb") as fh:          # one handle for the whole pass
            for lo in range(0, n, batch_size):
                rows = []
                for pos, size in self._offsets[lo:lo + batch_size]:
                    fh.seek(pos)
                    rows.append(parse_row(fh.read(size)))
                yield self.dag_collate_fn(rows, pin)
