"""Host side of the tensor-core edge layers (csrc/pdp_edge_nn.cu): the weight images the kernels stream, and the two calls.

A layer's weights are prepared once (and again whenever a parameter changes): split into the two tf32 terms
(hi = W truncated to tf32, lo = W - hi truncated to tf32), padded, cut into K-chunks of 16 and stored chunk after chunk
exactly as the kernel wants them in shared memory -- per chunk {hi, lo} x [n_tot rows][64 bytes, 16-byte units swizzled]
-- so that one bulk copy per chunk brings them in.  The nn.Linear / nn.GRUCell modules stay the owners of the parameters
(state-dict compatible with the reference); these objects only cache images of them."""
import ctypes
import os

import torch

from .. import _lib

ACT_NONE, ACT_LOGSIGMOID = 0, 1


def _tf32_split(w):
    bits = w.contiguous().view(torch.int32)
    hi = (bits & -8192).view(torch.float32)                        # 0xffffe000: the top 19 bits
    lo = ((w - hi).contiguous().view(torch.int32) & -8192).view(torch.float32)
    return hi, lo


def chunk_k():
    "K elements per chunk of the weight image: the kernel's compile-time constant"
    return int(_lib.load().pdp_edge_nn_chunk_k())


def _image(wp, passes, n_tot):
    """wp: [passes * n_tot, k_pad] -> the shared-memory image of every (pass, K chunk): {hi, lo} x one operand tile in the
    tensor core's K-major 64-byte-swizzled layout: [n_tot rows][4 units of 16 bytes][4 floats] with unit g of row n stored at
    unit g ^ ((n >> 1) & 3)."""
    k_pad = wp.shape[1]
    ck = chunk_k()
    chunks = k_pad // ck
    hi, lo = _tf32_split(wp)
    both = torch.stack((hi, lo), 0)                                 # [2, P * n_tot, k_pad]
    both = both.view(2, passes, n_tot, chunks, ck // 4, 4)          # [term, pass, row, chunk, group, j]
    if not int(_lib.load().pdp_edge_nn_swizzle()):
        raise _lib.PdpError("this library build expects un-swizzled weight images")
    rows = torch.arange(n_tot, device=wp.device)
    slot = torch.arange(ck // 4, device=wp.device).view(1, -1) ^ ((rows >> 1) & 3).view(-1, 1)      # [row, group] -> unit
    src = torch.empty_like(slot)
    src.scatter_(1, slot, torch.arange(ck // 4, device=wp.device).view(1, -1).expand(n_tot, -1))    # unit -> group stored there
    idx = src.view(1, 1, n_tot, 1, ck // 4, 1).expand(2, passes, n_tot, chunks, ck // 4, 4)
    both = torch.gather(both, 4, idx)                               # [term, pass, row, chunk, unit, j]
    return both.permute(1, 3, 0, 2, 4, 5).contiguous()             # [pass, chunk, term, row, unit, j]


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None and t.numel() else ctypes.c_void_p(0)


def _round_up(x, m):
    return (x + m - 1) // m * m


def _versions(params):
    return tuple((p.data_ptr(), p._version, tuple(p.shape)) for p in params if p is not None)


def use_tensor_cores():
    "PDP_B200_NN=torch keeps the dense layers on the library GEMMs (A/B comparisons in the tests and the profiler)"
    return os.environ.get("PDP_B200_NN", "tcgen05") != "torch"


class TensorLinear(object):
    "out = act([x1 | x2 | x3] W^T + b) * mask for an nn.Linear (weight [N, K], bias [N] or None)"

    def __init__(self, linear):
        self._linear = linear
        self._key = None

    def _prepare(self):
        lin = self._linear
        key = _versions((lin.weight, lin.bias))
        if key == self._key:
            return
        w = lin.weight.detach().to(torch.float32)
        n, k = w.shape
        if n > 256:
            raise _lib.PdpError("TensorLinear: %d outputs do not fit one accumulator block" % n)
        self.n_out, self.k = n, k
        self.n_blk, self.n_mma, self.passes = _round_up(n, 16), 1, 1
        n_tot = self.n_blk
        wp = torch.zeros(n_tot, _round_up(k, chunk_k()), dtype=torch.float32, device=w.device)
        wp[:n, :k] = w
        self.image = _image(wp, 1, n_tot)
        self.bias = torch.zeros(n_tot, dtype=torch.float32, device=w.device)
        if lin.bias is not None:
            self.bias[:n] = lin.bias.detach()
        self._key = key

    def __call__(self, sources, act=ACT_NONE, row_mask=None):
        """sources: 1-3 row-major float32 CUDA tensors [rows, k_i] whose concatenation is the layer's input"""
        self._prepare()
        src = [s.to(torch.float32).contiguous() for s in sources]
        rows = src[0].shape[0]
        if sum(s.shape[1] for s in src) != self.k or any(s.shape[0] != rows for s in src):
            raise _lib.PdpError("TensorLinear: input columns %s do not add up to %d" % ([tuple(s.shape) for s in src], self.k))
        dev = src[0].device
        if dev.type != "cuda":
            raise _lib.PdpError("tensor-core layers expect CUDA tensors (got %s); there is no CPU fallback" % dev)
        while len(src) < 3:
            src.append(None)
        L = _lib.load()
        mk = None if row_mask is None else row_mask.to(torch.float32).reshape(-1).contiguous()
        with torch.cuda.device(dev):
            out = torch.empty(rows, self.n_out, dtype=torch.float32, device=dev)
            k = [0 if s is None else s.shape[1] for s in src]
            _lib.check(L.pdp_edge_mlp_forward(_ptr(src[0]), k[0], _ptr(src[1]), k[1], _ptr(src[2]), k[2], rows, _ptr(self.image),
                                              _ptr(self.bias), self.n_blk, self.n_mma, self.passes, self.n_out, int(act), _ptr(mk),
                                              _ptr(out), ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)),
                       "pdp_edge_mlp_forward", L)
        return out


class TensorGRU(object):
    "h' = GRUCell([x1 | x2], h), optionally blended with h where row_mask == 0 (torch.nn.GRUCell semantics)"
    MAX_COLS = 304    # accumulator columns per pass: four per hidden unit (r, z, W_in x, W_hn h).  Two passes cover the 150 hidden
                      # units of the reference's models (2 x 76): the operand rows are staged once per pass, and wider passes do
                      # not fit twice into the 512 columns of tensor memory (304 + 304 share 96, which the epilogue drains first)

    def __init__(self, cell):
        self._cell = cell
        self._key = None

    def _prepare(self):
        c = self._cell
        key = _versions((c.weight_ih, c.weight_hh, c.bias_ih, c.bias_hh))
        if key == self._key:
            return
        wih, whh = c.weight_ih.detach().float(), c.weight_hh.detach().float()
        H, kx = whh.shape[1], wih.shape[1]
        dev = wih.device
        bih = c.bias_ih.detach().float() if c.bias_ih is not None else torch.zeros(3 * H, device=dev)
        bhh = c.bias_hh.detach().float() if c.bias_hh is not None else torch.zeros(3 * H, device=dev)
        passes = (4 * H + self.MAX_COLS - 1) // self.MAX_COLS
        nh = _round_up((H + passes - 1) // passes, 4)          # hidden units per pass: 16 accumulator columns = four whole units
        n_tot = 4 * nh
        if passes * n_tot > 768:
            raise _lib.PdpError("TensorGRU: %d hidden units do not fit three passes" % H)
        k = kx + H
        # [pass, unit, gate, K]: the four accumulator columns of a unit sit side by side
        wp = torch.zeros(passes, nh, 4, _round_up(k, chunk_k()), dtype=torch.float32, device=dev)
        bp = torch.zeros(passes, nh, 4, dtype=torch.float32, device=dev)
        for p in range(passes):
            u0, u1 = p * nh, min(H, (p + 1) * nh)
            m = u1 - u0
            if m <= 0:
                continue
            # the kernel's rows are [h | x]: the hidden state first, so that both big sources start at even columns (64-bit loads)
            for gate in range(2):        # r, z: input and hidden parts accumulate into the same column
                wp[p, :m, gate, :H] = whh[gate * H + u0: gate * H + u1]
                wp[p, :m, gate, H:k] = wih[gate * H + u0: gate * H + u1]
                bp[p, :m, gate] = bih[gate * H + u0: gate * H + u1] + bhh[gate * H + u0: gate * H + u1]
            wp[p, :m, 2, H:k] = wih[2 * H + u0: 2 * H + u1]      # W_in x   (+ b_in)
            wp[p, :m, 3, :H] = whh[2 * H + u0: 2 * H + u1]       # W_hn h   (+ b_hn), multiplied by r in the epilogue
            bp[p, :m, 2] = bih[2 * H + u0: 2 * H + u1]
            bp[p, :m, 3] = bhh[2 * H + u0: 2 * H + u1]
        self.hidden, self.kx = H, kx
        self.n_blk, self.n_mma, self.passes = n_tot, 1, passes
        self.image = _image(wp.view(passes * n_tot, -1), passes, n_tot)
        self.bias = bp.contiguous()
        self._key = key

    def __call__(self, sources, h, row_mask=None):
        self._prepare()
        src = [s.to(torch.float32).contiguous() for s in sources]
        h = h.to(torch.float32).contiguous()
        rows = h.shape[0]
        if sum(s.shape[1] for s in src) != self.kx or h.shape[1] != self.hidden or any(s.shape[0] != rows for s in src) or len(src) > 2:
            raise _lib.PdpError("TensorGRU: shapes %s / %s do not match the cell (%d -> %d)" %
                                ([tuple(s.shape) for s in src], tuple(h.shape), self.kx, self.hidden))
        dev = h.device
        if dev.type != "cuda":
            raise _lib.PdpError("tensor-core layers expect CUDA tensors (got %s); there is no CPU fallback" % dev)
        while len(src) < 2:
            src.append(None)
        L = _lib.load()
        mk = None if row_mask is None else row_mask.to(torch.float32).reshape(-1).contiguous()
        with torch.cuda.device(dev):
            out = torch.empty_like(h)
            k = [0 if s is None else s.shape[1] for s in src]
            _lib.check(L.pdp_edge_gru_forward(_ptr(src[0]), k[0], _ptr(src[1]), k[1], _ptr(h), self.hidden, rows, _ptr(self.image),
                                              _ptr(self.bias), self.n_blk, self.n_mma, self.passes, _ptr(mk), _ptr(out),
                                              ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)),
                       "pdp_edge_gru_forward", L)
        return out
