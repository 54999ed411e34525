"""CPU tests of the host side: the C-ABI library exports what include/pdp_b200.h declares (no compute calls
without a GPU), the product path refuses to run without CUDA, and the multi-GPU sharding logic
(world_size 2 over gloo) reproduces the unsharded result with the C oracle standing in for the kernels."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    src = open(os.path.join(ROOT, "include", "pdp_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pdp_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_bound_signatures():
    from pdp_solver_b200 import _lib
    assert set(_declared_functions()) == set(_lib.SIGNATURES), \
        set(_declared_functions()) ^ set(_lib.SIGNATURES)


@pytest.mark.parametrize("libname", ["libpdp_b200.so", "libpdp_b200_strict.so"])
def test_library_loads_and_exports_every_symbol(libname):
    path = os.path.join(ROOT, "pdp_solver_b200", "csrc", libname)
    if not os.path.exists(path):
        subprocess.check_call(["bash", os.path.join(ROOT, "pdp_solver_b200", "csrc", "build.sh")])
    lib = ctypes.CDLL(path)
    for fn in _declared_functions():
        assert hasattr(lib, fn), fn
    lib.pdp_version.restype = ctypes.c_char_p
    assert b"sm_100a" in lib.pdp_version()
    lib.pdp_workspace_bytes.restype = ctypes.c_size_t
    lib.pdp_workspace_bytes.argtypes = [ctypes.c_int64] * 4
    assert lib.pdp_workspace_bytes(1260, 100, 420, 1) > 0
    assert lib.pdp_workspace_bytes(-1, 0, 0, 0) == 0


def test_no_cpu_fallback():
    import torch
    from pdp_solver_b200 import _lib, cnfgen
    from pdp_solver_b200.engine import Context
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    gm, bvm, bfm, ef = [torch.from_numpy(x) for x in cnfgen.random_batch(2, 20, 3, 4.0, 0)]
    with pytest.raises(_lib.PdpError):
        Context(gm, bvm, bfm, ef)


def test_lpt_and_extract_roundtrip():
    from pdp_solver_b200 import cnfgen, sharding
    batch = cnfgen.mixed_batch([(30, 3, 4.0), (80, 3, 4.2), (10, 3, 3.0), (50, 5, 16.0), (50, 3, 4.1)], 3)
    nv, nf, ne = sharding.problem_sizes(*batch[:3])
    assert nv.tolist() == [30, 80, 10, 50, 50] and ne.sum() == batch[0].shape[1]
    parts = sharding.lpt_assign(ne, 2)
    assert sorted(parts[0] + parts[1]) == [0, 1, 2, 3, 4]
    loads = [int(ne[p].sum()) for p in parts]
    assert max(loads) - min(loads) <= int(ne.max())
    seen = np.zeros(nv.sum(), dtype=int)
    for p in parts:
        sub, vsel = sharding.extract_problems(batch, p)
        seen[vsel] += 1
        assert sub[1].max() + 1 == len(p) and sub[0][0].max() < sub[1].size and sub[0][1].max() < sub[2].size
        # the sub-batch holds exactly the clauses of its problems, literal for literal
        gm, bvm, bfm, ef = batch
        esel = np.isin(bvm[gm[0]], p)
        assert np.array_equal(vsel[sub[0][0]], gm[0][esel]) and np.array_equal(sub[3].reshape(-1), ef.reshape(-1)[esel])
    assert (seen == 1).all()


_WORKER = r'''
import os, sys
import numpy as np
import torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from oracle import pdp_oracle as po
from pdp_solver_b200 import cnfgen, sharding

def solve(sub):
    o = po.Oracle(*sub, strict=False)
    o.simplify()
    o.set_state(*po.init_state(sub[0].shape[1], False))
    o.run(60, 0.02, 20, True)
    pred = o.masks()["sol"]
    solved, _ = o.cnf_eval(pred)
    return pred, solved

dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
batch = cnfgen.mixed_batch([(40, 3, 3.6), (60, 3, 3.8), (25, 3, 3.0), (70, 3, 3.9), (30, 3, 3.5), (55, 3, 3.7)], 11)
pred, solved = sharding.solve_sharded(batch, solve, rank, world, dist)
if rank == 0:
    ref_pred, ref_solved = solve(batch)
    assert np.array_equal(pred, ref_pred), "sharded prediction differs from the unsharded one"
    assert np.array_equal(solved, ref_solved)
    print("SHARDED_OK", int(solved.sum()))
dist.destroy_process_group()
'''


def test_sharded_equals_unsharded_gloo_world2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29617", str(script), ROOT],
                         capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "SHARDED_OK" in out.stdout


def test_per_problem_max_and_argmax_follow_the_reference_rounding():
    """problem_max / problem_argmax (np-d-np, reinforce) = util.sparse_max / sparse_argmax (reference util.py:257-275) on a
    single-problem batch: fl(fl(max fl(fl(x - min) + 1) + min) - 1), zero floor, first index on ties"""
    import torch
    from pdp_solver_b200.nn.pdp_decimate import problem_argmax, problem_max
    rng = np.random.default_rng(5)
    for _ in range(20):
        n = int(rng.integers(1, 40))
        x = rng.random(n).astype(np.float32) * np.float32(rng.choice([1e-3, 1.0, 50.0]))
        if n > 3:
            x[int(rng.integers(0, n))] = x.max()            # a tie: the first index must win
        mn = x.min()
        y = ((x - mn).astype(np.float32) + np.float32(1)).astype(np.float32)
        want_max = ((max(y.max(), np.float32(0)) + mn).astype(np.float32) - np.float32(1)).astype(np.float32)
        bvm = torch.zeros(n, dtype=torch.int32)
        assert problem_max(torch.from_numpy(x), bvm, 1).numpy()[0] == want_max
        assert int(problem_argmax(torch.from_numpy(x), bvm, 1)[0]) == int(np.argmax(y))
    # three problems of different sizes in one batch: each behaves as if alone
    x = torch.tensor([0.2, 0.9, 0.9, 0.0, 0.0, 0.5, 0.1, 0.7], dtype=torch.float32)
    bvm = torch.tensor([0, 0, 0, 1, 1, 2, 2, 2], dtype=torch.int32)
    assert problem_argmax(x, bvm, 3).tolist() == [1, 3, 7]
    assert torch.allclose(problem_max(x, bvm, 3), torch.tensor([0.9, 0.0, 0.7]), atol=1e-6)


def _decode_image(image, n_tot):
    """inverse of nn/tensor_ops._image: [pass, chunk, term, row, unit, j] (64-byte swizzled rows) -> hi + lo as [pass, row, K]"""
    import torch
    passes, chunks, terms, rows, units, four = image.shape
    assert (terms, rows, units, four) == (2, n_tot, 4, 4)
    n = torch.arange(rows)
    slot_of_group = torch.arange(4).view(1, 4) ^ ((n >> 1) & 3).view(-1, 1)              # [row, group] -> unit it is stored at
    idx = slot_of_group.view(1, 1, 1, rows, 4, 1).expand(passes, chunks, 2, rows, 4, 4)
    groups = torch.gather(image, 4, idx)                                                  # [pass, chunk, term, row, group, j]
    w = groups.sum(2)                                                                     # hi + lo
    return w.permute(0, 2, 1, 3, 4).reshape(passes, rows, chunks * 16)


def test_tensor_core_weight_images_hold_the_layers_weights():
    """Host side of the tcgen05 layers (nn/tensor_ops.py), no GPU: the GRU cell's image has four rows per hidden unit
    (r, z, W_in x, W_hn h) with K ordered [h | x], the dense layer's image is W itself; hi + lo reproduces the fp32 weights
    to 2^-21, both parts are tf32 numbers, padding is zero."""
    import torch
    from pdp_solver_b200.nn import tensor_ops
    torch.manual_seed(3)
    cell = torch.nn.GRUCell(151, 150)
    tg = tensor_ops.TensorGRU(cell)
    tg._prepare()
    H, kx = 150, 151
    assert (tg.passes, tg.n_blk * tg.n_mma) == (2, 304)
    nh = tg.n_blk * tg.n_mma // 4
    w = _decode_image(tg.image, 4 * nh)                     # [pass, 4 nh, K padded]
    wih, whh = cell.weight_ih.detach(), cell.weight_hh.detach()
    assert ((tg.image.view(torch.int32) & 0x1fff) == 0).all()           # tf32: the low 13 mantissa bits are clear
    for p in range(tg.passes):
        for u in (0, 1, 37, nh - 1):
            unit = p * nh + u
            rows = w[p, 4 * u: 4 * u + 4]
            if unit >= H:
                assert (rows == 0).all()
                continue
            want = torch.zeros(4, w.shape[2])
            for gate in range(2):
                want[gate, :H] = whh[gate * H + unit]
                want[gate, H:H + kx] = wih[gate * H + unit]
            want[2, H:H + kx] = wih[2 * H + unit]
            want[3, :H] = whh[2 * H + unit]
            assert torch.allclose(rows, want, rtol=2.0 ** -21, atol=0)
            assert (rows[:, H + kx:] == 0).all()
        bias = tg.bias[p]
        unit = p * nh + 5
        assert torch.allclose(bias[5], torch.stack((cell.bias_ih[unit] + cell.bias_hh[unit], cell.bias_ih[H + unit] + cell.bias_hh[H + unit],
                                                    cell.bias_ih[2 * H + unit], cell.bias_hh[2 * H + unit])).detach())
    lin = torch.nn.Linear(151, 100)
    tl = tensor_ops.TensorLinear(lin)
    tl._prepare()
    wl = _decode_image(tl.image, tl.n_blk)[0]
    assert tl.n_blk == 112 and torch.allclose(wl[:100, :151], lin.weight.detach(), rtol=2.0 ** -21, atol=0)
    assert (wl[100:] == 0).all() and (wl[:, 151:] == 0).all() and torch.equal(tl.bias[:100], lin.bias.detach())


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (CPU only: the unmodified reference when it can be loaded, else the C port) prints one
    JSON line with the bench contract's keys, `impl: reference`, a cpu_baseline describing the run and an e2e object
    without copies."""
    import json
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--workload", "config4",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference"
    if "unavailable" in line:
        pytest.skip("reference arm unavailable: %s" % line["unavailable"])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "dtype", "data",
                "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["metric"] == "edge_updates_per_s" and line["value"] > 0 and "workload" in line["config"]
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0 and line["e2e"]["value"] == line["value"]
