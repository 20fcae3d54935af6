"""The N > 1 path on CPU (gloo, world_size 2): the contiguous-slice partition, the max-over-ranks
timing reduction bench.py uses, and the reference arm under torchrun (rank 0 alone prints one JSON
line).  The GPU side of the same path: tests/test_gpu_parity.py::test_multi_device_context_matches_single_device
and ::test_config5_sharded_2p24_with_product_gather (skipped on single-GPU boxes; `gpurun --gpus 2`), and
bench.py's `product` / `config5` / `strong` records under torchrun (NCCL all_gather of the Fp12 partials)."""
import json
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, q):
    import torch.distributed as dist
    from zkvm_pairings_b200 import sharding
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = sharding.slice_bounds(n, world, rank)
    slowest = sharding.max_over_ranks(100.0 + 7.0 * rank, world)          # rank 1 is slower
    rate = sharding.whole_job_rate(hi - lo, 3, world, slowest)
    dist.barrier()
    q.put((rank, lo, hi, slowest, rate, sharding.synthetic_first_index(rank, 1 << 20)))
    dist.destroy_process_group()


def test_slices_and_max_reduction_world_size_2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q, port, n, world = ctx.Queue(), _free_port(), 1001, 2
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, lo0, hi0, m0, rate0, f0), (r1, lo1, hi1, m1, rate1, f1) = res
    assert (lo0, hi0, lo1, hi1) == (0, 500, 500, 1001)                    # contiguous, disjoint, covering
    assert m0 == m1 == 107.0                                              # both ranks see the slowest time
    assert abs(rate0 - 2 * 500 * 3 / 0.107) < 1e-6 and f1 == 1 << 20 and f0 == 0


def _product_worker(rank, world, port, n, q):
    """One rank of a sharded product with the ORACLE standing in for the GPU: Miller product of its slice, gather of
    the 576-byte partials (gloo), product of the partials, one final exponentiation."""
    import numpy as np
    import torch
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import coracle
    import util
    from zkvm_pairings_b200 import sharding
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g1, i1, g2, i2 = util.oracle_points(coracle, 0xA11, 0, n)
    i1[1] = 1
    lo, hi = sharding.slice_bounds(n, world, rank)
    ml, _ = coracle.miller_product(g1[lo:hi], i1[lo:hi], g2[lo:hi], i2[lo:hi])
    parts = sharding.gather_partials(torch.from_numpy(ml.view(np.int64)), world).numpy().view(np.uint64)
    acc = parts[0]
    for r in range(1, world):
        acc = coracle.tower_op("fp12_mul", acc[None], parts[r][None])[0]
    gt = coracle.final_exp_batch(acc[None])[0]
    eml, egt = coracle.miller_product(g1, i1, g2, i2)
    q.put((rank, bool(np.array_equal(acc, eml)), bool(np.array_equal(gt, egt)), bool(np.array_equal(parts[rank], ml))))
    dist.barrier()
    dist.destroy_process_group()


def test_product_partial_gather_world_size_2():
    """SURVEY 8e's one exchange step on CPU: contiguous slices, all_gather of the Fp12 partials, product, one final
    exponentiation -- bit-identical to the unsharded product on every rank (uneven split: 7 pairs over 2 ranks)."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q, port, world = ctx.Queue(), _free_port(), 2
    procs = [ctx.Process(target=_product_worker, args=(r, world, port, 7, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [(0, True, True, True), (1, True, True, True)]


def test_slice_bounds_cover_every_partition():
    from zkvm_pairings_b200.sharding import slice_bounds
    for n in (0, 1, 7, 8, 1 << 20, (1 << 24) + 3):
        for parts in (1, 2, 3, 4, 8):
            b = [slice_bounds(n, parts, i) for i in range(parts)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(parts - 1))
            assert max(hi - lo for lo, hi in b) - min(hi - lo for lo, hi in b) <= 1
    with pytest.raises(ValueError):
        slice_bounds(10, 2, 2)


def test_reference_arm_under_torchrun_prints_one_json_line():
    """bench.py --impl reference with 2 ranks: rank 0 runs the CPU oracle and prints ONE JSON line, rank 1 exits 0."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
           "--warmup", "0", "--ref-seconds", "1"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "pairings_per_sec" and d["n_gpus"] == 2
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "port" and d["e2e"]["h2d_bytes_per_step"] == 0
