"""CPU tier, world_size 2 over gloo: the N > 1 path shards the reads in contiguous blocks, every rank aligns
its own block against the replicated contig table, and the ordered gather reproduces the single-process
answer (and the oracle's).  The per-rank aligner here is the CPU emulator of the kernels (no GPU needed)."""
import os
import pickle
import random
import socket
import sys
import tempfile

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    import emul_lib
    import gen
    from stitch_b200 import sharding
    from stitch_b200._abi import make_opts
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = random.Random(77)
    contigs = [gen.rand_seq(rng, rng.randint(200, 500)) for _ in range(3)]
    reads = [gen.chimeric_read(rng, contigs, 150, 2, strands=True, wrap=True) for _ in range(7)]
    named = [(f"c{k}", s) for k, s in enumerate(contigs)]
    kw = dict(double_strand=True, circular=True)
    lo, hi = sharding.block_range(len(reads), rank, world)
    local = emul_lib.EmulAligners(make_opts(**kw), named, strip=8).batch(reads[lo:hi], raw=False)
    keys = [[a.key() for a in chains] for chains in local]
    allkeys = sharding.gather_in_order(keys)
    if rank == 0:
        with open(out_path, "wb") as f:
            pickle.dump((allkeys, [r for r in reads], named, kw), f)
    dist.barrier()
    dist.destroy_process_group()


def test_block_range_partitions():
    from stitch_b200.sharding import block_range
    for n in (0, 1, 7, 8, 1000):
        for world in (1, 2, 3, 8):
            spans = [block_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[k][1] == spans[k + 1][0] for k in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.timeout(300)
def test_two_ranks_match_single_process(oracle):
    import torch.multiprocessing as mp
    from stitch_b200._abi import make_opts
    with tempfile.TemporaryDirectory() as d:
        out = os.path.join(d, "res.pkl")
        mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
        allkeys, reads, named, kw = pickle.load(open(out, "rb"))
    exp, _ = oracle.OracleAligners(make_opts(**kw), named).batch(reads, raw=False)
    assert len(allkeys) == len(reads)
    for got, chains in zip(allkeys, exp):
        assert got == [a.key() for a in chains]
