"""CPU-side checks of the boundary: the C-ABI library loads and exports every symbol the header
declares, fails loudly without a CUDA device (no CPU fallback), the forest reader mirrors the
reference's, and the multi-GPU sharding logic works under gloo with world_size 2."""
import ctypes as C
import os
import re
import socket

import numpy as np
import pytest

from helpers import FORESTS
from oraclelib import ROOT


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "gpc_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(gpc_[a-z_0-9]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    import opengpc_b200 as g
    from opengpc_b200 import capi
    lib = g.load_library()
    names = _declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/gpc_b200.h but not exported"
    assert set(capi.SYMBOLS) <= set(names)


def test_header_is_plain_c():
    """The boundary is a C ABI: include/gpc_b200.h compiles as C99 (plain pointers and sizes, no C++ or torch
    types), and the ctypes binding knows every symbol it declares."""
    import subprocess
    from opengpc_b200 import capi
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-fsyntax-only", "-x", "c",
                        os.path.join(ROOT, "include", "gpc_b200.h")], capture_output=True, text=True)
    assert r.returncode == 0 and not r.stderr.strip(), r.stderr
    assert sorted(capi.SYMBOLS) == _declared_symbols()


def test_no_cpu_fallback():
    """Without a device gpc_create must fail with GPC_E_CUDA; nothing computes on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    import opengpc_b200 as g
    from opengpc_b200 import capi
    with pytest.raises(g.GpcError) as e:
        g.Context(device=0, max_w=64, max_h=64, max_batch=1)
    assert e.value.status == capi.GPC_E_CUDA
    assert "no CPU fallback" in str(e.value)


def test_forest_reader_matches_oracle(oracle):
    import opengpc_b200 as g
    for name, path in FORESTS.items():
        f, o = g.read_forest(path), oracle.read_forest(path)
        assert (f.n_tests, f.type, f.n_discarded) == (o.n_tests, o.type, o.n_discarded)
        for k in ("ix", "iy", "jx", "jy", "tau"):
            assert list(getattr(f, k))[:f.n_tests] == list(getattr(o, k))[:o.n_tests], (name, k)
    assert g.read_forest(FORESTS["deep"]).n_ferns == 16 and g.read_forest(FORESTS["deep"]).n_discarded == 160
    with pytest.raises(g.GpcError) as e:
        g.read_forest("/nonexistent/forest.txt")
    from opengpc_b200 import capi
    assert e.value.status == capi.GPC_E_IO


def test_struct_layouts():
    import opengpc_b200 as g
    from opengpc_b200 import capi
    assert g.SUPPORT_DTYPE.itemsize == 12 and g.CORR_DTYPE.itemsize == 16
    assert C.sizeof(capi.GpcSettings) == 24
    assert C.sizeof(capi.GpcForest) == 4 * (3 + 5 * 32 + 1)


def test_shard_pairs_partition():
    from opengpc_b200.shard import owner_of, shard_pairs
    for n, world in ((256, 8), (7, 2), (3, 4), (0, 2)):
        parts = [shard_pairs(n, r, world) for r in range(world)]
        allp = np.sort(np.concatenate(parts))
        assert np.array_equal(allp, np.arange(n))
        for r, p in enumerate(parts):
            assert all(owner_of(i, world) == r for i in p)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from opengpc_b200.shard import gather_support_counts, reduce_timing, shard_pairs
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        n_pairs = 11
        mine = shard_pairs(n_pairs, rank, world)
        counts = 100 + 3 * mine                         # stand-in for per-pair support counts
        dist.barrier()
        ms, units = reduce_timing(10.0 + rank, len(mine), dist)
        allc = gather_support_counts(counts, mine, n_pairs, dist)
        q.put((rank, ms, units, allc.tolist()))
    finally:
        dist.destroy_process_group()


def test_sharding_under_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ms, units, allc in res:
        assert ms == 11.0                               # max over ranks
        assert units == 11                              # every pair processed exactly once
        assert allc == [100 + 3 * i for i in range(11)]
