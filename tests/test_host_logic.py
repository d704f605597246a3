"""Host-side logic that needs no GPU: launch-chain flags, the bench's ring bookkeeping, kernel hints, build staleness."""
import os

import numpy as np
import pytest


def test_chunk_is_independent_bounds_the_overlap():
    """A flagged launch only waits for its predecessor to have begun; chunk i reuses the buffer of chunk i - R.  Whatever R, the
    launch that last wrote a chunk's buffer must be older than the latest unflagged launch at or before the chunk."""
    from deepmimo_b200.channels import chunk_is_independent
    for r in (1, 2, 3, 4, 7):
        flags = [chunk_is_independent(i, r) for i in range(40)]
        assert flags[0] is False
        for i in range(40):
            last_plain = max(j for j in range(i + 1) if not flags[j])
            # everything before `last_plain` has completed when chunk i starts; chunk i - r must be among it
            assert i - r < last_plain or i - r < 0, (r, i)
        if r > 1:
            assert sum(flags) == 40 - len(range(0, 40, r))          # as many flagged launches as the contract allows
        else:
            assert not any(flags)


def test_ring_segments_cover_what_the_ring_holds():
    import bench
    for n, chunk, r in ((200000, 8192, 3), (4096, 4096, 1), (10, 4, 3), (26076, 8192, 3), (8192 * 6, 8192, 3), (5, 8, 2)):
        segs = bench.ring_segments(n, chunk, r)
        # replay the ring
        held = {}
        for i, a in enumerate(range(0, n, chunk)):
            for k in range(min(chunk, n - a)):
                held[(i % r, k)] = a + k
        got = {}
        for u0, b, r0, rows in segs:
            for k in range(rows):
                got[(b, r0 + k)] = u0 + k
        assert got == held, (n, chunk, r)


def test_kernel_hint_mapping(monkeypatch):
    from deepmimo_b200 import _lib
    monkeypatch.delenv("DMK_FD_KERNEL", raising=False)
    assert _lib.kernel_hint_from_env() == 0
    for name, val in (("tile", 1), ("ffma", 2), ("tc", 3), ("tc1", 4), ("small", 5), ("small1", 6), ("mma", 7), ("rows", 8), ("auto", 0), ("TC", 3)):
        monkeypatch.setenv("DMK_FD_KERNEL", name)
        assert _lib.kernel_hint_from_env() == val
    monkeypatch.setenv("DMK_FD_KERNEL", "fastest")
    with pytest.raises(ValueError):
        _lib.kernel_hint_from_env()
    src = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "dmk.h")).read()
    for name, val in (("DMK_KERNEL_TILE", 1), ("DMK_KERNEL_FFMA", 2), ("DMK_KERNEL_TC ", 3), ("DMK_KERNEL_TC1", 4), ("DMK_KERNEL_SMALL ", 5),
                      ("DMK_KERNEL_SMALL1", 6), ("DMK_KERNEL_MMA", 7), ("DMK_KERNEL_ROWS", 8)):
        assert f"{name.strip()}" in src and f"= {val}" in src.split(name.strip())[1][:12], name


def test_build_staleness_follows_source_content(tmp_path, monkeypatch):
    from deepmimo_b200 import build
    assert os.path.exists(build.LIB), "libdmk.so must have been built (python __graft_entry__.py)"
    assert not build.is_stale()
    h0 = build.source_hash()
    monkeypatch.setenv("DMK_NVCC_EXTRA", "-DDMK_TC_TRACE")          # a different build -> a different hash -> stale
    assert build.source_hash() != h0 and build.is_stale()
    monkeypatch.delenv("DMK_NVCC_EXTRA")
    os.utime(os.path.join(build.CSRC, "dmk_api.cu"))                 # a touched file with the same content is NOT stale
    assert not build.is_stale()


def test_pinned_cap_env(monkeypatch):
    from deepmimo_b200.channels import pinned_cap_bytes
    monkeypatch.delenv("DMK_PINNED_CAP_GIB", raising=False)
    assert pinned_cap_bytes() == 16 << 30
    monkeypatch.setenv("DMK_PINNED_CAP_GIB", "0.5")
    assert pinned_cap_bytes() == 1 << 29


def test_bench_workloads_are_consistent_and_scenarios_build():
    """Every bench workload names a scenario the generator knows, with user counts for the GPU leg and both CPU legs; the oracle
    runs on a handful of its users (the parity block of the bench does exactly that)."""
    import importlib.util
    import numpy as np
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    from deepmimo_b200.synth import scenario, coef_count
    from util import oracle_on_users
    for name, (cfg, var) in bench.WORKLOADS.items():
        assert cfg in bench.DEFAULT_USERS and cfg in bench.CPU_SAMPLE_USERS and cfg in bench.CPU_BASELINE_USERS, name
        s = scenario(cfg, 6, **var)
        assert coef_count(s) > 0
        o = oracle_on_users(s, np.arange(3), procs=1)
        assert o["H"].shape[0] == 3 and not np.isnan(o["H"].view(np.float32)).any(), name
