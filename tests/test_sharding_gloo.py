"""The N>1 path on CPU: world_size-2 gloo processes shard the pair list and gather the match tables."""
import os
import subprocess
import sys

import numpy as np

from conftest import ROOT

WORKER = r"""
import os, sys
sys.path.insert(0, os.environ["DIMB_ROOT"])
import numpy as np, torch, torch.distributed as dist
from dim_b200.sharded import shard_pairs, gather_match_tables, shard_images, store_slot, images_per_rank, all_gather_blocks
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
# phase-1 -> phase-2 exchange: 7 images dealt i % world, one "feature block" of 48 bytes per image in rank-major slots
n_img, slot_bytes = 7, 48
ipr = images_per_rank(n_img, world)
store = torch.zeros(world * ipr, slot_bytes, dtype=torch.uint8)
block = lambda i: torch.full((slot_bytes,), 10 + i, dtype=torch.uint8)
for i in shard_images(n_img, world, rank):
    store[store_slot(i, n_img, world)] = block(i)
got = all_gather_blocks(store, n_img, dist)
assert got == (world - 1) * ipr * slot_bytes, got
for i in range(n_img):
    assert torch.equal(store[store_slot(i, n_img, world)], block(i)), i
assert sorted(store_slot(i, n_img, world) for i in range(n_img)) == sorted(set(store_slot(i, n_img, world) for i in range(n_img)))
print("EXCHANGE_OK", rank)
n = 11
costs = [(i * 7919) % 13 + 1 for i in range(n)]
mine = shard_pairs(n, world, rank, costs)
rng = lambda i: np.random.default_rng(i)
tables = [rng(i).integers(0, 2048, (int(rng(i).integers(0, 40)), 2)).astype(np.int64) for i in mine]
full = gather_match_tables(mine, tables, n, dist)
if rank == 0:
    for i in range(n):
        exp = rng(i).integers(0, 2048, (int(rng(i).integers(0, 40)), 2)).astype(np.int64)
        assert np.array_equal(full[i], exp), i
    print("GATHER_OK", sum(len(t) for t in full))
else:
    assert full is None
dist.destroy_process_group()
"""


def test_shard_pairs_partitions():
    from dim_b200.sharded import shard_pairs
    for world in (1, 2, 3, 8):
        for costs in (None, [(i * 31) % 17 + 1 for i in range(50)]):
            parts = [shard_pairs(50, world, r, costs) for r in range(world)]
            assert sorted(sum(parts, [])) == list(range(50))
            if costs is not None and world > 1:
                loads = [sum(costs[i] for i in p) for p in parts]
                assert max(loads) - min(loads) <= max(costs)  # LPT balance bound
    assert shard_pairs(10, 4, 1) == [1, 5, 9]


def test_image_sharding_and_store_slots():
    from dim_b200.sharded import images_per_rank, shard_images, store_slot
    for n, world in ((100, 8), (7, 2), (5, 1), (3, 4)):
        parts = [shard_images(n, world, r) for r in range(world)]
        assert sorted(sum(parts, [])) == list(range(n))
        ipr = images_per_rank(n, world)
        slots = [store_slot(i, n, world) for i in range(n)]
        assert len(set(slots)) == n and max(slots) < world * ipr
        for r in range(world):  # a rank's images occupy its own contiguous region of the store
            assert all(r * ipr <= store_slot(i, n, world) < (r + 1) * ipr for i in parts[r])


def test_gather_single_process():
    from dim_b200.sharded import gather_match_tables
    out = gather_match_tables([2, 0], [np.zeros((0, 2)), np.array([[1, 2], [3, 4]])], 3)
    assert out[1] is None and out[2].shape == (0, 2) and np.array_equal(out[0], [[1, 2], [3, 4]])


def test_world_size_2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = {**os.environ, "DIMB_ROOT": ROOT}
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29591", str(script)], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "GATHER_OK" in r.stdout
    assert r.stdout.count("EXCHANGE_OK") == 2


def test_shard_pairs_properties():
    """Property test (hypothesis): every pair goes to exactly one rank, the deal is the same on every rank, and the cost-aware
    deal is within the classical LPT bound (max load <= 4/3 OPT + ... <= mean + max cost)."""
    from hypothesis import given, settings
    from hypothesis import strategies as st

    from dim_b200.sharded import shard_pairs

    @settings(max_examples=60, deadline=None)
    @given(st.integers(0, 200), st.integers(1, 8), st.booleans(), st.integers(0, 2 ** 31 - 1))
    def check(n, world, with_costs, seed):
        import numpy as np
        costs = np.random.default_rng(seed).integers(1, 4096 * 4096, n).astype(float) if with_costs else None
        parts = [shard_pairs(n, world, r, costs) for r in range(world)]
        flat = sorted(i for p in parts for i in p)
        assert flat == list(range(n))
        assert all(p == sorted(p) for p in parts)
        if with_costs and n:
            loads = [sum(costs[i] for i in p) for p in parts]
            assert max(loads) <= sum(costs) / world + max(costs) + 1e-6
    check()
