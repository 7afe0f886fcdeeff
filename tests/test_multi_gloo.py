"""CPU tier: the N > 1 host logic of multiple-structures mode over torch.distributed (gloo, world_size 2)."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

TYPES = ["C.a", "N.b", "O.c"]


def _fake_result(i):
    """Deterministic stand-in for analyzeStructure(); structure 5 fails (returns 0)."""
    if i == 5:
        return 0
    rng = np.random.default_rng(100 + i)
    return {"pdbid": "s%03d" % i,
            "diffs": {t: float(rng.normal()) if (i + k) % 4 else float("nan") for k, t in enumerate(TYPES)},
            "slopes": {t: float(rng.normal()) for t in TYPES},
            "stats": {"density_electron_ratio": 1.0 + 0.01 * i, "voxel_volume": 0.125, "num_voxels_aggregated": 1000 + i,
                      "total_aggregated_electrons": 500.0 + i, "total_aggregated_density": 600.0 + 2 * i, "num_atoms_analyzed": 10 * i,
                      "num_residue_clouds_analyzed": 2 * i, "num_domain_clouds_analyzed": i, "atom_overlap_completeness": 0.5,
                      "execution_time": 0.1},
            "atomtype_overlap_completeness": {t: i + k for k, t in enumerate(TYPES)},
            "atomtype_overlap_incompleteness": {t: 1 for t in TYPES}}


def _worker(rank, world, port, n, out):
    import torch.distributed as dist
    from pdb_eda_b200 import multi
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    costs = [(i * 7) % 11 + 1 for i in range(n)]
    mine = multi.shardStructures(costs, world)[rank]
    results = {i: _fake_result(i) for i in mine}
    summary = multi.gatherResults(results, mine, n, TYPES, "cpu")
    out[rank] = (summary["cumulative"], summary["rows"], summary["medianDiffs"], summary["overallStdDevDiffs"],
                 summary["atomTypeOverlapCompleteness"], mine)
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def test_shard_is_a_balanced_partition():
    from pdb_eda_b200 import multi
    costs = [5, 1, 9, 3, 3, 8, 2, 7]
    shards = multi.shardStructures(costs, 3)
    assert sorted(i for s in shards for i in s) == list(range(8))
    loads = [sum(costs[i] for i in s) for s in shards]
    assert max(loads) - min(loads) <= max(costs)
    assert multi.shardStructures(costs, 3) == shards
    assert multi.shardStructures([], 2) == [[], []]


@pytest.mark.timeout(120)
def test_gather_matches_single_process():
    from pdb_eda_b200 import multi
    n, world = 13, 2
    manager = mp.Manager()
    out = manager.dict()
    mp.spawn(_worker, args=(world, _free_port(), n, out), nprocs=world, join=True)
    single = multi.gatherResults({i: _fake_result(i) for i in range(n)}, list(range(n)), n, TYPES, "cpu")
    assert sorted(out[0][5] + out[1][5]) == list(range(n)) and not set(out[0][5]) & set(out[1][5])
    for rank in range(world):
        cumulative, rows, medianDiffs, overallStd, completeness, _ = out[rank]
        assert cumulative["structures"] == n - 1 == single["cumulative"]["structures"]      # the failed structure is skipped
        assert cumulative["num_voxels_aggregated"] == single["cumulative"]["num_voxels_aggregated"]
        assert cumulative["atomtype_overlap_completeness"] == single["cumulative"]["atomtype_overlap_completeness"]
        np.testing.assert_array_equal(rows, single["rows"])
        assert 5 not in rows[:, 0]
        for t in TYPES:
            assert medianDiffs[t] == single["medianDiffs"][t] and completeness[t] == single["atomTypeOverlapCompleteness"][t]
        assert overallStd == single["overallStdDevDiffs"]
