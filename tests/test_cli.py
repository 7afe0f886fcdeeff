"""CLI shims of the voxel path: argument surface on CPU, end-to-end runs on the GPU."""
import gzip
import io
import json
import os

import numpy as np
import pytest


def test_single_parser_surface():
    from pdb_eda_b200 import singleStructure
    p = singleStructure.buildParser()
    a = p.parse_args(["1cbs", "-", "density", "--residue", "--radius", "2.5", "--atom-mask", "m.json", "--optimized-radii", "--out-format", "csv"])
    assert a.submode == "density" and a.residue and a.radius == 2.5 and a.atom_mask == "m.json" and a.optimized_radii and a.out_format == "csv"
    a = p.parse_args(["1cbs", "out.json", "blob", "--green", "--red", "--num-sd", "3.5", "--include-pdbid"])
    assert a.green and a.red and a.num_sd == 3.5 and a.include_pdbid
    with pytest.raises(SystemExit):
        p.parse_args(["1cbs", "-", "nosuchmode"])
    assert singleStructure.numpyConverter(np.float64(1.5)) == 1.5 and singleStructure.numpyConverter(np.arange(3)) == [0, 1, 2]


def _write_entry(folder, pdbid, seed):
    from pdb_eda_b200 import structure, synthetic
    cell, n = (24.0, 24.0, 24.0, 90, 90, 90), (48, 48, 48)
    st = synthetic.polyAlaStructure(45, (0, 0, 0), cell[:3], seed=seed)
    a, b = synthetic.mapPair(st, n, cell, seed=seed + 1)
    open(os.path.join(folder, pdbid + ".ccp4"), "wb").write(synthetic.ccp4Bytes(a, cell, n))
    open(os.path.join(folder, pdbid + "_diff.ccp4"), "wb").write(synthetic.ccp4Bytes(b, cell, n))
    text = structure.formatPDB(st, remark290=synthetic.cartesianOperators("P 21 21 21", cell), cell=cell, spaceGroup="P 21 21 21")
    with gzip.open(os.path.join(folder, "pdb" + pdbid + ".ent.gz"), "wt") as fh:
        fh.write(text)
    return text


@pytest.mark.gpu
def test_single_and_multiple_end_to_end(tmp_path):
    from pdb_eda_b200 import densityAnalysis, multipleStructures, singleStructure, synthetic
    densityAnalysis.setGlobals(synthetic.defaultParams())
    folder = str(tmp_path)
    ids = ["1aaa", "2bbb", "3ccc"]
    for k, pdbid in enumerate(ids):
        _write_entry(folder, pdbid, 10 + 3 * k)
    files = ["--pdb-file", os.path.join(folder, "pdb1aaa.ent.gz"), "--density-file", os.path.join(folder, "1aaa.ccp4"), "--diff-file",
             os.path.join(folder, "1aaa_diff.ccp4")]
    out = os.path.join(folder, "blob.json")
    singleStructure.main(["1AAA", out, "blob", "--green", "--red", "--include-pdbid"] + files)
    rows = json.load(open(out))
    an = densityAnalysis.fromFile(files[1], files[3], files[5])
    want = an.calculateAtomSpecificBlobStatistics(an.greenBlobList + an.redBlobList)
    assert len(rows) == len(want) > 0 and rows[0]["pdbid"] == "1aaa"
    assert [r["num_voxels"] for r in rows] == [w[3] for w in want]
    np.testing.assert_allclose([r["distance_to_atom"] for r in rows], [w[0] for w in want], rtol=1e-12)
    out = os.path.join(folder, "res.csv")
    singleStructure.main(["1aaa", out, "difference", "--residue", "--type", "ALA", "--out-format", "csv"] + files)
    lines = open(out).read().strip().splitlines()
    assert lines[0].split(",")[:5] == ["model", "chain", "residue_number", "residue_name", "mean_occupancy"] and len(lines) == 46
    out = os.path.join(folder, "cloud.json")
    singleStructure.main(["1aaa", out, "cloud", "--atom"] + files)
    assert len(json.load(open(out))) == len(an.atomCloudDescriptions)
    # multiple mode over a directory of cached entries, one of which is missing
    idfile = os.path.join(folder, "ids.txt")
    open(idfile, "w").write(" ".join(ids + ["9zzz"]))
    out = os.path.join(folder, "multi.json")
    multipleStructures.main([idfile, out, "--data-dir", folder])
    res = json.load(open(out))
    assert sorted(res["entries"]) == ids and res["cumulative"]["structures"] == 3
    out2 = os.path.join(folder, "multi_workers.json")
    multipleStructures.main([idfile, out2, "--data-dir", folder, "--workers", "2"])        # host worker pool on the same GPU
    res2 = json.load(open(out2))
    assert sorted(res2["entries"]) == ids and res2["cumulative"]["num_voxels_aggregated"] == res["cumulative"]["num_voxels_aggregated"]
    assert res2["entries"]["2bbb"]["stats"]["density_electron_ratio"] == res["entries"]["2bbb"]["stats"]["density_electron_ratio"]
    np.testing.assert_allclose(res["entries"]["1aaa"]["stats"]["density_electron_ratio"], an.densityElectronRatio, rtol=1e-12)


@pytest.mark.gpu
def test_optimize_service_and_pinned_loader(tmp_path):
    """Persistent optimiser inner loop and the pinned load path (8f-1)."""
    import copy
    from pdb_eda_b200 import ccp4, densityAnalysis, multi, multipleStructures, synthetic
    folder = str(tmp_path)
    ids = ["1aaa", "2bbb"]
    for k, pdbid in enumerate(ids):
        _write_entry(folder, pdbid, 20 + 3 * k)
    threshold = ccp4.PINNED_MIN_BYTES
    ccp4.PINNED_MIN_BYTES = 1024
    try:
        dm = ccp4.read(os.path.join(folder, "1aaa.ccp4"))
        assert dm._host32 is None and ccp4._ring[0][0].is_pinned()           # file -> page-locked ring -> HBM, no host copy
        with pytest.raises(AssertionError):
            ccp4.parse(io.BytesIO(open(os.path.join(folder, "1aaa.ccp4"), "rb").read()[:-8]), "x")
    finally:
        ccp4.PINNED_MIN_BYTES = threshold
    ref = ccp4.parse(io.BytesIO(open(os.path.join(folder, "1aaa.ccp4"), "rb").read()), "x")
    assert ref._host32 is not None                                           # small maps: host array, uploaded on first use
    assert np.array_equal(dm.densityArray, ref.densityArray) and dm.meanDensity == ref.meanDensity
    with pytest.raises(AssertionError):
        ccp4.parse(io.BytesIO(open(os.path.join(folder, "1aaa.ccp4"), "rb").read()[:-8]), "x")
    params = synthetic.defaultParams()
    service = multi.OptimizeService(ids, multipleStructures.makeLoader(folder))
    first = service.evaluate(params)
    bigger = copy.deepcopy(params)
    bigger["radii"] = {t: r * 1.25 for t, r in params["radii"].items()}
    second = service.evaluate(bigger)
    assert first[0] != second[0]                                             # the radii matter
    again = service.evaluate(params)
    assert again[0] == first[0] and again[5] == first[5]                     # and the service is stateless across iterations
    densityAnalysis.setGlobals(bigger)
    fresh = multi.gatherResults({i: multi.analyzeStructure(multipleStructures.makeLoader(folder)(p), list(bigger["radii"]), optimizer=True)
                                 for i, p in enumerate(ids)},
                                [0, 1], 2, list(bigger["radii"]), "cpu")
    assert fresh["medianDiffs"].keys() == second[0].keys()
    np.testing.assert_allclose([fresh["medianDiffs"][t] for t in second[0]], [second[0][t] for t in second[0]], rtol=1e-9)   # batched vs per-structure path
    densityAnalysis.setGlobals(params)
