"""Consumes the dump of oracle/pcl_fixture/dump_pcl_fixture.cpp (real PCL 1.8.1) when somebody has produced it, and pins the
oracle's recalled third-party behaviour to it.  Skipped (with the reason) while tests/golden/pcl_fixture.bin does not exist.
The generator itself is tested unconditionally: it must be pure integer / IEEE arithmetic (bit-reproducible in C++)."""
import os
import struct
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle", "pcl_fixture"))
DUMP = os.path.join(ROOT, "tests", "golden", "pcl_fixture.bin")


def _sections(path):
    raw = open(path, "rb").read()
    assert raw[:8] == b"VGSPCL01"
    pos, out = 8, {}
    while pos < len(raw):
        tag, nbytes = struct.unpack_from("<IQ", raw, pos)
        pos += 12
        out[tag] = raw[pos:pos + nbytes]
        pos += nbytes
    return out


def test_fixture_cloud_is_deterministic():
    import fixture_cloud
    u = fixture_cloud.lcg_stream(4)
    assert [int(v * 4294967296.0) for v in u] == [87628868, 71072467, 2332836374, 2726892157]
    c = fixture_cloud.cloud()
    assert c.shape == (20000, 3) and c.dtype == np.float32
    assert np.isfinite(c).all() and float(c[:, 0].max()) < 6.4 and float(c[:, 2].min()) > 0.09
    # surfaces alternate point by point: insertion order grows PCL's bounding box more than once
    assert abs(float(c[1, 1]) - 3.2) < 0.01 and abs(float(c[0, 2]) - 0.1) < 0.01


@pytest.mark.skipif(not os.path.exists(DUMP), reason="tests/golden/pcl_fixture.bin not produced yet: needs PCL 1.8.1 "
                    "(oracle/pcl_fixture/README.md); until then oracle parity with PCL is unpinned")
def test_oracle_matches_pcl_dump(built_lib):
    import fixture_cloud
    from oracle import oracle
    s = _sections(DUMP)
    xyz = fixture_cloud.cloud()
    np.testing.assert_array_equal(np.frombuffer(s[1], np.float32).reshape(-1, 3).view(np.uint32), xyz.view(np.uint32))
    r = oracle.run(xyz, math=0, voxel_size=0.15, graph_size=0.5)
    np.testing.assert_array_equal(np.frombuffer(s[2], np.float64), r.bbox)                       # PclOctree: dynamic bounding box
    np.testing.assert_array_equal(np.frombuffer(s[3], np.uint32).reshape(-1, 3), r.point_key)    # genOctreeKeyforPoint
    np.testing.assert_array_equal(np.frombuffer(s[4], np.uint32).reshape(-1, 3), r.unit_key)     # leaf-iterator order = voxel ids
    np.testing.assert_array_equal(np.frombuffer(s[5], np.uint32).astype(np.int64), r.unit_offsets)
    np.testing.assert_array_equal(np.frombuffer(s[6], np.int32), r.unit_points)                  # container order inside a leaf
    np.testing.assert_array_equal(np.frombuffer(s[7], np.float32).view(np.uint32), r.unit_center.reshape(-1).view(np.uint32))
    mats = np.frombuffer(s[8], np.float32).reshape(-1, 3, 3)
    vals = np.frombuffer(s[9], np.float32).reshape(-1, 3)
    vecs = np.frombuffer(s[10], np.float32).reshape(-1, 9)
    for m, ev, evec in zip(mats, vals, vecs):                                                    # pcl_eigen33 / pcl_compute_roots
        ev_o, evec_o = oracle.eigen33(m, math=0)
        np.testing.assert_array_equal(ev_o.view(np.uint32), ev.view(np.uint32))
        np.testing.assert_array_equal(np.asarray(evec_o, np.float32).reshape(-1).view(np.uint32), evec.view(np.uint32))
    np.testing.assert_array_equal(np.frombuffer(s[11], np.uint32).astype(np.int64), r.adj_offsets)  # FLANN radius search
    np.testing.assert_array_equal(np.frombuffer(s[12], np.int32), r.adj_idx)                        # order (dist2, index)
    v = oracle.vccs(xyz, voxel_res=0.05, seed_res=0.25, schedule=0)                                 # PCL's sequential schedule
    np.testing.assert_array_equal(np.frombuffer(s[14], np.uint32).astype(np.int32), v.point_label)
    assert int(np.frombuffer(s[15], np.uint32)[0]) == v.max_label
