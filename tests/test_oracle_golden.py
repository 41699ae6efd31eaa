"""CPU: the oracle (oracle/pcst_oracle.c + oracle/ref_oracle.py) against golden vectors produced
by executing the reference (oracle/gen_golden.py).  This is the pin that makes the oracle a valid
checker for the CUDA path: integer/index outputs and loss-form minima bit-exact, floats within the
stated tolerance."""
import hashlib

import numpy as np
import pytest

from pointcloud_style_transfer_b200 import synthetic as S


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def test_square_distance_bit_exact(golden, oracle):
    g = golden("square_distance")
    assert np.array_equal(bits(oracle.square_distance(g["src"], g["dst"])), bits(g["out"]))


def test_c1_encoder_indices_bit_exact_and_features(golden, oracle):
    g = golden("c1_encoder")
    sd = {k[3:]: v for k, v in g.items() if k.startswith("sd.")}
    out = oracle.encoder_forward(g["x"], sd, g["start1"], g["start2"])
    assert np.array_equal(out["fps1"], g["fps1"])
    assert np.array_equal(out["group1"], g["group1"])
    assert np.array_equal(out["fps2"], g["fps2"])
    assert np.array_equal(out["group2"], g["group2"])
    # fp32 MLP: different GEMM blocking than oneDNN -> tolerance, SURVEY.md A.7
    np.testing.assert_allclose(out["l1_points"], g["l1_points"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(out["l2_points"], g["l2_points"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(out["feature"], g["feature"], rtol=1e-4, atol=1e-5)


def test_c1_chamfer_minima_bit_exact(golden, oracle):
    g = golden("c1_chamfer")
    assert np.array_equal(bits(oracle.nn_min(g["pred"], g["target"], 0)), bits(g["loss_rowmin"]))
    assert np.array_equal(bits(oracle.nn_min(g["target"], g["pred"], 0)), bits(g["loss_colmin"]))
    cd = oracle.chamfer_distance_chunked_optimized(g["pred"], g["target"])
    np.testing.assert_allclose(cd, g["chamfer_loss"], rtol=1e-6)
    np.testing.assert_allclose(cd, g["chamfer_loss_chunk100"], rtol=1e-6)


def test_c1_metrics_within_one_ulp(golden, oracle):
    # torch's vectorised CPU sqrt is not correctly rounded: 1-ulp tolerance (rtol 1.2e-7)
    g = golden("c1_chamfer")
    np.testing.assert_allclose(oracle.nn_min(g["pred"], g["target"], 1), g["metric_rowmin"], rtol=1.2e-7)
    np.testing.assert_allclose(oracle.nn_min(g["target"], g["pred"], 2), g["metric_colmin"], rtol=1.2e-7)
    np.testing.assert_allclose(oracle.metrics_chamfer_distance(g["pred"], g["target"]), g["metric_cd"], rtol=1e-6)
    np.testing.assert_allclose(oracle.metrics_chamfer_distance(g["pred"], g["target"], False), g["metric_cd_oneway"], rtol=1e-6)
    np.testing.assert_allclose(oracle.metrics_hausdorff_distance(g["pred"], g["target"]), g["metric_hausdorff"], rtol=1.2e-7)


def test_lattice_ties(golden, oracle):
    g = golden("lattice")
    fps = oracle.farthest_point_sample(g["x"], 256, g["start"])
    assert np.array_equal(fps, g["fps"])
    new_xyz = oracle.index_points(g["x"], fps)
    assert np.array_equal(oracle.query_ball_point(0.25, 16, g["x"], new_xyz), g["group"])
    assert np.array_equal(bits(oracle.nn_min(g["x"], g["y"], 0)), bits(g["loss_rowmin"]))
    assert np.array_equal(bits(oracle.nn_min(g["y"], g["x"], 0)), bits(g["loss_colmin"]))


def test_edge_ball_query(golden, oracle):
    g = golden("edge_ball_query")
    tiny = oracle.query_ball_point(float(g["r_tiny"]), 8, g["x"], g["q"])
    assert np.array_equal(tiny, g["tiny"])
    assert (tiny[0, 5:] == 300).all()  # empty balls -> sentinel N
    assert np.array_equal(oracle.query_ball_point(float(g["r_huge"]), 300, g["x"], g["q"][:, :5]), g["huge"])


def test_edge_fps(golden, oracle):
    g = golden("edge_fps")
    assert np.array_equal(oracle.farthest_point_sample(g["x"], 100, g["start"]), g["fps"])
    assert np.array_equal(oracle.farthest_point_sample(np.zeros((1, 50, 3), np.float32), 10, g["same_start"]), g["same_fps"])


@pytest.mark.parametrize("name", ["lidar", "uniform"])
def test_c2_120k_indices_bit_exact(golden, oracle, name):
    g = golden("c2_120k_" + name)
    cloud = S.lidar_scan(0) if name == "lidar" else S.uniform_cloud(0, 1, 120000)
    assert hashlib.sha256(cloud.numpy().tobytes()).hexdigest() == str(g["sha256"]), "synthetic generator drifted"
    x = cloud.numpy()
    f1 = oracle.farthest_point_sample(x, 512, g["start1"])
    assert np.array_equal(f1, g["fps1"])
    c1 = oracle.index_points(x, f1)
    assert np.array_equal(oracle.query_ball_point(0.2, 32, x, c1), g["group1"])
    f2 = oracle.farthest_point_sample(c1, 128, g["start2"])
    assert np.array_equal(f2, g["fps2"])
    assert np.array_equal(oracle.query_ball_point(0.4, 64, c1, oracle.index_points(c1, f2)), g["group2"])


def test_upsample_knn_and_sklearn_metrics(golden, oracle):
    g = golden("upsample_knn")
    out = oracle.upsample_knn(g["coarse"], g["original"], g["indices"])
    np.testing.assert_allclose(out, g["out"], rtol=1e-6, atol=1e-7)
    assert abs(oracle.coverage_score(g["pred"], g["target"], 0.05) - float(g["coverage_005"])) < 1e-12
    assert abs(oracle.coverage_score(g["pred"], g["target"], 0.01) - float(g["coverage_001"])) < 1e-12
    assert abs(oracle.uniformity_score(g["pred"], 8) - float(g["uniformity_8"])) < 1e-9
    assert abs(oracle.uniformity_score(g["target"], 4) - float(g["uniformity_4"])) < 1e-9


def test_voxel_downsample_oracle_against_reference_golden(golden, oracle):
    """models/diffusion_model.py:69-122: representatives (deterministic part) bit-exact; the full function with the
    reference's RNG stream reproduces the reference's indices."""
    import torch

    g = golden("voxel_downsample")
    target = int(g["target"])
    for b in range(2):
        pts = g["clouds"][b]
        vs = oracle.voxel_size_like_reference(pts, target)
        assert vs.tobytes() == g["voxel_size"][b].tobytes()
        assert np.array_equal(oracle.voxel_representatives(pts, pts.min(axis=0), vs), g["rep%d" % b])
    torch.manual_seed(77)
    idx = oracle.voxel_grid_downsample(g["clouds"], target, lambda n: torch.randperm(n).numpy())
    assert np.array_equal(idx, g["indices"])
    assert np.array_equal(oracle.voxel_grid_downsample(g["clouds"][:, :100], 200, None), g["small_indices"])


def test_voxel_size_scalar_arithmetic_matches_torch(oracle):
    """The oracle's numpy restatement of the 0-dim tensor expression (:80-84) against torch's own evaluation."""
    import torch

    rs = np.random.RandomState(3)
    for _ in range(300):
        n = int(rs.randint(50, 400))
        pts = (rs.randn(n, 3) * rs.uniform(0.01, 30.0, size=3)).astype(np.float32)
        target = int(rs.randint(5, 40))
        t = torch.from_numpy(pts)
        r = t.max(axis=0)[0] - t.min(axis=0)[0]
        r[r < 1e-6] = 1.0
        ref = ((r.prod() / target) ** (1 / 3) * 1.2).numpy()
        assert oracle.voxel_size_like_reference(pts, target).tobytes() == np.float32(ref).tobytes()
