"""CPU: the C-ABI shared library loads and exports every symbol include/pcst.h declares; argument
validation paths that need no GPU return the documented status codes."""
import ctypes
import os
import re

import pytest

from pointcloud_style_transfer_b200 import _lib

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(REPO, "include", "pcst.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pcst_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    from pointcloud_style_transfer_b200.build import build

    build()
    return _lib.load()


def test_every_declared_symbol_is_exported_and_bound(lib):
    syms = header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/pcst.h but not exported by libpcst.so"
    assert sorted(_lib.SIGNATURES) == syms, "ctypes prototypes and include/pcst.h disagree"


def test_version_and_error_string(lib):
    assert lib.pcst_version().decode().startswith("pcst ")
    assert isinstance(_lib.last_error(), str)


def test_invalid_arguments_return_status_not_crash(lib):
    null = ctypes.c_void_p(0)
    assert lib.pcst_fps_f32(null, 1, 16, 4, null, null, null, null, 0, null) == -1
    assert "null" in _lib.last_error()
    assert lib.pcst_ball_query_f32(null, null, 1, 16, 4, ctypes.c_float(0.1), 4, null, null, 0, null) == -1
    assert lib.pcst_nn_min_f32(null, null, 1, 16, 16, 0, null, null, null, 0, null) == -1
    assert lib.pcst_knn_f32(null, null, 1, 4, 4, 2, null, null, null, 0, null) == -1
    one = ctypes.c_void_p(256)  # never dereferenced: the size checks fail first
    assert lib.pcst_nn_min_f32(one, one, 1, 0, 16, 0, one, null, null, 0, null) == -1
    assert lib.pcst_nn_min_f32(one, one, 1, 16, 16, 7, one, null, null, 0, null) == -1
    assert lib.pcst_ball_query_f32(one, one, 1, 16, 4, ctypes.c_float(0.1), 17, one, null, 0, null) == -1  # nsample > N
    assert lib.pcst_knn_f32(one, one, 1, 4, 4, 17, one, one, null, 0, null) == -1
    # workspace too small -> PCST_ERR_WORKSPACE
    assert lib.pcst_nn_min_f32(one, one, 1, 16, 16, 0, one, null, null, 0, null) == -4
    with pytest.raises(_lib.PcstError):
        _lib.check(-4)


def test_workspace_queries(lib):
    assert lib.pcst_nn_min_workspace_bytes(1, 120000, 120000) >= 2 * 120000 * 16
    assert lib.pcst_ball_query_workspace_bytes(1, 120000, 512) >= 120000 * 16
    assert lib.pcst_fps_workspace_bytes(1, 120000, 512) == 0          # register-resident
    assert lib.pcst_fps_workspace_bytes(1, 500000, 512) >= 500000 * 4  # streaming fallback keeps distances in HBM
    assert lib.pcst_nn_min_workspace_bytes(0, 1, 1) == 0


def test_shared_mlp_plan_queries_are_host_only(lib):
    """The launch plan of the shared MLP (cluster split, packed-blob size) is plain host arithmetic: no GPU needed.
    Shapes: the three stages of the reference encoder (models/pointnet2_encoder.py:117-119) on 1 and 32 scans."""
    import ctypes
    C3 = ctypes.c_int * 3
    sa1, sa2, sa3 = (512, 32, 0, (64, 64, 128)), (128, 64, 128, (128, 128, 256)), (1, 128, 256, (256, 512, 256))
    for S_, K, D, c in (sa1, sa2, sa3):
        assert lib.pcst_sa_mlp_pick_cluster(1, S_, K, D, C3(*c), 0) == 1           # fp32 path never splits
        assert lib.pcst_sa_mlp_pick_cluster(32, S_, K, D, C3(*c), 1) in (1, 2, 4)  # many tiles: little or no split
    # one scan: SA1 has 128 row tiles (no room to split), SA2 64 (two CTAs per tile), SA3 one (eight)
    assert [lib.pcst_sa_mlp_pick_cluster(1, S_, K, D, C3(*c), 1) for S_, K, D, c in (sa1, sa2, sa3)] == [1, 2, 8]
    for S_, K, D, c in (sa1, sa2, sa3):
        kp0 = -(-(3 + D) // 16) * 16
        weights = 2 * (kp0 * c[0] + c[0] * c[1] + c[1] * c[2])  # bf16, layer-0 K padded to 16
        for cl in (1, 2, 4):
            n = lib.pcst_sa_mlp_packed_bytes(D, C3(*c), 1, cl)
            assert n >= weights + 2 * 4 * sum(c) and n < 1.1 * weights + 8 * sum(c) + 4096
        assert lib.pcst_sa_mlp_packed_bytes(D, C3(*c), 0, 1) >= 4 * ((3 + D) * c[0] + c[0] * c[1] + c[1] * c[2])
    assert lib.pcst_sa_mlp_packed_bytes(0, C3(64, 64, 128), 1, 8) == 0              # 64 / 8 columns per CTA: no such plan
    assert lib.pcst_sa_mlp_packed_bytes(0, C3(64, 64, 100), 1, 1) == 0              # widths must be multiples of 32
    # wider than the tensor-core path takes (Cout_2 > 512): the fp32 layout is what both precisions pack
    assert lib.pcst_sa_mlp_packed_bytes(0, C3(64, 64, 1024), 1, 1) == lib.pcst_sa_mlp_packed_bytes(0, C3(64, 64, 1024), 0, 1)


def test_denoiser_plan_queries_are_host_only(lib):
    """The fused NoisePredictor's step table (csrc/noise_mlp_tc.cu) is host arithmetic: blob size, workspace and the
    pack's launch count for the reference's configuration (feature_dim 256, time_embed_dim 128, 6 blocks,
    models/diffusion_model.py:38-61) and the limits of what the kernel takes."""
    F, T, nb = 256, 128, 6
    weights = 2 * (16 * 128 + 128 * 256 + 256 * F + nb * 2 * (F * 2 * F) + F * 256 + 256 * 128 + 128 * 16)   # bf16, K padded
    side = 4 * (F * T + F * F + 3 * F + nb * F + 128 + 256 + nb * 2 * F + 256 + 128 + 16)                  # fp32 vectors
    n = lib.pcst_noise_predictor_packed_bytes(F, T, nb)
    assert weights + side <= n <= weights + side + 64 * 1024
    # 4 point-encoder steps (the 256-wide layer in two N chunks), 8 per block (4 hidden chunks x 2), 4 output steps;
    # plus the bias / projection copies
    assert lib.pcst_noise_predictor_pack_launches(F, T, nb) == (4 + 8 * nb + 4) + 5 + nb
    assert lib.pcst_noise_predictor_workspace_bytes(2, F, nb) >= 2 * (nb + 1) * F * 4
    for bad in ((250, T, nb), (272, T, nb), (F, 127, nb), (F, T, 9), (8, T, nb)):
        assert lib.pcst_noise_predictor_packed_bytes(*bad) == 0
        assert lib.pcst_noise_predictor_pack_launches(*bad) == 0
    # narrow widths: the hidden layer is one chunk of 2F <= 128 columns (2 steps per block)
    assert lib.pcst_noise_predictor_pack_launches(64, 32, 2) == (4 + 2 * 2 + 4) + 5 + 2
    assert lib.pcst_chamfer_shard_payload_floats(120000) == 120000 + 256


def test_denoiser_step_table_is_consistent_for_every_supported_size(lib):
    """The fused denoiser's schedule (MMA steps running one hidden chunk ahead of the epilogues, events alternating between
    two mbarriers) is derived on the host from the buffers each step touches; the library's self-check replays it for every
    feature width and block count the kernel accepts -- including hidden layers that split into 1, 2, 3 or 4 chunks and
    sizes no GPU test runs -- and reports the first violated dependency rule (deadlock, parity ambiguity, RAW, WAR, range)."""
    for F in range(16, 257, 16):
        for nb in range(0, 9):
            assert lib.pcst_noise_predictor_plan_selfcheck(F, 128, nb) == 0, (F, nb)
    assert lib.pcst_noise_predictor_plan_selfcheck(250, 128, 6) == -1


def test_tuning_knobs(lib):
    _lib.set_tuning("nn_min.splits", 3)
    assert _lib.get_tuning("nn_min.splits") == 3
    _lib.set_tuning("nn_min.splits", 0)
    with pytest.raises(_lib.PcstError):
        _lib.set_tuning("no.such.knob", 1)


def test_pcst_tune_environment_is_applied_at_load():
    """PCST_TUNE="key=value,..." sets knobs for a whole process when the library is loaded; an unknown key fails loudly."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("from pointcloud_style_transfer_b200 import _lib; _lib.load(); "
            "print(_lib.get_tuning('fps.cluster'), _lib.get_tuning('sa_mlp.regs'))")
    env = dict(os.environ, PCST_TUNE="fps.cluster=8, sa_mlp.regs=168", PYTHONPATH=root)
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, cwd=root)
    assert out.returncode == 0, out.stderr
    assert out.stdout.split() == ["8", "168"]
    env["PCST_TUNE"] = "no.such.knob=1"
    bad = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, cwd=root)
    assert bad.returncode != 0 and "unknown key" in bad.stderr


def test_cpu_tensors_raise_loudly():
    import torch

    from pointcloud_style_transfer_b200.models.pointnet2_encoder import farthest_point_sample, query_ball_point
    from pointcloud_style_transfer_b200.models.losses import chamfer_distance_chunked_optimized

    x = torch.rand(1, 64, 3)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        farthest_point_sample(x, 8)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        query_ball_point(0.2, 8, x, x[:, :4])
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        chamfer_distance_chunked_optimized(x, x)
