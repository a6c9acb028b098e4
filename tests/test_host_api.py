"""Host-side logic of the drop-in layer that does not need a GPU: geometry, layouts, error behaviour
mirrored from kymatio's frontends, batch sharding and the gloo world_size-2 gather."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import wst_b200
import wst_b200.numpy
import wst_b200.torch


def test_geometry_matches_oracle():
    from oracle import compute_padding, num_coefficients
    for M in (32, 64, 128, 512):
        for J in (1, 2, 3, 4, 5):
            if 2 ** J <= M:
                assert wst_b200.compute_padding(M, M, J) == compute_padding(M, M, J)
            for L in (6, 8):
                for mo in (1, 2):
                    assert wst_b200.num_coefficients(J, L, mo) == num_coefficients(J, L, mo)


def test_layout_permutations():
    C, K = 3, 5
    f = np.arange(2 * C * 2 * K, dtype=np.float32).reshape(2, C, 2, K)
    blk = wst_b200.to_block(f)
    itl = wst_b200.to_interleaved(f)
    assert blk.shape == itl.shape == (2, C * 2 * K)
    for c in range(C):
        np.testing.assert_array_equal(itl[:, c * 2 * K:(c + 1) * 2 * K:2], blk[:, c * 2 * K:c * 2 * K + K])
        np.testing.assert_array_equal(itl[:, c * 2 * K + 1:(c + 1) * 2 * K:2], blk[:, c * 2 * K + K:(c + 1) * 2 * K])
    t = torch.from_numpy(f)
    np.testing.assert_array_equal(wst_b200.to_interleaved(t).numpy(), itl)


def test_frontend_constructor_errors():
    with pytest.raises(RuntimeError, match="smallest dimension should be larger than 2\\^J"):
        wst_b200.numpy.Scattering2D(J=6, shape=(32, 32))
    with pytest.raises(RuntimeError, match="out_type"):
        wst_b200.Scattering2D(J=2, shape=(32, 32), out_type="dict")
    with pytest.raises(RuntimeError, match="frontend"):
        wst_b200.Scattering2D(J=2, shape=(32, 32), frontend="jax")
    S = wst_b200.numpy.Scattering2D(J=2, shape=(32, 32))
    with pytest.raises(TypeError, match="NumPy array"):
        S([[1.0]])
    with pytest.raises(RuntimeError, match="at least two dimensions"):
        S(np.zeros(4, np.float32))
    with pytest.raises(RuntimeError, match="spatial size \\(32,32\\)"):
        S(np.zeros((16, 16), np.float32))
    St = wst_b200.torch.Scattering2D(J=2, shape=(32, 32))
    with pytest.raises(TypeError, match="PyTorch Tensor"):
        St(np.zeros((32, 32), np.float32))
    with pytest.raises(RuntimeError, match="contiguous"):
        St(torch.zeros(32, 64)[:, ::2])
    assert len(S._meta()) == 81 and S._meta()[17]["n"] == (0, 8)


def test_shard_ranges_cover_batch():
    for B in (0, 1, 7, 8, 1000, 1_000_000):
        for G in (1, 2, 4, 8):
            r = [wst_b200.shard_range(B, g, G) for g in range(G)]
            assert r[0][0] == 0 and r[-1][1] == B
            assert all(r[i][1] == r[i + 1][0] for i in range(G - 1))
            assert max(h - l for l, h in r) - min(h - l for l, h in r) <= 1
    with pytest.raises(ValueError):
        wst_b200.shard_range(10, 2, 2)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _gather_worker(rank, world, port, B, F, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    full = torch.arange(B * F, dtype=torch.float32).reshape(B, F)
    lo, hi = wst_b200.shard_range(B, rank, world)
    got = wst_b200.gather_features(full[lo:hi].clone(), B)
    if B % world == 0:                       # equal shards: the asynchronous form into a caller-provided matrix
        out = torch.full((B, F), -1.0)
        res, work = wst_b200.gather_features(full[lo:hi].clone(), B, async_op=True, out=out)
        work.wait()
        assert res is out and torch.equal(out, got)
    else:
        try:
            wst_b200.gather_features(full[lo:hi].clone(), B, async_op=True)
            raise AssertionError("ragged async gather must be refused")
        except ValueError:
            pass
    torch.save(got, os.path.join(out_dir, "r%d.pt" % rank))
    dist.destroy_process_group()


@pytest.mark.parametrize("B", [7, 8, 1])
def test_gloo_gather_world2(tmp_path, B):
    """N>1 path on CPU: ragged contiguous shards, gathered feature matrix is in input order on every rank."""
    F, world = 6, 2
    mp.spawn(_gather_worker, args=(world, _free_port(), B, F, str(tmp_path)), nprocs=world, join=True)
    full = torch.arange(B * F, dtype=torch.float32).reshape(B, F)
    for r in range(world):
        assert torch.equal(torch.load(os.path.join(tmp_path, "r%d.pt" % r)), full)


def test_feature_names_match_reference_contract():
    """train_and_save_model.py:400-427: the names the trainer zips with the feature columns."""
    adv = wst_b200.get_feature_names("advanced_stats")
    wst = wst_b200.get_feature_names("wst")
    hyb = wst_b200.get_feature_names("hybrid")
    assert len(adv) == 54 and adv[0] == "R_mean" and adv[17] == "R_edge_density" and adv[18] == "G_mean"
    assert len(wst) == 486 and wst[0] == "R_wst_mean_0" and wst[81] == "R_wst_std_0" and wst[162] == "G_wst_mean_0"
    assert hyb == adv + wst and len(hyb) == 540
    assert len(wst_b200.get_feature_names("wst", K=217)) == 3 * 2 * 217
    with pytest.raises(ValueError, match="Unknown feature method: foo"):
        wst_b200.get_feature_names("foo")
    ref_path = "/root/reference/src/training/train_and_save_model.py"
    if os.path.exists(ref_path):                       # build container only: the reference's own function
        import importlib.util
        spec = importlib.util.spec_from_file_location("ref_train", ref_path)
        ref = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(ref)
        for m in ("advanced_stats", "wst", "hybrid"):
            assert wst_b200.get_feature_names(m) == ref.get_feature_names(m)


def test_dispatchers_reject_unknown_methods_before_touching_the_gpu():
    img = np.zeros((3, 32, 32), np.float32)
    with pytest.raises(ValueError, match="Unknown feature method: nope"):
        wst_b200.extract_features(img, "nope")
    with pytest.raises(ValueError, match="Unknown feature method: nope"):
        wst_b200.extract_features_inference(img, "nope")
    m = wst_b200.ModelInferenceFeatures()
    m.feature_method = "nope"
    with pytest.raises(ValueError, match="Unknown feature method"):
        m.extract_features(img)


def test_torch_library_ops_are_registered_with_fake_kernels():
    """the batched ops are visible to the dispatcher; shapes propagate without a device."""
    from torch._subclasses.fake_tensor import FakeTensorMode
    assert str(torch.ops.wst.scattering2d_features.default._schema).startswith(
        "wst::scattering2d_features(Tensor x, SymInt J, SymInt L, SymInt max_order, SymInt layout, bool full_maps)")
    with FakeTensorMode():
        x = torch.empty((5, 3, 128, 128), device="cuda")
        assert torch.ops.wst.scattering2d_features(x, 2, 8, 2, 0, False).shape == (5, 486)
        assert torch.ops.wst.scattering2d_features(x, 4, 8, 2, 1, True).shape == (5, 3, 417, 8, 8)
        assert torch.ops.wst.scattering2d_maps(x, 2, 8, 2).shape == (5, 3, 81, 32, 32)
        r = torch.empty((2, 1, 120, 128), device="cuda")                       # rectangular: 32 x 32 outputs at J=2
        assert torch.ops.wst.scattering2d_maps(r, 2, 8, 2).shape == (2, 1, 81, 30, 32)
        u = torch.empty((7, 64, 64, 3), device="cuda", dtype=torch.uint8)      # load_rgb_image's input order
        assert torch.ops.wst.scattering2d_features(u, 3, 8, 2, 0, False).shape == (7, 1302)
    with pytest.raises(NotImplementedError):                                    # no CPU kernel: no CPU fallback
        torch.ops.wst.scattering2d_features(torch.zeros(1, 3, 32, 32), 2, 8, 2, 0, False)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        wst_b200.scattering_features(torch.zeros(1, 3, 32, 32), 2)


def test_device_index_resolution():
    from wst_b200._api import _device_index
    assert _device_index(3) == 3 and _device_index("cuda:2") == 2 and _device_index(torch.device("cuda", 1)) == 1
    with pytest.raises(RuntimeError, match="CUDA devices only"):
        _device_index("cpu")
