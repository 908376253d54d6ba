"""CPU tests (no GPU needed): the oracle against the golden vectors of the unmodified reference, the
reference's only shipped known answer, host-side logic, and the C-ABI surface of the library."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from oracle import spair_oracle as so
from tests import helpers
from tests.helpers import assert_close

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_backbone_geometry_known_answer():
    """The one known answer the reference ships (test_notebook.ipynb:350)."""
    g = so.backbone_geometry(so.DEFAULT_TOPOLOGY, (128, 128))
    assert list(g["rf_size"]) == [31, 31] and list(g["grid_cell_size"]) == [12, 12]
    assert list(g["n_grid_cells"]) == [11, 11] and list(g["pre_padding"]) == [9, 9]
    assert list(g["post_padding"]) == [14, 14] and list(g["required_image_size"]) == [151, 151]
    from spair_pytorch_b200.modules import receptive_field_geometry
    pad, n_cells, cell = receptive_field_geometry(so.DEFAULT_TOPOLOGY, (128, 128))
    assert pad == (9, 14, 9, 14) and list(n_cells) == [11, 11] and list(cell) == [12, 12]
    pad, n_cells, cell = receptive_field_geometry(so.CELL8_TOPOLOGY, (256, 256))
    assert pad == (7, 7, 7, 7) and list(n_cells) == [32, 32] and list(cell) == [8, 8]


@pytest.mark.parametrize("name,step", [("tiny", 1), ("tiny", 1001), ("A", 1001)])
def test_oracle_reproduces_reference_golden(name, step):
    """The restatement in oracle/spair_oracle.py, fed the golden's inputs and the seeded parameters,
    reproduces what the unmodified reference produced (bit-exact forward, gradients to rounding)."""
    g = helpers.load_golden("model_%s_step%d.npz" % (name, step))
    net = helpers.build_model(name)
    params = so.params_from_state_dict(net.state_dict())
    noise = so.Noise(*(torch.from_numpy(g[k]) for k in ("eps_where", "eps_attr", "eps_depth", "u_pres")))
    out = so.forward_backward(params, torch.from_numpy(g["x"]), int(g["step"]), noise, helpers.oracle_config(name))
    for key in ("loss", "recon_x", "z_where", "z_pres", "z_attr", "z_depth"):
        assert np.array_equal(out[key].detach().numpy(), g[key]), key
    for n, m in out["kl"].items():
        assert np.array_equal(m.detach().numpy(), g["kl/" + n]), n
    for k, p in params.items():
        if k.startswith("attn."):
            assert p.grad is None and "gnone/" + k in g.files
            continue
        got = p.grad.flatten()[torch.from_numpy(g["gidx/" + k])]
        assert_close(got, g["gval/" + k], "grad " + k, rtol=1e-5, atol=1e-6 * float(g["gstat/" + k][1]))


def test_draw_noise_replays_reference_stream():
    """Golden noise tensors were captured next to a real reference run seeded with noise_seed."""
    g = helpers.load_golden("model_tiny_step1.npz")
    n = so.draw_noise(int(g["noise_seed"]), g["x"].shape[0], (5, 5), 50)
    assert np.array_equal(n.eps_where.numpy(), g["eps_where"]) and np.array_equal(n.u_pres.numpy(), g["u_pres"])


def test_product_reference_noise_stream_matches_reference_golden():
    """SPAIR.set_reference_noise (product side, no oracle import): seeding the CPU generator like the reference run that
    produced the golden and replaying its draw order gives the golden's noise tensors bit for bit, i.e. the drop-in can be
    fed "the same inputs and seeds" as the reference without test infrastructure."""
    for name in ("tiny", "A"):
        g = helpers.load_golden("model_%s_step1.npz" % name)
        net = helpers.build_model(name, "cpu")
        torch.manual_seed(int(g["noise_seed"]))
        net.set_reference_noise(g["x"].shape[0])
        eps_where, eps_attr, eps_depth, u_pres = net._noise           # image-major [B, HW, k] copies held for the next forward
        B = g["x"].shape[0]

        def img_major(a):
            t = torch.from_numpy(a)
            return t.permute(0, 2, 3, 1).reshape(B, -1, t.shape[1])

        assert torch.equal(eps_where, img_major(g["eps_where"])) and torch.equal(eps_attr, img_major(g["eps_attr"]))
        assert torch.equal(eps_depth, img_major(g["eps_depth"]).squeeze(-1)) and torch.equal(u_pres, img_major(g["u_pres"]).squeeze(-1))


def test_closed_form_kl_scan_matches_oracle():
    """fp64 restatement of the count-prior recurrence against the oracle's op sequence."""
    torch.manual_seed(0)
    B, Hc = 3, 6
    HW = Hc * Hc
    pres = torch.rand(B, 1, Hc, Hc)
    cfg = so.OracleConfig()
    for step in (1, 1001, 5000):
        kl = so.compute_kl({}, {}, pres, pres, step, cfg)["pres_dist"].view(B, HW).double()
        _, cd0, _ = so.count_prior_distribution(step, HW, cfg)
        for b in range(B):
            cd, k = cd0.double().clone(), 0.0
            for i in range(HW):
                q = (torch.arange(HW + 1, dtype=torch.float64) - k).clamp(0, HW - i) / (HW - i)
                pz, pi = float((cd * q).sum()), float(pres.view(B, HW)[b, i])
                want = pi * (np.log(pi + 1e-9) - np.log(pz + 1e-9)) + (1 - pi) * (np.log(1 - pi + 1e-9) - np.log(1 - pz + 1e-9))
                assert abs(want - float(kl[b, i])) <= 1e-5 + 1e-5 * abs(want)
                s = float(torch.round(torch.tensor(pi)))
                cd = cd * (s * q + (1 - s) * (1 - q))
                cd = cd / cd.sum().clamp(min=1e-6)
                k += s


# ------------------------------------------------------------------------------------------
# host logic
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("L", [1, 2, 3])
def test_context_offsets_match_reference_order(L):
    from spair_pytorch_b200.schedule import context_offsets
    assert context_offsets(L) == so.context_offsets(L)
    assert len(context_offsets(L)) == (2 * L + 1) ** 2 // 2


@pytest.mark.parametrize("Hc,Wc,L,T", [(11, 11, 1, 31), (16, 16, 1, 46), (32, 32, 1, 94), (5, 7, 2, 19), (1, 9, 1, 9)])
def test_wavefront_schedule(Hc, Wc, L, T):
    from spair_pytorch_b200.schedule import build_schedule
    s = build_schedule(Hc, Wc, L)
    assert s.n_wavefronts == T == Wc + (L + 1) * (Hc - 1)
    assert sorted(s.order.tolist()) == list(range(Hc * Wc))
    assert all(s.wf_pos[s.order[i]] == i for i in range(Hc * Wc))
    done = set()
    for t in range(T):                      # every dependency is satisfied by an earlier wavefront
        cells = s.cells_of(t).tolist()
        for c in cells:
            h, w = divmod(c, Wc)
            for dh, dw in s.offsets:
                if 0 <= h + dh < Hc and 0 <= w + dw < Wc:
                    assert (h + dh) * Wc + (w + dw) in done
        done.update(cells)
    assert s.max_cells <= (Wc + L) // (L + 1) + 1


def test_state_dict_layout_is_the_reference_checkpoint_abi():
    net = helpers.build_model("A")
    sd = net.state_dict()
    assert sum(v.numel() for v in sd.values()) == 1462260
    expect = {"virtual_edge_element": (56,), "backbone.net.conv_0.weight": (128, 1, 4, 4), "backbone.net.conv_out.bias": (100,),
              "box_network.body.dense0.weight": (100, 324), "box_network.output_layers.0.weight": (8, 100),
              "box_network.output_layers.1.bias": (100,), "object_encoder.dense0.weight": (256, 784),
              "object_encoder.out.weight": (100, 128), "z_network.body.dense0.weight": (100, 478),
              "z_network.output_layers.0.bias": (2,), "obj_network.dense0.weight": (100, 479), "obj_network.out.weight": (1, 100),
              "object_decoder.out.weight": (1568, 256), "attn.gamma": (1,), "attn.query_conv.weight": (6, 55, 1, 1)}
    for k, shape in expect.items():
        assert tuple(sd[k].shape) == shape, k
    g = helpers.load_golden("model_A_step1.npz")
    assert sorted("psum/" + k for k in sd) == sorted(f for f in g.files if f.startswith("psum/"))


def test_drop_in_package_names():
    """What reference train.py:12-16 imports must resolve, and cfg must be one shared module."""
    import spair
    from spair import config as cfg
    from spair import debug_tools, metric  # noqa: F401
    from spair.dataloader import SimpleScatteredMNISTDataset  # noqa: F401
    from spair.models import SPAIR  # noqa: F401
    import spair.modules as m
    import spair_pytorch_b200.config as impl_cfg
    assert cfg is impl_cfg
    for name in ("Backbone", "build_MLP", "SequentialMultipleOutput", "latent_to_mean_std", "clamped_sigmoid",
                 "exponential_decay", "stn", "to_C_H_W", "to_H_W_C", "safe_log", "compute_backbone_feature_shape"):
        assert hasattr(m, name), name
    with pytest.raises(AssertionError):
        m.to_H_W_C(torch.zeros(2, 4, 4, 4))       # reference modules.py:293 rejects C == H


def test_scalar_helpers_match_oracle():
    from spair_pytorch_b200 import modules as m
    t = torch.linspace(-15, 15, 62).view(-1, 2)
    for a, b in zip(m.latent_to_mean_std(t), so.latent_to_mean_std(t)):
        assert torch.equal(a, b)
    assert torch.equal(m.clamped_sigmoid(t), so.clamped_sigmoid(t))
    assert torch.equal(m.clamped_sigmoid(t, True), so.clamped_sigmoid(t, True))
    for step in (0, 1, 999, 1000, 1001, 5000):
        for kw in (so.OracleConfig().wheel, so.OracleConfig().count_prior):
            assert torch.equal(m.exponential_decay(step, "cpu", **kw), so.exponential_decay(step, **kw))


# ------------------------------------------------------------------------------------------
# C-ABI surface
# ------------------------------------------------------------------------------------------
def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "spair_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for m in re.finditer(r"\bint\s+(spair_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        args = [a.strip() for a in m.group(2).split(",") if a.strip() and a.strip() != "void"]
        out[m.group(1)] = [ctypes.c_void_p if "*" in a else (ctypes.c_float if a.startswith("float") else ctypes.c_int)
                           for a in args]
    return out


def test_library_exports_every_declared_symbol_and_binding_matches_header():
    """No compute call: the shared library loads, exports everything include/spair_b200.h declares,
    and the ctypes signatures in kernels.py agree with the header argument by argument."""
    from spair_pytorch_b200 import kernels as K
    declared = _declared_symbols()
    assert len(declared) >= 20
    handle = K.lib()
    for name, kinds in declared.items():
        assert hasattr(handle, name), "missing export " + name
        assert K._SIGNATURES[name] == kinds, "binding of %s disagrees with the header" % name
    assert set(K._SIGNATURES) == set(declared)
    header = open(os.path.join(ROOT, "include", "spair_b200.h")).read()
    header_version = int(re.search(r"#define\s+SPAIR_ABI_VERSION\s+(\d+)", header).group(1))
    assert handle.spair_abi_version() == header_version == K.ABI_VERSION


def test_base_grid_matches_torch_affine_grid_bit_exactly():
    from spair_pytorch_b200 import kernels as K
    for n in list(range(1, 70)) + [128, 255, 256, 1024]:
        assert torch.equal(K.base_grid(n), torch.linspace(-1, 1, n) * (n - 1) / n), n


def test_invalid_arguments_rejected_without_gpu():
    from spair_pytorch_b200 import kernels as K
    assert K.lib().spair_base_grid(0, None) == -1
    assert K.lib().spair_render_num_tiles(2, 128, 128) == 2 * 4 * 4      # 32x32 canvas tiles


def test_device_scene_generator_schema():
    """Procedural scenes: values in [0,1], boxes inside the canvas, counts in 1..max — same item schema
    as the reference's HDF5 dataset (dataloader.py:23-33)."""
    from spair_pytorch_b200.dataloader import scattered_sprites, scattered_sprites_gpu
    g = torch.Generator().manual_seed(0)
    for shape in ((1, 128, 128), (3, 64, 64)):
        img, bbox, count = scattered_sprites_gpu(6, shape, "cpu", g, max_sprites=9, sprite_px=(8, 20))
        assert img.shape == (6,) + shape and float(img.min()) >= 0.0 and float(img.max()) <= 1.0
        assert bbox.shape == (6, 9, 4) and count.shape == (6, 1) and 1 <= float(count.min()) and float(count.max()) <= 9
        live = bbox[..., 2] > 0
        assert bool(((bbox[..., 0] + bbox[..., 2])[live] <= shape[2]).all()) and bool(((bbox[..., 1] + bbox[..., 3])[live] <= shape[1]).all())
        assert int(live.sum()) == int(count.sum())
        assert float(img.flatten(1).max(1).values.min()) > 0.5        # every image has at least one visible sprite
    img2, box2, cnt2 = scattered_sprites(3, (1, 128, 128), seed=1, return_boxes=True)
    assert img2.shape == (3, 1, 128, 128) and box2.shape == (3, 9, 4) and cnt2.shape == (3, 1)
